"""Executed-instruction mix per SASS opcode from an `ncu --page source --print-source cuda,sass --csv` export (dev tool):
python tools/ncu_opmix.py export.csv [divisor]   -- divisor: e.g. warp-iterations of the hot loop, to get instructions per iteration."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
div = float(sys.argv[2]) if len(sys.argv) > 2 else None
h = None; agg = {}; smp = {}; tot = 0; ts = 0
for r in rows:
    if not r: continue
    if r[0] == "Line No": h = r; continue
    if h is None or len(r) < 8 or r[2] in ("-", "Address"): continue
    try:
        ie = int(r[h.index("Instructions Executed")]); s = int(r[h.index("# Samples")])
    except ValueError:
        continue
    toks = r[3].split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    parts = op.split(".")
    op = ".".join(parts[:2]) if parts[0] in ("IMAD", "LDS", "STG", "LDG", "STS") and len(parts) > 1 else parts[0]
    agg[op] = agg.get(op, 0) + ie; smp[op] = smp.get(op, 0) + s; tot += ie; ts += s
print("warp-instructions", tot, "samples", ts)
for k, v in sorted(agg.items(), key=lambda x: -x[1])[:40]:
    extra = f"  {v / div:7.2f} / iteration" if div else ""
    print(f"{k:14s} {v / tot * 100:5.1f}% inst {smp[k] / max(ts, 1) * 100:5.1f}% smp{extra}")
