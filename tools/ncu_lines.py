"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export per CUDA source line (dev tool)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; h = None; data = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": h = r; continue
    if h is None or r[0] in ("", "Function Name"): continue
    try:
        ie = int(r[h.index("Instructions Executed")]); smp = int(r[h.index("# Samples")])
    except (ValueError, IndexError):
        continue
    data.append((ie, smp, cur, r[0], r[1].strip()[:100]))
ti = sum(d[0] for d in data) or 1; ts = sum(d[1] for d in data) or 1
print("total warp-instructions", ti, "samples", ts)
for d in sorted(data, reverse=True)[:top]:
    print(f"{d[0]/ti*100:5.1f}% inst {d[1]/ts*100:5.1f}% smp  {d[2]}:{d[3]:>4}  {d[4]}")
if len(sys.argv) > 3:      # line ranges "file:lo-hi,..." -> share of samples
    for spec in sys.argv[3].split(","):
        f, rg = spec.split(":"); lo, hi = map(int, rg.split("-"))
        s = sum(d[1] for d in data if d[2] == f and lo <= int(d[3]) <= hi); i = sum(d[0] for d in data if d[2] == f and lo <= int(d[3]) <= hi)
        print(f"{spec:45s} {s/ts*100:5.1f}% smp {i/ti*100:5.1f}% inst")
    byfile = {}
    for d in data: byfile[d[2]] = byfile.get(d[2], 0) + d[1]
    print({k: round(v / ts * 100, 1) for k, v in byfile.items()})
