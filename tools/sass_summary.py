#!/usr/bin/env python
"""Per-kernel SASS opcode summary of libppnn_b200.so (evidence of what the hot kernels are made of):

    python tools/sass_summary.py > profiles/r2_sass_summary.md

Counts, per kernel of the shipped library, the mnemonics that matter on this path: UBLKCP (1-D
cp.async.bulk = TMA bulk copy), UTMALDG (tensor-map TMA; not used: 257-float pitches are not 16-byte
multiples), SYNCS (mbarrier), FFMA2 / FADD2 / FMUL2 (packed fp32x2), FFMA, LDS / STS, LDG / STG, RED / ATOM,
LDL / STL (local-memory spills), and tensor-core opcodes (none by design: the path is memory-bound)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ppnn_b200", "libppnn_b200.so")
COLS = ["UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "LDS", "STS", "LDG", "STG", "RED", "ATOM", "LDL", "STL",
        "HMMA", "UTCMMA", "total"]


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.split("\n")
    return [o.strip() for o in out[:len(names)]]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["total"] += 1
            for c in COLS[:-1]:
                if op == c or (c in ("RED", "ATOM") and op.startswith(c)) or (c == "HMMA" and op.endswith("MMA") and op != "UTCMMA" and "FFMA" not in op and op not in ("FFMA", "FFMA2", "DFMA", "IMMA") and "MMA" in op):
                    counts[cur][c] += 1
    names = list(counts)
    pretty = demangle(names)
    print("# SASS opcode summary of `ppnn_b200/libppnn_b200.so` (sm_100a)\n")
    print("`python tools/sass_summary.py` (cuobjdump -sass, static instruction counts per kernel; every template instantiation "
          "the library ships).  UBLKCP = 1-D `cp.async.bulk` (TMA bulk copy), SYNCS = mbarrier operations, FFMA2/FADD2/FMUL2 = "
          "packed fp32x2 arithmetic, LDL/STL = local-memory (spill) accesses, HMMA/UTCMMA = tensor-core opcodes (none by design).\n")
    print("| kernel | " + " | ".join(COLS) + " |")
    print("|---|" + "---|" * len(COLS))
    for n, p in sorted(zip(names, pretty), key=lambda t: t[1]):
        short = re.sub(r"^void ppnn::", "", p)
        # keep the template arguments, drop the parameter list: cut at the "(" that follows the closing ">" (or the name)
        m2 = re.match(r"^([A-Za-z0-9_:]+(?:<.*?>)?)\(ppnn::|^([A-Za-z0-9_:]+(?:<.*?>)?)\(", short)
        if m2:
            short = m2.group(1) or m2.group(2)
        short = short.replace("(int)", "").replace("(bool)", "")
        print(f"| `{short}` | " + " | ".join(str(counts[n][c]) for c in COLS) + " |")


if __name__ == "__main__":
    sys.exit(main())
