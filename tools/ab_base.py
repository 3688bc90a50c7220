"""A/B helper (dev tool, build container): build/ab/libppnn_base.so = the library with the listed .cu files taken
from a git revision (default HEAD) and every other object from the current build, so that one gpurun call can time
both on the same box:  PPNN_B200_LIB=build/ab/libppnn_base.so python tools/run_fine.py ...

    python tools/ab_base.py [--rev REV] fine_step.cu [other.cu ...]
"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as G  # noqa: E402

args = sys.argv[1:]
rev = "HEAD"
if args and args[0] == "--rev":
    rev, args = args[1], args[2:]
src = os.path.join(ROOT, "build", "base_src")
out = os.path.join(ROOT, "build", "ab")
os.makedirs(src, exist_ok=True)
os.makedirs(out, exist_ok=True)
subprocess.run(f"git -C {ROOT} archive {rev} ppnn_b200/csrc include | tar -x -C {src}", shell=True, check=True)
objs = []
for cu in sorted(glob.glob(os.path.join(G.CSRC, "*.cu"))):
    name = os.path.basename(cu)
    o = os.path.join(ROOT, "build", "obj", name[:-3] + ".o")
    if name in args:
        bo = os.path.join(out, "base_" + os.path.basename(o))
        cmd = ["nvcc"] + G.NVCC_FLAGS + ["-c", "-o", bo, name]
        print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True, cwd=os.path.join(src, "ppnn_b200", "csrc"))
        objs.append(bo)
    else:
        objs.append(o)
lib = os.path.join(out, "libppnn_base.so")
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib] + objs, check=True)
print("built", lib, "from", rev, "for", args)
