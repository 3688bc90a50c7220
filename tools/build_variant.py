"""Dev tool (build container): build/variants/NAME.so = the current library with the listed .cu files recompiled with
extra -D flags, for same-box timing of compile-time choices (tools/tune_step.py walks build/variants/*.so).

    python tools/build_variant.py NAME file.cu[,file2.cu] -DPPNN_X=1 [-DPPNN_Y=2 ...]
"""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as G  # noqa: E402

name, files, flags = sys.argv[1], sys.argv[2].split(","), sys.argv[3:]
out = os.path.join(ROOT, "build", "variants")
os.makedirs(os.path.join(out, "obj_" + name), exist_ok=True)
objs, procs = [], []
for cu in sorted(glob.glob(os.path.join(G.CSRC, "*.cu"))):
    base = os.path.basename(cu)
    o = os.path.join(ROOT, "build", "obj", base[:-3] + ".o")
    if base in files:
        o = os.path.join(out, "obj_" + name, base[:-3] + ".o")
        procs.append(subprocess.Popen(["nvcc"] + G.NVCC_FLAGS + flags + ["-c", "-o", o, base], cwd=G.CSRC))
    objs.append(o)
assert all(p.wait() == 0 for p in procs)
lib = os.path.join(out, name + ".so")
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib] + objs, check=True)
print("built", lib, flags)
