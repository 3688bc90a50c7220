"""Time the coarse-step kernel variants in build/variants/*.so (dev tool, GPU box)."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, torch
sys.path.insert(0, %r)
import bench
from ppnn_b200 import CoarseStep, _lib as L
dev = torch.device("cuda")
F, B = 257, 1920
c = bench.CASE
step = CoarseStep(F, c["coarse"], c["rhsu"], c["rhsv"], c["dt"], c["length"])
bstep = CoarseStep(F, 33, "burgers2Dpu", "burgers2Dpv", 0.02, 3.2)
u = bench.synth_fields(B, F, dev, 10)
mu = bench.make_mu(B, dev).reshape(-1).contiguous()
nn = torch.randn(B, 2, F - 1, F - 1, device=dev)
out = torch.empty_like(u)
om, os_ = L.f2(bench.STATS["outmean"]), L.f2(bench.STATS["outstd"])
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
res = {}
res["update"] = t(lambda: step.step_update_(u, mu, nn, om, os_, out))
res["checksum"] = float(out.double().sum())
res["pde_only"] = t(lambda: step.step_update_(u, mu, None, None, None, out))
st = [L.f2([0.0, 0.0]), L.f2([0.3, 0.3]), L.f2([0.0, 0.0]), L.f2([1e-3, 1e-3])]
res["inputs"] = t(lambda: step.step_inputs(u, mu, *st))
res["burgers_update"] = t(lambda: bstep.step_update_(u, mu * 0.05, nn, om, os_, out))
res["pde_du"] = t(lambda: step.pde_du(u, mu))
import ctypes as C
lib = L.load()
x = u + 1e-3 * torch.randn_like(u)
partial = torch.zeros(B, 2, 6, dtype=torch.float64, device=dev)
outs = [torch.empty(B, 2, F - 1, F - 1, device=dev) for _ in range(3)]
m6, s6 = (C.c_float * 6)(0.1, 0.2, 0.0, 0.0, 0.0, 0.0), (C.c_float * 6)(0.3, 0.3, 1e-3, 1e-3, 1e-3, 1e-3)
for nm, st_ in (("rd", step), ("b33", bstep)):
    muv = mu if nm == "rd" else mu * 0.05
    res["stats_" + nm] = t(lambda: L.check(lib.ppnn_dataset_stats(C.byref(st_.cfg), u.data_ptr(), x.data_ptr(), muv.data_ptr(), B, 0,
                                                                   partial.data_ptr(), L.stream()), "stats"))
    res["emit_" + nm] = t(lambda: L.check(lib.ppnn_dataset_emit(C.byref(st_.cfg), u.data_ptr(), x.data_ptr(), muv.data_ptr(), B, 0, m6, s6,
                                                                 outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), L.stream()), "emit"))
res = {k: round(v, 4) if k != "checksum" else v for k, v in res.items()}
print(json.dumps(res))
''' % ROOT
for lib in sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so"))):
    for extra in ({},):
        env = dict(os.environ, PPNN_B200_LIB=lib, **extra)
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-600:]
        print(os.path.basename(lib), extra, line, flush=True)
