"""Time the two data-set passes alone (dev tool, GPU box):  python tools/run_stats.py [rd|burgers33] [stats|emit]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from ppnn_b200 import CoarseStep, _lib as L
dev = torch.device("cuda")
case = sys.argv[1] if len(sys.argv) > 1 else "rd"
what = sys.argv[2] if len(sys.argv) > 2 else "stats"
B = 1920
step = {"rd": lambda: CoarseStep(257, 49, "rdpu", "rdpv", 0.002, 6.4),
        "burgers33": lambda: CoarseStep(257, 33, "burgers2Dpu", "burgers2Dpv", 0.02, 3.2)}[case]()
u = bench.synth_fields(B, 257, dev, 3)
x = u + 1e-3 * torch.randn_like(u)
mu = torch.full((B,), 0.5 if case == "rd" else 0.04, device=dev)
partial = torch.zeros(B, 2, 6, dtype=torch.float64, device=dev)
outs = [torch.empty(B, 2, 256, 256, device=dev) for _ in range(3)] if what == "emit" else None
lib = L.load()
m6, s6 = (C.c_float * 6)(0.1, 0.2, 0.0, 0.0, 0.0, 0.0), (C.c_float * 6)(0.3, 0.3, 1e-3, 1e-3, 1e-3, 1e-3)
def run():
    if what == "emit":
        L.check(lib.ppnn_dataset_emit(C.byref(step.cfg), u.data_ptr(), x.data_ptr(), mu.data_ptr(), B, 0, m6, s6,
                                      outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), L.stream()), "emit")
    else:
        L.check(lib.ppnn_dataset_stats(C.byref(step.cfg), u.data_ptr(), x.data_ptr(), mu.data_ptr(), B, 0,
                                       partial.data_ptr(), L.stream()), "stats")
for _ in range(3): run()
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): run()
e.record(); torch.cuda.synchronize()
ms = a.elapsed_time(e) / 20
bpp = 20 if what == "emit" else 8
chk = partial.sum(0).flatten().tolist() if what == "stats" else [float(o.double().sum()) for o in outs]
print(f"dataset {what} {case}: {ms:.4f} ms, {bpp * B * 2 * 257 * 257 / ms / 1e6:.0f} GB/s, check {['%.9e' % c for c in chk]}")
