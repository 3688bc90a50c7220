"""Time the adjoint kernel alone (dev tool, GPU box)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from ppnn_b200 import CoarseStep, _lib as L
dev = torch.device("cuda")
case = sys.argv[1] if len(sys.argv) > 1 else "rd"
B = 1920
step = CoarseStep(257, 49, "rdpu", "rdpv", 0.002, 6.4) if case == "rd" else CoarseStep(257, 49, "burgers2DFu", "burgers2DFv", 0.02, 3.2)
u = bench.synth_fields(B, 257, dev, 3)
g = torch.randn_like(u)
mu = torch.full((B,), 0.5 if case == "rd" else 0.04, device=dev)
gu, gmu = torch.empty_like(u), torch.empty(B, device=dev)
lib = L.load()
def run():
    L.check(lib.ppnn_pde_du_bwd(C.byref(step.cfg), u.data_ptr(), mu.data_ptr(), g.data_ptr(), gu.data_ptr(), gmu.data_ptr(), B, 1, L.stream()), "bwd")
for _ in range(3): run()
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): run()
e.record(); torch.cuda.synchronize()
ms = a.elapsed_time(e) / 10
print(f"adjoint {case}: {ms:.4f} ms, {8 * B * 2 * 257 * 257 / ms / 1e6:.0f} GB/s")
