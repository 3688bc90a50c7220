"""Time the adjoint kernel alone (dev tool, GPU box):  python tools/run_adjoint.py [rd|burgers|burgers33] [saved|plain] [gnn]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from ppnn_b200 import CoarseStep, _lib as L
dev = torch.device("cuda")
case = sys.argv[1] if len(sys.argv) > 1 else "rd"
saved = (sys.argv[2] if len(sys.argv) > 2 else "plain") == "saved"
gnn_on = len(sys.argv) > 3 and sys.argv[3] == "gnn"
B = 1920
step = {"rd": lambda: CoarseStep(257, 49, "rdpu", "rdpv", 0.002, 6.4),
        "burgers": lambda: CoarseStep(257, 49, "burgers2DFu", "burgers2DFv", 0.02, 3.2),
        "burgers33": lambda: CoarseStep(257, 33, "burgers2Dpu", "burgers2Dpv", 0.02, 3.2)}[case]()
u = bench.synth_fields(B, 257, dev, 3)
g = torch.randn_like(u)
mu = torch.full((B,), 0.5 if case == "rd" else 0.04, device=dev)
gu, gmu = torch.empty_like(u), torch.empty(B, device=dev)
gnn = torch.empty(B, 2, 256, 256, device=dev) if gnn_on else None
coarse = torch.empty(B, 2, step.CH - 1, step.CW - 1, device=dev)
lib = L.load()
os_ = L.f2([3e-3, 2e-3])
L.check(lib.ppnn_pde_du_fwd_save(C.byref(step.cfg), u.data_ptr(), mu.data_ptr(), gu.data_ptr(), coarse.data_ptr(), B, L.stream()), "fwd")
def run():
    if gnn_on:
        f = lib.ppnn_step_update_bwd_saved if saved else lib.ppnn_step_update_bwd
        L.check(f(C.byref(step.cfg), (coarse if saved else u).data_ptr(), mu.data_ptr(), g.data_ptr(), os_, gu.data_ptr(),
                  gmu.data_ptr(), gnn.data_ptr(), B, L.stream()), "bwd")
    else:
        f = lib.ppnn_pde_du_bwd_saved if saved else lib.ppnn_pde_du_bwd
        L.check(f(C.byref(step.cfg), (coarse if saved else u).data_ptr(), mu.data_ptr(), g.data_ptr(), gu.data_ptr(),
                  gmu.data_ptr(), B, 1, L.stream()), "bwd")
for _ in range(3): run()
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): run()
e.record(); torch.cuda.synchronize()
ms = a.elapsed_time(e) / 10
bpp = 12 if gnn_on else 8
print(f"adjoint {case} {'saved' if saved else 'plain'}{' +gnn' if gnn_on else ''}: {ms:.4f} ms, {bpp * B * 2 * 257 * 257 / ms / 1e6:.0f} GB/s")
