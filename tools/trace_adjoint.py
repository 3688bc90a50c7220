"""CTA timeline of the periodic adjoint kernel from a -DPPNN_TRACE build of coarse_adjoint.cu (dev tool, GPU box):
PPNN_B200_LIB=build/trace/adj_trace.so python tools/trace_adjoint.py [rd|burgers] [saved|plain]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, bench
from ppnn_b200 import CoarseStep, _lib as L
dev = torch.device("cuda")
case = sys.argv[1] if len(sys.argv) > 1 else "rd"
saved = (sys.argv[2] if len(sys.argv) > 2 else "saved") == "saved"
B = 1920
step = {"rd": lambda: CoarseStep(257, 49, "rdpu", "rdpv", 0.002, 6.4),
        "burgers": lambda: CoarseStep(257, 49, "burgers2DFu", "burgers2DFv", 0.02, 3.2)}[case]()
u = bench.synth_fields(B, 257, dev, 3)
g = torch.randn_like(u)
mu = torch.full((B,), 0.5 if case == "rd" else 0.04, device=dev)
gu, gmu = torch.empty_like(u), torch.empty(B, device=dev)
coarse = torch.empty(B, 2, step.CH - 1, step.CW - 1, device=dev)
lib = L.load()
raw = C.CDLL(os.environ["PPNN_B200_LIB"])
L.check(lib.ppnn_pde_du_fwd_save(C.byref(step.cfg), u.data_ptr(), mu.data_ptr(), gu.data_ptr(), coarse.data_ptr(), B, L.stream()), "fwd")
f = lib.ppnn_pde_du_bwd_saved if saved else lib.ppnn_pde_du_bwd
for _ in range(4):
    L.check(f(C.byref(step.cfg), (coarse if saved else u).data_ptr(), mu.data_ptr(), g.data_ptr(), gu.data_ptr(), gmu.data_ptr(), B, 1, L.stream()), "bwd")
torch.cuda.synchronize()
n = 2 * B
buf = (C.c_ulonglong * (8 * n))()
raw.ppnn_trace_read_adj.argtypes = [C.c_void_p, C.c_long]
assert raw.ppnn_trace_read_adj(buf, 8 * n) == 0
t = np.frombuffer(buf, dtype=np.uint64).reshape(n, 8).astype(np.int64)
t0 = t[:, 0].min()
T = t[:, :5] - t0
sm = t[:, 7]
d = lambda a, b: f"{np.mean(T[:, b] - T[:, a]) / 1e3:.2f}"
print(f"{case} {'saved' if saved else 'plain'}: span {T[:, 4].max() / 1e3:.1f} us; entry->stream end {d(0, 1)} us, fold + W pass + coarse state {d(1, 2)} us, "
      f"grad_mu + exchange {d(2, 3)} us, stencils^T + scatter {d(3, 4)} us; total per CTA {d(0, 4)} us; CTAs per SM {n / len(set(sm.tolist())):.2f}")
gaps = []
for s_ in sorted(set(sm.tolist())):
    m = sm == s_
    st_, en_ = np.sort(T[m, 0]), np.sort(T[m, 4])
    for k in range(2, len(st_)):
        gaps.append(st_[k] - en_[k - 2])
print(f"  CTA end -> next CTA entry on the same SM: mean {np.mean(gaps) / 1e3:.2f} us")
