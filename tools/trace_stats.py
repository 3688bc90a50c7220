"""CTA timeline of the statistics pass from a -DPPNN_TRACE build of coarse_dataset.cu (dev tool, GPU box):
PPNN_B200_LIB=build/trace/libppnn_trace.so python tools/trace_stats.py [rd|burgers33]
Prints per-CTA prologue / stream durations and, per SM, how long no resident CTA was streaming."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, bench
from ppnn_b200 import CoarseStep, _lib as L
dev = torch.device("cuda")
case = sys.argv[1] if len(sys.argv) > 1 else "rd"
B = 1920
step = {"rd": lambda: CoarseStep(257, 49, "rdpu", "rdpv", 0.002, 6.4),
        "burgers33": lambda: CoarseStep(257, 33, "burgers2Dpu", "burgers2Dpv", 0.02, 3.2)}[case]()
u = bench.synth_fields(B, 257, dev, 3)
x = u + 1e-3 * torch.randn_like(u)
mu = torch.full((B,), 0.5 if case == "rd" else 0.04, device=dev)
partial = torch.zeros(B, 2, 6, dtype=torch.float64, device=dev)
lib = L.load()
raw = C.CDLL(os.environ["PPNN_B200_LIB"])
for _ in range(4):
    L.check(lib.ppnn_dataset_stats(C.byref(step.cfg), u.data_ptr(), x.data_ptr(), mu.data_ptr(), B, 0,
                                   partial.data_ptr(), L.stream()), "stats")
torch.cuda.synchronize()
n = 2 * B
buf = (C.c_ulonglong * (6 * n))()
raw.ppnn_trace_read.argtypes = [C.c_void_p, C.c_long]
assert raw.ppnn_trace_read(buf, 6 * n) == 0
t = np.frombuffer(buf, dtype=np.uint64).reshape(n, 6).astype(np.int64)
t0 = t[:, 0].min()
start, pro, end, sm = t[:, 0] - t0, t[:, 1] - t0, t[:, 2] - t0, t[:, 3]
print(f"kernel span {(end.max()) / 1e3:.1f} us; prologue mean {np.mean(pro - start) / 1e3:.2f} us (p10 {np.percentile(pro - start, 10) / 1e3:.2f}, "
      f"p90 {np.percentile(pro - start, 90) / 1e3:.2f}); stream mean {np.mean(end - pro) / 1e3:.2f} us (p10 {np.percentile(end - pro, 10) / 1e3:.2f}, "
      f"p90 {np.percentile(end - pro, 90) / 1e3:.2f}); CTAs per SM {n / len(set(sm.tolist())):.2f}")
gat = t[:, 4] - t0
print(f"  entry -> gather done {np.mean(gat - start) / 1e3:.2f} us, gather done -> rhs done {np.mean(pro - gat) / 1e3:.2f} us")
gaps = []
for s_ in sorted(set(sm.tolist())):
    m = sm == s_
    st_, en_ = np.sort(start[m]), np.sort(end[m])
    # a new CTA starts when a slot frees: pair each start after the first two with the end that preceded it
    ns = int(os.environ.get("SLOTS", "2"))
    for k in range(ns, len(st_)):
        gaps.append(st_[k] - en_[k - ns])
print(f"  CTA end -> next CTA entry on the same SM: mean {np.mean(gaps) / 1e3:.2f} us")
idle = []
for s in sorted(set(sm.tolist())):
    m = sm == s
    ev = sorted([(a, 1) for a in pro[m]] + [(b, -1) for b in end[m]])
    lvl, last, none = 0, 0, 0
    for tt, d in ev:
        if lvl == 0: none += tt - last
        lvl += d; last = tt
    idle.append(none / end[m].max())
print(f"share of an SM's time with no CTA streaming: mean {np.mean(idle):.3f}, min {np.min(idle):.3f}, max {np.max(idle):.3f}")
s0 = sorted(set(sm.tolist()))[0]
m = sm == s0
o = np.argsort(start[m])
print("SM", s0, "first 8 CTAs (block, start, prologue end, end) us:")
for k in o[:8]:
    print("  ", int(np.nonzero(m)[0][k]), f"{start[m][k] / 1e3:.2f} {pro[m][k] / 1e3:.2f} {end[m][k] / 1e3:.2f}")
