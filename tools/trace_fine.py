"""CTA timeline of the endpoint-mesh fine solver kernel from a -DPPNN_TRACE build of fine_step.cu (dev tool, GPU box):
PPNN_B200_LIB=build/trace/fine_trace.so python tools/trace_fine.py [N] [B]   -- how much of a step of an L2-resident
ensemble is pipeline fill / drain, and what a step costs inside a CUDA graph."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, bench
from ppnn_b200 import FineBurgers
n = int(sys.argv[1]) if len(sys.argv) > 1 else 257
b = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda")
f = bench.synth_fields(b, n, dev, 5) + 0.6
u, v = f[:, 0].contiguous(), f[:, 1].contiguous()
u2, v2 = torch.empty_like(u), torch.empty_like(v)
mu = torch.full((b,), 0.04, device=dev)
fb = FineBurgers(mode="as_written", endpoint=True)
for _ in range(5):
    fb.step_(u, v, mu, u2, v2); fb.step_(u2, v2, mu, u, v)
torch.cuda.synchronize()
raw = C.CDLL(os.environ["PPNN_B200_LIB"])
buf = (C.c_ulonglong * 4096)()
raw.ppnn_trace_read_fine.argtypes = [C.c_void_p, C.c_long]
assert raw.ppnn_trace_read_fine(buf, 4096) == 0
t = np.frombuffer(buf, dtype=np.uint64).reshape(1024, 4).astype(np.int64)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
print(f"{b} x {n}^2, {len(t)} CTAs: entry spread {(t[:, 0].max() - t0) / 1e3:.2f} us; entry -> first bands ready {np.mean(t[:, 1] - t[:, 0]) / 1e3:.2f} us; "
      f"first bands ready -> last store {np.mean(t[:, 2] - t[:, 1]) / 1e3:.2f} us; kernel span {(t[:, 2].max() - t0) / 1e3:.2f} us")
# the same step inside a CUDA graph (what the generator replays)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    fb.step_(u, v, mu, u2, v2); fb.step_(u2, v2, mu, u, v)
    s.synchronize()
    with torch.cuda.graph(g, stream=s):
        for _ in range(100):
            fb.step_(u, v, mu, u2, v2); fb.step_(u2, v2, mu, u, v)
g.replay(); torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); g.replay(); g.replay(); e.record(); torch.cuda.synchronize()
print(f"graph replay: {a.elapsed_time(e) * 1e3 / 400:.2f} us per step")
