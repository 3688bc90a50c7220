"""Time the fine-grid solver step (dev tool, GPU box): python tools/run_fine.py [N] [B] [mode]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from ppnn_b200 import FineBurgers
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1025
b = int(sys.argv[2]) if len(sys.argv) > 2 else 64
mode = sys.argv[3] if len(sys.argv) > 3 else "as_written"
dev = torch.device("cuda")
f = bench.synth_fields(b, n, dev, 5)
if len(sys.argv) > 4 and sys.argv[4] == "pos":      # all-positive fields like the reference's initial conditions
    f = f + 0.6
u, v = f[:, 0].contiguous(), f[:, 1].contiguous()
u2, v2 = torch.empty_like(u), torch.empty_like(v)
mu = torch.full((b,), 0.04, device=dev)
fb = FineBurgers(mode=mode, endpoint=(n % 2 == 1))
for _ in range(3):
    fb.step_(u, v, mu, u2, v2)
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
reps = 20
for _ in range(reps // 2):
    fb.step_(u, v, mu, u2, v2)
    fb.step_(u2, v2, mu, u, v)
e.record()
torch.cuda.synchronize()
ms = a.elapsed_time(e) / reps
pts = b * 2 * n * n
print(f"[{os.path.basename(os.environ.get('PPNN_B200_LIB', 'libppnn_b200.so'))}] fine {mode} {'pos' if len(sys.argv) > 4 else 'mixed'} N={n} B={b}: {ms:.4f} ms/step, {pts / ms / 1e6:.1f} Gpt/s, {8 * pts / ms / 1e6:.0f} GB/s")
