"""Time the on-chip solvers (dev tool, GPU box): python tools/run_ks.py"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ppnn_b200 import KS1D, KS2D
dev = torch.device("cuda")
tag = os.path.basename(os.environ.get("PPNN_B200_LIB", "libppnn_b200.so"))

def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

x2 = torch.linspace(0, 8 * math.pi, 65, device=dev)
w = (torch.cos(x2.reshape(-1, 1) / 4) * (1 + torch.sin(x2.reshape(1, -1) / 4))).reshape(1, 1, 65, 65)
ks2 = KS2D(dt=1e-3, dx=math.pi / 8)
for b in (1, 148, 296, 1024):
    wb = (w * torch.linspace(0.9, 1.1, b, device=dev).reshape(-1, 1, 1, 1)).contiguous()
    n = 1000
    print(f"[{tag}] ks2d {b}x65x65: {t(lambda: ks2.run(wb, n)) / n * 1e3:.2f} us/step")
x = torch.linspace(0, 8 * math.pi, 129, device=dev).reshape(1, 1, -1)
u = torch.cos(x / 16 + torch.randn(256, 1, 1, device=dev)) * (1 + torch.sin(x / 16 + torch.randn(256, 1, 1, device=dev)))
ks1 = KS1D(dt=1e-4, dx=math.pi / 8)
print(f"[{tag}] ks1d 256x129: {t(lambda: ks1.run(u, 20000)) / 20000 * 1e3:.3f} us/step")
