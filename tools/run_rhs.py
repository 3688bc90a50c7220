"""Time one fused rhs.py call on a large true-periodic grid (dev tool, GPU box; BASELINE config 5's "fused Burgers RHS"):
python tools/run_rhs.py [name] [N] [B] [adv_acc] [diff_acc]"""
import os, sys, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch
import ppnn_b200.operators as ops, ppnn_b200.rhs as rhs
name = sys.argv[1] if len(sys.argv) > 1 else "burgers2Dpu"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
aa = int(sys.argv[4]) if len(sys.argv) > 4 else 1
da = int(sys.argv[5]) if len(sys.argv) > 5 else 2
h = max(aa, da // 2)
assert aa == h and da // 2 == h, "one common halo"
torch.manual_seed(5)
u1 = torch.randn(B, 1, N, N, device="cuda"); v1 = torch.randn(B, 1, N, N, device="cuda")
mu = torch.rand(B, 1, 1, 1, device="cuda") * 0.05 + 0.02
ax = torch.cat((u1[:, :, -h:], u1, u1[:, :, :h]), 2); ay = torch.cat((u1[:, :, :, -h:], u1, u1[:, :, :, :h]), 3)
dudx, dudy, d2x, d2y = ops.dudx_2D(accuracy=aa), ops.dudy_2D(accuracy=aa), ops.d2udx2_2D(accuracy=da), ops.d2udy2_2D(accuracy=da)
f = getattr(rhs, name)
run = lambda: f(u1, v1, mu, dudx, dudy, d2x, d2y, ax, ay, 0.1, 0.08, 0.01, 0.0064)
for _ in range(3): y = run()
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): y = run()
e.record(); torch.cuda.synchronize()
ms = a.elapsed_time(e) / 10
print(f"{name} acc {aa}/{da} {B} x {N}^2: {ms:.4f} ms, {20 * y.numel() / ms / 1e6:.0f} GB/s (20 B/pt), sha {hashlib.sha1(y.cpu().numpy().tobytes()).hexdigest()[:12]}")
