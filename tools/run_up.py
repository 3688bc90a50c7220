"""Time mesh_convertor.up alone (dev tool, GPU box):  python tools/run_up.py [C] [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ppnn_b200.utility.utils import mesh_convertor
C = int(sys.argv[1]) if len(sys.argv) > 1 else 33
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2160
m = mesh_convertor(257, C, dim=2)
x = torch.randn(B, 1, C, C, device="cuda")
for _ in range(3): y = m.up(x)
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): y = m.up(x)
e.record(); torch.cuda.synchronize()
ms = a.elapsed_time(e) / 10
print(f"up {C}->257 B={B}: {ms:.4f} ms, {y.numel() * 4 / ms / 1e6:.0f} GB/s written, checksum {float(y.double().sum()):.6e}")
