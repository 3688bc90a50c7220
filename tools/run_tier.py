"""Run the Tier-A pde_du chain (train2D.py:38-79 around the drop-in modules) a few times (dev tool, GPU box; for
ncu launch lists):  python tools/run_tier.py [burgers_pPDE|RD_pPDE|burgers_fPDE] [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch
from test_gpu_tier_chain import CASES, make_tier_a, timed
name = sys.argv[1] if len(sys.argv) > 1 else "burgers_pPDE"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2160
c = CASES[name]
f = make_tier_a(c)
u = torch.rand(B, 2, 257, 257, device="cuda") - 0.45
mu = torch.linspace(c["mu"][0], c["mu"][1], B, device="cuda").reshape(-1, 1, 1, 1)
with torch.no_grad():
    print(f"tier A {name} B={B}: {timed(lambda: f(u, mu), reps=5):.4f} ms")
