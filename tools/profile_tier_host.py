"""Host-side cost of the Tier-A pde_du chain at rollout batch sizes (dev tool, GPU box): cProfile over 300 calls."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch
from test_gpu_tier_chain import CASES, make_tier_a
import ppnn_oracle as O
name = sys.argv[1] if len(sys.argv) > 1 else "RD_pPDE"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 6
c = CASES[name]
f = make_tier_a(c)
ostep = O.CoarseStep(257, c["cmesh"], c["rhsu"], c["rhsv"], c["dt"], c["length"])
u = torch.rand(B, 2, 257, 257, device="cuda") - 0.45
mu = torch.linspace(c["mu"][0], c["mu"][1], B, device="cuda").reshape(-1, 1, 1, 1)
with torch.no_grad():
    for fn, label in ((lambda: f(u, mu), "tier A"), (lambda: ostep.pde_du(u, mu), "ATen eager")):
        for _ in range(20): fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(300): fn()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"{label} {name} B={B}: host {1e6 * (t1 - t0) / 300:.1f} us per call issued, {1e6 * (t2 - t0) / 300:.1f} us incl. drain")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(300): f(u, mu)
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(22)
