"""Where the wall-clock time of a dataset-generation run goes (dev tool, GPU box):
python tools/datagen_profile.py [mesh] [samples] [steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
mesh = int(sys.argv[1]) if len(sys.argv) > 1 else 1025
b = int(sys.argv[2]) if len(sys.argv) > 2 else 64
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1001
every = 200
t = [time.perf_counter()]
def lap(name):
    torch.cuda.synchronize()
    t.append(time.perf_counter())
    print(f"{name:28s} {t[-1] - t[-2]:.3f} s", flush=True)
torch.cuda.set_device(0)
torch.zeros(1, device="cuda")
lap("cuda context")
from ppnn_b200 import datagen as G, dist as D, FineBurgers, _lib as L
L.load()
lap("import + library load")
coef = D.fourier_coefficients(10, 0, b)
lap("coefficients (host RNG)")
dev = torch.device("cuda", 0)
u0 = G.fourier_initial_condition(coef[:, 0], mesh, dev)
v0 = G.fourier_initial_condition(coef[:, 1], mesh, dev)
lap("initial conditions")
T = len(range(0, steps, every))
w = G.SnapshotWriter(b, T, mesh, mesh, dev)
lap(f"writer buffers (result {w.host.numel() * 4 / 1e9:.2f} GB)")
mu = torch.linspace(0.025, 0.065, 8).repeat(b // 8 + 1)[:b].to(dev)
s = FineBurgers()
s.run(u0, v0, mu, steps, every=every, graph=True, sink=w)
lap(f"solver.run {steps} steps")
w.finish()
lap("writer.finish")
w2 = G.SnapshotWriter(b, T, mesh, mesh, dev)
s.run(u0, v0, mu, steps, every=every, graph=True, sink=w2)
w2.finish()
lap("second run incl. writer set-up")
