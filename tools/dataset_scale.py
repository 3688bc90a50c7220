"""PdeDataset at the reference's scale: 216 trajectories x 100 steps (21 600 samples) of 257^2, burgers_pPDE."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppnn_b200 import CoarseStep, PdeDataset
dev = torch.device("cuda")
P, T, F = 216, 100, 257
g = torch.Generator(device=dev).manual_seed(0)
fdata = torch.rand(P, T + 1, 2, F, F, device=dev, generator=g) * 0.5 + 0.1
mus = torch.linspace(0.02, 0.07, 6, device=dev).repeat_interleave(36).reshape(-1, 1).repeat(1, T).reshape(-1)
step = CoarseStep(F, 33, "burgers2Dpu", "burgers2Dpv", 0.02, 3.2)
torch.cuda.synchronize(); t0 = time.perf_counter()
ds = PdeDataset(fdata, mus, step)
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"PdeDataset over {len(ds)} samples: {t1 - t0:.3f} s; mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
print("stats", ds.inmean.flatten().tolist(), ds.instd.flatten().tolist(), ds.pdestd.flatten().tolist(), ds.outstd.flatten().tolist())
assert all(torch.isfinite(t).all() for t in (ds.u0_normd, ds.pdeu_normd, ds.du_normd))
# cross-check the statistics on the device in fp64
u0 = fdata[:, :-1].reshape(-1, 2, F, F)[:, :, :-1, :-1]
m = torch.stack([u0[:, c].double().mean() for c in range(2)]); s = torch.stack([u0[:, c].double().std() for c in range(2)])
print("inmean err", (ds.inmean.flatten().double() - m).abs().max().item(), "instd rel err", ((ds.instd.flatten().double() - s) / s).abs().max().item())
print("u0_normd mean/std", ds.u0_normd[:, 0].double().mean().item(), ds.u0_normd[:, 0].double().std().item())
t0 = time.perf_counter(); n = 0
for batch in ds.batches(800): n += batch[0].shape[0]
torch.cuda.synchronize(); print(f"one epoch of device batches (800): {time.perf_counter() - t0:.3f} s, {n} samples")
