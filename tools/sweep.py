"""BASELINE config 5: stencil roofline sweep (GPU box).  Operators on pre-padded input (Tier A), the fused RHS
kernel and the fused fine-grid Burgers step, fp32, N x N grids x GLOBAL batch B, on 1/2/4/8 GPUs.

    python bench.py --sweep [quick|full] > profiles/r2_sweep_n1.md                               (1 GPU)
    python -m torch.distributed.run --nproc-per-node G ... bench.py --sweep full --gpus G        (G GPUs)

(`bench.py --sweep` is the entry because the CPU-reference column is one of bench.py's CPU legs; run directly,
`python tools/sweep.py [--quick]`, this file prints the GPU columns only.)

Multi-GPU: the batch is sharded, rank r owns B/G samples, no communication; cells with B < G are skipped; the
time of a cell is the max over ranks and GB/s is the whole job's.  Algorithmic bytes: operator
8 B / output point; fused RHS on pre-padded halos 20 B / output point; fused Burgers step 8 B / point-update
(16 B per grid point).  Inputs larger than L2 are streamed; smaller ones are L2-resident and marked (L2).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ppnn_b200.operators as ops  # noqa: E402
import ppnn_b200.rhs as rhs  # noqa: E402
from ppnn_b200 import FineBurgers  # noqa: E402

PEAK = 6539.5
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

OPS = ("d2udx2_2D Central2 acc 6 (x)", "d2udy2_2D Central2 acc 6 (y)", "dudx_2D Upwind1 acc 3 (x)",
       "dudy_2D Upwind1 acc 3 (y)", "d2udx2_2D Central2 acc 2 (x)", "dudy_2D Upwind1 acc 1 (y)",
       "rhs.burgers2Dpu fused (acc 3/6)", "fused Burgers step (u,v)")
# units of work per grid point of one sample, bytes per unit
UNITS = {name: (1, 8) for name in OPS}
UNITS["rhs.burgers2Dpu fused (acc 3/6)"] = (1, 20)
UNITS["fused Burgers step (u,v)"] = (2, 8)


def sweep(dev, world=1, rank=0, quick=False, cpu=None, barrier=None, max_over_ranks=None, out=print):
    """``cpu(opname, N) -> (G units/s, cores)`` or None; ``barrier()`` / ``max_over_ranks(ms)`` for world > 1."""
    barrier = barrier or torch.cuda.synchronize
    max_over_ranks = max_over_ranks or (lambda t: t)

    def timed(fn, reps):
        for _ in range(3):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b)) / reps

    say = out if rank == 0 else (lambda *a, **k: None)
    Ns = (256, 1024, 4096) if quick else (256, 512, 1024, 2048, 4096, 8192)
    Bs = (1, 16, 256) if quick else (1, 4, 16, 64, 256)
    cap = 6e9 if quick else 24e9           # bytes per array per GPU
    cpu_col = " CPU reference G/s (cores) | GPU / CPU |" if cpu else ""
    say(f"| op | N | B | GPUs | ms | GB/s (algorithmic, all GPUs) | % of {world} x {PEAK:.0f} GB/s |{cpu_col} note |")
    say("|---|---|---|---|---|---|---|" + ("---|---|" if cpu else "") + "---|")
    cpu_cache = {}
    for N in Ns:
        for B in Bs:
            if B < world or B % world:
                continue
            Bl = B // world
            if Bl * N * N * 4 > cap:
                continue
            g = torch.Generator(device=dev).manual_seed(1 + rank)
            u = torch.rand(Bl, 1, N, N, device=dev, generator=g) - 0.4
            v = torch.rand(Bl, 1, N, N, device=dev, generator=g) - 0.4
            mu = torch.full((Bl, 1, 1, 1), 0.04, device=dev)
            note = "(L2)" if 2 * Bl * N * N * 4 < 100e6 else ""
            reps = 20 if Bl * N * N < 5e7 else 6
            h = 3
            ux = torch.cat((u[:, :, -h:], u, u[:, :, :h]), dim=2)
            uy = torch.cat((u[:, :, :, -h:], u, u[:, :, :, :h]), dim=3)
            h1 = 1
            ux1 = torch.cat((u[:, :, -h1:], u, u[:, :, :h1]), dim=2)
            uy1 = torch.cat((u[:, :, :, -h1:], u, u[:, :, :, :h1]), dim=3)
            o3 = (ops.dudx_2D("Upwind1", 3), ops.dudy_2D("Upwind1", 3), ops.d2udx2_2D("Central2", 6), ops.d2udy2_2D("Central2", 6))
            fb = FineBurgers(endpoint=False)
            u3, v3 = u[:, 0].contiguous(), v[:, 0].contiguous()
            uo, vo = torch.empty_like(u3), torch.empty_like(v3)
            muv = mu.reshape(-1).contiguous()
            fns = (lambda: ops.d2udx2_2D("Central2", 6)(ux), lambda: ops.d2udy2_2D("Central2", 6)(uy),
                   lambda: ops.dudx_2D("Upwind1", 3)(ux), lambda: ops.dudy_2D("Upwind1", 3)(uy),
                   lambda: ops.d2udx2_2D("Central2", 2)(ux1), lambda: ops.dudy_2D("Upwind1", 1)(uy1),
                   lambda: rhs.burgers2Dpu(u, v, mu, *o3, ux, uy, 0.0125, 0.0125, 0.0125 ** 2, 0.0125 ** 2),
                   lambda: fb.step_(u3, v3, muv, uo, vo))
            for name, fn in zip(OPS, fns):
                ms = timed(fn, reps)
                units, bpu = UNITS[name]
                work = units * B * N * N
                gbs = bpu * work / ms / 1e6
                extra = ""
                if cpu:
                    if (name, N) not in cpu_cache:
                        cpu_cache[(name, N)] = cpu(name, N) if rank == 0 else (None, 0)
                    cv, cores = cpu_cache[(name, N)]
                    extra = f" {cv:.3f} ({cores}) | {work / ms / 1e6 / cv:.0f}x |" if cv else " - | - |"
                say(f"| {name} | {N} | {B} | {world} | {ms:.4f} | {gbs:.0f} | {100 * gbs / (world * PEAK):.1f} |{extra} {note} |",
                    flush=True)
            del u, v, ux, uy, ux1, uy1, u3, v3, uo, vo
            torch.cuda.empty_cache()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    sweep(torch.device("cuda"), quick=ap.parse_args().quick)
