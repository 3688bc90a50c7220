"""BASELINE config 5: stencil roofline sweep (GPU box).  Operators on pre-padded input
(Tier A), the fused RHS kernel and the fused fine-grid Burgers step, fp32, N x N grids x batch B.

    python tools/sweep.py [--quick] > profiles/r1_sweep.md

Algorithmic bytes: operator 8 B / output point; fused RHS on pre-padded halos 20 B /
output point; fused Burgers step 8 B / point-update (16 B per grid point).  Inputs larger than L2 are
streamed; smaller ones are L2-resident and marked (L2).
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


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda")
    Ns = (256, 1024, 4096) if a.quick else (256, 512, 1024, 2048, 4096, 8192)
    Bs = (1, 16, 256) if a.quick else (1, 4, 16, 64, 256)
    cap = 6e9 if a.quick else 24e9           # bytes per array
    print(f"| op | N | B | ms | GB/s (algorithmic) | % of {PEAK:.0f} GB/s | note |")
    print("|---|---|---|---|---|---|---|")
    for N in Ns:
        for B in Bs:
            if B * N * N * 4 > cap:
                continue
            g = torch.Generator(device=dev).manual_seed(1)
            u = torch.rand(B, 1, N, N, device=dev, generator=g) - 0.4
            v = torch.rand(B, 1, N, N, device=dev, generator=g) - 0.4
            mu = torch.full((B, 1, 1, 1), 0.04, device=dev)
            note = "(L2)" if 2 * B * N * N * 4 < 100e6 else ""
            reps = 20 if B * N * N < 5e7 else 6
            h = 3
            ux = torch.cat((u[:, :, -h:], u, u[:, :, :h]), dim=2)
            uy = torch.cat((u[:, :, :, -h:], u, u[:, :, :, :h]), dim=3)
            pts = B * N * N
            cases = [
                ("d2udx2_2D Central2 acc 6 (x)", lambda: ops.d2udx2_2D("Central2", 6)(ux), 8),
                ("d2udy2_2D Central2 acc 6 (y)", lambda: ops.d2udy2_2D("Central2", 6)(uy), 8),
                ("dudx_2D Upwind1 acc 3 (x)", lambda: ops.dudx_2D("Upwind1", 3)(ux), 8),
                ("dudy_2D Upwind1 acc 3 (y)", lambda: ops.dudy_2D("Upwind1", 3)(uy), 8),
            ]
            h1 = 1
            ux1 = torch.cat((u[:, :, -h1:], u, u[:, :, :h1]), dim=2)
            uy1 = torch.cat((u[:, :, :, -h1:], u, u[:, :, :, :h1]), dim=3)
            cases += [("d2udx2_2D Central2 acc 2 (x)", lambda: ops.d2udx2_2D("Central2", 2)(ux1), 8),
                      ("dudy_2D Upwind1 acc 1 (y)", lambda: ops.dudy_2D("Upwind1", 1)(uy1), 8)]
            o3 = (ops.dudx_2D("Upwind1", 3), ops.dudy_2D("Upwind1", 3), ops.d2udx2_2D("Central2", 6), ops.d2udy2_2D("Central2", 6))
            cases.append(("rhs.burgers2Dpu fused (acc 3/6)",
                          lambda: rhs.burgers2Dpu(u, v, mu, *o3, ux, uy, 0.0125, 0.0125, 0.0125 ** 2, 0.0125 ** 2), 20))
            fb = FineBurgers(endpoint=False)
            u3, v3 = u[:, 0].contiguous(), v[:, 0].contiguous()
            uo, vo = torch.empty_like(u3), torch.empty_like(v3)
            muv = mu.reshape(-1).contiguous()
            cases.append(("fused Burgers step (u,v)", lambda: fb.step_(u3, v3, muv, uo, vo), 16))
            for name, fn, bpp in cases:
                ms = timed(fn, reps)
                gbs = bpp * pts / ms / 1e6
                print(f"| {name} | {N} | {B} | {ms:.4f} | {gbs:.0f} | {100 * gbs / PEAK:.1f} | {note} |", flush=True)
            del u, v, ux, uy, ux1, uy1, u3, v3, uo, vo
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
