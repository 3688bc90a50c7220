#!/usr/bin/env python
"""Training THROUGH the PDE-preserving step (BASELINE config 3, cases/burgers_fPDE.yaml):
a k-step unrolled PPNN rollout whose every step is the fused kernel
``u <- padBC_rd(net(norm(u), norm(mu))*outstd + outmean) + u + pde_du(u, mu)`` and whose
backward runs the hand-written adjoint kernel, so the loss gradient reaches the network
parameters of every step.  One process per GPU; with WORLD_SIZE > 1 the network is wrapped
in DistributedDataParallel (NCCL all-reduce of the 6.2 MB of fp32 gradients over NVLink) --
the PDE path itself needs no collective, each rank trains on its own shard of trajectories.

    python examples/train_unrolled.py --steps 10 --unroll 4 --batch 16
    torchrun --nproc-per-node 8 examples/train_unrolled.py --steps 50 --unroll 4 --batch 100

Synthetic data (random periodic fields); the network has the layer shapes of the reference's
``cnn2d`` (models.py:77-96).  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ppnn_b200 import CoarseStep  # noqa: E402
from ppnn_b200.nets import ConvResNet  # noqa: E402
import bench  # noqa: E402  (synthetic field generator)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--unroll", type=int, default=4)
    ap.add_argument("--batch", type=int, default=16, help="trajectories per GPU per step")
    ap.add_argument("--lr", type=float, default=1e-3)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    torch.manual_seed(10)                                   # same initial weights on every rank
    net = ConvResNet(uses_pde=False).to(dev)
    model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local]) if world > 1 else net
    opt = torch.optim.Adam(model.parameters(), lr=a.lr)
    # cases/burgers_fPDE.yaml: F = 257, C = 49, length 3.2, dt 0.02, rhs burgers2DF{u,v}, nu in [0.02, 0.07]
    step = CoarseStep(257, 49, "burgers2DFu", "burgers2DFv", 0.02, 3.2)
    B = a.batch
    u0 = bench.synth_fields(B, 257, dev, 100 + rank)
    mu = torch.linspace(0.02, 0.07, 6, device=dev)[torch.arange(B, device=dev) % 6].reshape(B, 1, 1, 1)
    target = bench.synth_fields(B, 257, dev, 200 + rank)    # stand-in label trajectory end state
    inmean, instd = torch.zeros(1, 2, 1, 1, device=dev), torch.full((1, 2, 1, 1), 0.3, device=dev)
    outmean, outstd = [0.0, 0.0], [1e-2, 1e-2]
    mun = (mu - 0.045) / 0.017
    losses, t_fwd_bwd = [], []
    for it in range(a.steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        u = u0
        for _ in range(a.unroll):
            nn_out = model((u[:, :, :-1, :-1] - inmean) / instd, mun)
            u = step.step_update(u, mu, nn_out, outmean, outstd)      # fused forward; adjoint kernel on backward
        loss = torch.nn.functional.mse_loss(u, target)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        torch.cuda.synchronize()
        t_fwd_bwd.append(time.perf_counter() - t0)
        losses.append(loss.item())
    gnorm = torch.sqrt(sum((p.grad.float() ** 2).sum() for p in net.parameters() if p.grad is not None)).item()
    if rank == 0:
        steady = sorted(t_fwd_bwd[1:] or t_fwd_bwd)
        ms = 1e3 * steady[len(steady) // 2]
        print(json.dumps({"example": "train_unrolled", "case": "burgers_fPDE", "n_gpus": world, "unroll": a.unroll,
                          "batch_per_gpu": B, "ms_per_train_step": ms,
                          "traj_steps_per_s": world * B * a.unroll / (ms * 1e-3),
                          "loss_first": losses[0], "loss_last": losses[-1], "grad_norm": gnorm,
                          "params": sum(p.numel() for p in net.parameters())}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
