#!/usr/bin/env python
"""The whole flow of the reference's ``train2D.py`` on this repo's hot path, end to end on one GPU:

    1. generate fine-grid Burgers trajectories           operators.py:403-492  -> ppnn_b200.datagen.generate
    2. build the normalised training set                 train2D.py:84-157     -> ppnn_b200.PdeDataset
    3. one-step training of the network                  train2D.py:199-223    (PyTorch, unchanged)
    4. autoregressive test rollout                       train2D.py:230-245    -> CoarseStep.prepare_rollout

Nothing leaves the GPU between the stages except the generator's snapshots (its wire format is a host
tensor, like the reference's ``torch.save``).  Sizes default to a few seconds of work; the reference's
are ``--mesh 257 --para 6 --repeat 36 --gen-steps 20001``.

    python examples/train_one_step.py
    python examples/train_one_step.py --mesh 257 --coarse 33 --para 3 --repeat 2 --gen-steps 2001 --epochs 5

Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ppnn_b200 import CoarseStep, PdeDataset  # noqa: E402
from ppnn_b200 import datagen  # noqa: E402
from ppnn_b200.nets import ConvResNet  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", type=int, default=257)
    ap.add_argument("--coarse", type=int, default=33)
    ap.add_argument("--para", type=int, default=3)
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--gen-steps", type=int, default=2001, help="fine solver steps (dt = 1e-4)")
    ap.add_argument("--every", type=int, default=200, help="snapshot interval = one coarse step of dt 0.02")
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--lr", type=float, default=1e-3)
    a = ap.parse_args(argv)
    dev = torch.device("cuda", torch.cuda.current_device())
    t0 = time.perf_counter()

    # 1. trajectories [P, T, 2, F, F] (pinned host tensor, the reference's wire format)
    fdata = datagen.generate(mesh=a.mesh, para=a.para, repeat=a.repeat, steps=a.gen_steps, every=a.every, device=dev)
    P, T1 = fdata.shape[:2]
    T = T1 - 1
    mu_traj = torch.linspace(0.025, 0.065, a.para).repeat(a.repeat)                 # operators.py:411-414
    mus = mu_traj.reshape(-1, 1).repeat(1, T).reshape(-1, 1, 1, 1)                  # train2D.py `mus`: one per sample
    t_gen = time.perf_counter() - t0

    # 2. training set: pdeu = pde_du(u_t), label = (u_{t+1} - u_t) - pdeu, statistics, normalised arrays -- in HBM
    step = CoarseStep(a.mesh, a.coarse, "burgers2Dpu", "burgers2Dpv", dt=a.every * 1e-4, length=3.2)
    t1 = time.perf_counter()
    ds = PdeDataset(fdata, mus, step, device=dev)
    torch.cuda.synchronize()
    t_set = time.perf_counter() - t1

    # 3. one-step training (train2D.py:199-223)
    torch.manual_seed(0)
    model = ConvResNet(uses_pde=True, grid=a.mesh - 1).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=a.lr)
    crit = torch.nn.MSELoss()
    losses = []
    t2 = time.perf_counter()
    for _ in range(a.epochs):
        tot, n = 0.0, 0
        for u0, du, mu, pdeu in ds.batches(a.batch, shuffle=True):
            loss = crit(model(u0, mu, pdeu), du)
            opt.zero_grad()
            loss.backward()
            opt.step()
            tot += loss.item() * u0.shape[0]
            n += u0.shape[0]
        losses.append(tot / n)
    torch.cuda.synchronize()
    t_train = time.perf_counter() - t2

    # 4. test rollout from the first state of every trajectory against the generated data (train2D.py:230-249)
    model.eval()
    init = fdata[:, 0].to(dev)
    with torch.no_grad():
        plan = step.prepare_rollout(model, P, mu_traj.to(dev), ds.stats(), uses_pde=True, graph=True, device=dev)
        t3 = time.perf_counter()
        pred = plan.run(init, T)
        torch.cuda.synchronize()
        t_roll = time.perf_counter() - t3
    label = fdata[:, 1:].to(dev)
    err = ((pred - label).norm() / label.norm()).item()
    print(json.dumps({"trajectories": P, "samples": len(ds), "mesh": a.mesh, "coarse": a.coarse,
                      "seconds": {"generate": round(t_gen, 3), "dataset": round(t_set, 4), "train": round(t_train, 3),
                                  "rollout": round(t_roll, 4)},
                      "loss_per_epoch": [round(x, 5) for x in losses], "rollout_rel_l2": round(err, 5),
                      "finite": bool(torch.isfinite(pred).all())}))
    return losses, err


if __name__ == "__main__":
    main()
