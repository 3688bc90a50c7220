"""Device-resident training set: the ``myset`` of the reference (train2D.py:84-157).

The reference builds its training arrays on the host: ``pde_du`` of all 21 600
samples goes ``.to(device)`` ... ``.cpu()`` (2 x 5.7 GB over PCIe), and the three
normalisations allocate ~10 full-size temporaries.  A B200 holds the whole data
set (11.5 GB of trajectories + 3 x 11.3 GB of normalised arrays) in its 180 GB of
HBM, so here the trajectories are uploaded once (pinned, chunked, on a side stream)
and everything else is two streaming kernel passes (``ppnn_dataset_stats`` /
``ppnn_dataset_emit``, csrc/coarse_dataset.cu); batches are sliced on the device.

Attribute names, shapes and the ``__getitem__`` tuple follow the reference class so
that the training loop of train2D.py:199-223 reads the same.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterator, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib as L
from .dist import scalar_stats
from .fused import CoarseStep

_NAMES = ("in", "pde", "out")


def upload(host: torch.Tensor, device, chunk_bytes: int = 256 << 20) -> torch.Tensor:
    """Host -> HBM through two pinned staging buffers on a side stream (the host
    tensor need not be pinned; a pinned one is copied directly)."""
    device = torch.device(device)
    host = host.detach()
    if host.is_cuda:
        return host.to(device=device, dtype=torch.float32).contiguous()
    host = host.to(torch.float32).contiguous()
    out = torch.empty(host.shape, dtype=torch.float32, device=device)
    if host.is_pinned():
        out.copy_(host, non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
        return out
    flat_h, flat_d = host.reshape(-1), out.reshape(-1)
    n = flat_h.numel()
    step = max(1, min(n, chunk_bytes // 4))
    stage = [torch.empty(step, dtype=torch.float32).pin_memory() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]
    side = torch.cuda.Stream(device)
    # `out` was allocated on the current stream: a recycled block may still be read by kernels queued
    # there, so the side stream must not write it before they finish
    side.wait_stream(torch.cuda.current_stream(device))
    with torch.cuda.stream(side):
        for k, lo in enumerate(range(0, n, step)):
            hi = min(n, lo + step)
            s = k & 1
            if k >= 2:
                done[s].synchronize()                     # staging buffer s is free again
            stage[s][:hi - lo].copy_(flat_h[lo:hi])       # pageable -> pinned (host memcpy)
            flat_d[lo:hi].copy_(stage[s][:hi - lo], non_blocking=True)
            done[s].record(side)
    side.synchronize()
    return out


def finalize_stats(sums: torch.Tensor, counts: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(sum, sum of squares) per statistic and channel -> mean and unbiased std.

    ``sums`` [2 channels, 6] fp64 as ``ppnn_dataset_stats`` lays them out, ``counts``
    [3] fp64 = number of values behind (u, pdeu, label).  With a process group the
    shards of all ranks are combined by ONE all-reduce (the data set shards by
    trajectory).  Returns mean, std as [3, 2] float32."""
    buf = torch.cat((sums.reshape(-1).double(), counts.reshape(-1).double()))
    if group is not None or (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    s = buf[:12].reshape(2, 3, 2)                          # [channel, statistic, (sum, sumsq)]
    n = buf[12:15].reshape(1, 3)
    mean = s[:, :, 0] / n
    var = (s[:, :, 1] - s[:, :, 0] * s[:, :, 0] / n) / (n - 1)   # Tensor.std is unbiased
    return mean.t().contiguous().float(), var.clamp_min(0).sqrt().t().contiguous().float()


class PdeDataset(torch.utils.data.Dataset):
    """``myset`` (train2D.py:109-157) with every array resident in HBM.

    fdata  [P, T+1, 2, F, F] trajectories (host or device), already cut to the time
           window of train2D.py:84-85
    mus    one parameter per SAMPLE, [P*T] or [P*T,1,1,1] (train2D.py ``mus``)
    step   the case's CoarseStep; None = ``params.pde == False`` (no pdeu arrays)

    Construction is COLLECTIVE when a process group is initialised (world size > 1): every rank of
    ``group`` (default: the world) must build its shard, or the all-reduces of the statistics hang.
    """

    def __init__(self, fdata: torch.Tensor, mus: torch.Tensor, step: Optional[CoarseStep],
                 device="cuda", group=None, keep_trajectories: bool = False):
        if fdata.dim() != 5 or fdata.shape[2] != 2 or fdata.shape[1] < 2:
            raise RuntimeError(f"fdata has shape {tuple(fdata.shape)}, expected [P,T+1,2,F,F]")
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("PdeDataset lives in GPU memory; there is no CPU fallback")
        P, T1, _, FH, FW = fdata.shape
        T = T1 - 1
        B = P * T
        traj = upload(fdata, device)
        mu = mus.detach().to(device=device, dtype=torch.float32).reshape(-1).contiguous()
        if mu.numel() != B:
            raise RuntimeError(f"mus has {mu.numel()} entries, expected P*T = {B}")
        self.pde = step is not None
        PH, PW = FH - 1, FW - 1
        if self.pde:
            if (step.FH, step.FW) != (FH, FW):
                raise RuntimeError(f"CoarseStep is for a {step.FH}x{step.FW} mesh, data is {FH}x{FW}")
            lib = L.load()
            nxt = traj.reshape(-1)[2 * FH * FW:]               # the same trajectories one state later
            partial = torch.empty(B, 2, 6, dtype=torch.float64, device=device)
            with L.on_device(traj):
                L.check(lib.ppnn_dataset_stats(C.byref(step.cfg), traj.data_ptr(), nxt.data_ptr(), mu.data_ptr(),
                                               B, T, partial.data_ptr(), L.stream()), "ppnn_dataset_stats")
            counts = torch.tensor([B * PH * PW, B * FH * FW, B * PH * PW], dtype=torch.float64, device=device)
            mean, std = finalize_stats(partial.sum(dim=0), counts, group)
            self.u0_normd = torch.empty(B, 2, PH, PW, device=device)
            self.pdeu_normd = torch.empty_like(self.u0_normd)
            self.du_normd = torch.empty_like(self.u0_normd)
            m6 = (C.c_float * 6)(*mean.reshape(-1).tolist())
            s6 = (C.c_float * 6)(*std.reshape(-1).tolist())
            with L.on_device(traj):
                L.check(lib.ppnn_dataset_emit(C.byref(step.cfg), traj.data_ptr(), nxt.data_ptr(), mu.data_ptr(),
                                              B, T, m6, s6, self.u0_normd.data_ptr(), self.pdeu_normd.data_ptr(),
                                              self.du_normd.data_ptr(), L.stream()), "ppnn_dataset_emit")
        else:
            # params.pde == False (train2D.py:134-139): plain normalisation, not the PDE path;
            # element-wise ATen on the device, one trajectory at a time to bound temporaries
            sums = torch.zeros(2, 6, dtype=torch.float64, device=device)
            for p in range(P):
                u0 = traj[p, :-1, :, :-1, :-1].double()
                du = (traj[p, 1:] - traj[p, :-1])[:, :, :-1, :-1].double()
                sums[:, 0] += u0.sum(dim=(0, 2, 3)); sums[:, 1] += (u0 * u0).sum(dim=(0, 2, 3))
                sums[:, 4] += du.sum(dim=(0, 2, 3)); sums[:, 5] += (du * du).sum(dim=(0, 2, 3))
            counts = torch.tensor([B * PH * PW, 2, B * PH * PW], dtype=torch.float64, device=device)
            mean, std = finalize_stats(sums, counts, group)
            shape = (1, 2, 1, 1)
            self.u0_normd = torch.empty(B, 2, PH, PW, device=device)
            self.du_normd = torch.empty_like(self.u0_normd)
            for p in range(P):
                sl = slice(p * T, (p + 1) * T)
                self.u0_normd[sl] = (traj[p, :-1, :, :-1, :-1] - mean[0].reshape(shape)) / std[0].reshape(shape)
                self.du_normd[sl] = ((traj[p, 1:] - traj[p, :-1])[:, :, :-1, :-1] - mean[2].reshape(shape)) / std[2].reshape(shape)
        shape = (1, 2, 1, 1)
        self.inmean, self.instd = mean[0].reshape(shape).to(device), std[0].reshape(shape).to(device)
        if self.pde:
            self.pdemean, self.pdestd = mean[1].reshape(shape).to(device), std[1].reshape(shape).to(device)
        self.outmean, self.outstd = mean[2].reshape(shape).to(device), std[2].reshape(shape).to(device)
        self.mu = mu.reshape(-1, 1, 1, 1)
        # train2D.py:149-150 takes mean / std of `mus` over the WHOLE data set: with sharded trajectories
        # the (count, sum, sum of squares) of the local parameters are all-reduced like the other statistics
        self.mumean, self.mustd = scalar_stats(self.mu, group)
        self.mu_normd = (self.mu - self.mumean) / self.mustd                    # train2D.py:151
        self.trajectories = traj if keep_trajectories else None

    # ---- the reference's Dataset protocol (train2D.py:153-160) ----
    def __getitem__(self, index):
        if self.pde:
            return self.u0_normd[index], self.du_normd[index], self.mu_normd[index], self.pdeu_normd[index]
        return self.u0_normd[index], self.du_normd[index], self.mu_normd[index]

    def __len__(self) -> int:
        return self.u0_normd.shape[0]

    def stats(self) -> Dict[str, torch.Tensor]:
        """The statistics train2D.py:163-179 saves next to the checkpoints."""
        out = {"inmean": self.inmean, "instd": self.instd, "outmean": self.outmean, "outstd": self.outstd,
               "mumean": self.mumean, "mustd": self.mustd}
        if self.pde:
            out.update(pdemean=self.pdemean, pdestd=self.pdestd)
        return out

    def batches(self, batch_size: int, shuffle: bool = True, generator: Optional[torch.Generator] = None,
                drop_last: bool = False) -> Iterator[Tuple[torch.Tensor, ...]]:
        """What ``DataLoader(dataset, batch_size, shuffle=True)`` yields (train2D.py:182),
        gathered on the device: no workers, no pinned staging, no H2D copy per batch."""
        n = len(self)
        dev = self.u0_normd.device
        order = torch.randperm(n, device=dev, generator=generator) if shuffle else torch.arange(n, device=dev)
        for lo in range(0, n, batch_size):
            idx = order[lo:lo + batch_size]
            if drop_last and idx.numel() < batch_size:
                return
            yield self[idx]
