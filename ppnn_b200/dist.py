"""Multi-GPU plumbing for the PDE-preserving path: one process per GPU
(``torchrun``), samples sharded across ranks, NO collective on the data path.

Every sample (trajectory / parameter value / initial condition) is independent
(SURVEY.md 8(e)), so dataset generation, the ``pde_du`` precompute and rollouts
shard by sample.  The only exchanges are off the hot path:

* dataset normalisation statistics (train2D.py:115-151 computes per-channel
  mean / std over the WHOLE dataset): an all-reduce of per-channel count / sum /
  sum of squares (``channel_stats``);
* NN gradients in training: ``torch.nn.parallel.DistributedDataParallel`` over NCCL
  (not here -- the networks stay plain PyTorch).

Works with any backend (``nccl`` on the B200 box, ``gloo`` in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size) of the running job; (0, 1) when not distributed."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def bind_to_gpu_numa(device_index: int) -> Optional[List[int]]:
    """Pin this process to the CPU cores NVML reports as local to GPU ``device_index`` (its NUMA node), so that
    pinned host buffers allocated afterwards are first-touched next to the GPU's PCIe root and the host-side
    copies of the staging pipelines do not cross the socket interconnect.  Returns the core list, or None when
    NVML or the affinity call is unavailable (nothing changes then).  Call before allocating pinned memory."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = device_index
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                phys = int(ids[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cores = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced shard ``[lo, hi)`` of ``n`` samples for ``rank``: the first
    ``n % world_size`` ranks own one extra sample, the union over ranks is ``[0, n)``."""
    if world_size < 1 or not (0 <= rank < world_size) or n < 0:
        raise ValueError(f"bad shard request n={n} rank={rank} world={world_size}")
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sample_generator(seed: int, global_index: int, device="cpu") -> torch.Generator:
    """RNG stream of ONE sample, derived from its GLOBAL index, so that the union of
    the shards is bit-identical to the single-process run whatever the world size."""
    g = torch.Generator(device=device)
    g.manual_seed((int(seed) * 1_000_003 + int(global_index)) % (2 ** 63 - 1))
    return g


def fourier_coefficients(seed: int, lo: int, hi: int, modes: int = 4) -> torch.Tensor:
    """[hi-lo, 2 fields, (2*modes+1)^2, 2] N(0,1) coefficients of the reference's random
    Fourier initial condition (operators.py:436-441), one RNG stream per global sample."""
    m = (2 * modes + 1) ** 2
    out = torch.empty(hi - lo, 2, m, 2)
    for k, idx in enumerate(range(lo, hi)):
        out[k] = torch.randn(2, m, 2, generator=sample_generator(seed, idx))
    return out


def channel_stats(x: torch.Tensor, dims: Sequence[int] = (0, 2, 3), group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-channel mean and (unbiased, like ``Tensor.std``) standard deviation of the
    GLOBAL dataset from local shards ``x`` [b, C, H, W]: one all-reduce of
    (count, sum, sum of squares) in fp64.  Shapes are ``[1, C, 1, 1]`` like
    train2D.py:115-116."""
    xd = x.double()
    cnt = torch.tensor([float(xd.numel() // xd.shape[1])], dtype=torch.float64, device=x.device)
    s1 = xd.sum(dim=tuple(dims))
    s2 = (xd * xd).sum(dim=tuple(dims))
    buf = torch.cat((cnt, s1, s2))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    c = x.shape[1]
    n, s1, s2 = buf[0], buf[1:1 + c], buf[1 + c:]
    mean = s1 / n
    var = (s2 - n * mean * mean) / (n - 1)
    shape = [1, c] + [1] * (x.dim() - 2)
    return mean.float().reshape(shape), var.clamp_min(0).sqrt().float().reshape(shape)


def scalar_stats(x: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Mean and unbiased standard deviation (``Tensor.mean()`` / ``Tensor.std()``, train2D.py:149-150) of
    ALL values of the global data set from the local shard ``x``: one fp64 all-reduce of (count, sum of
    deviations from a pivot, sum of squared deviations).  The pivot (the global maximum, itself
    all-reduced) keeps the variance formula well conditioned when |mean| >> std.  0-d float32 results
    on ``x``'s device."""
    xd = x.detach().double().reshape(-1)
    ws = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    pivot = xd.max() if xd.numel() else torch.tensor(float("-inf"), dtype=torch.float64, device=x.device)
    if ws > 1:
        pivot = pivot.clone()
        dist.all_reduce(pivot, op=dist.ReduceOp.MAX, group=group)
    d = xd - pivot
    buf = torch.stack((torch.tensor(float(xd.numel()), dtype=torch.float64, device=x.device), d.sum(), (d * d).sum()))
    if ws > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    n, s1, s2 = buf[0], buf[1], buf[2]
    mean = pivot + s1 / n
    var = (s2 - s1 * s1 / n) / (n - 1)
    return mean.float(), var.clamp_min(0).sqrt().float()


def gather_shards(local: torch.Tensor, n_total: int, group=None) -> Optional[torch.Tensor]:
    """Concatenate the per-rank shards (possibly ragged) on rank 0 (None elsewhere)."""
    rank, ws = world()
    if ws == 1 or not (dist.is_available() and dist.is_initialized()):
        return local
    sizes = [hi - lo for lo, hi in (shard_bounds(n_total, r, ws) for r in range(ws))]
    nmax = max(sizes)
    # collectives need equal shapes: pad the ragged shards to the largest one, trim after
    padded = local.contiguous()
    if padded.shape[0] < nmax:
        pad = torch.zeros((nmax - padded.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat((padded, pad), dim=0)
    pieces = [torch.empty_like(padded) for _ in range(ws)] if rank == 0 else None
    dist.gather(padded, pieces, dst=0, group=group)
    return torch.cat([p[:n] for p, n in zip(pieces, sizes)], dim=0) if rank == 0 else None


def shard_file(path: str, rank: int, world_size: int) -> str:
    """File name of one rank's shard: ``data.pt`` -> ``data.rank003-of-008.pt``."""
    if world_size == 1:
        return path
    root, ext = os.path.splitext(path)
    return f"{root}.rank{rank:03d}-of-{world_size:03d}{ext}"


def merge_shard_files(path: str, world_size: int, key: str = "u") -> torch.Tensor:
    """Merge per-rank ``{'u': tensor}`` files (reference wire format, train2D.py:86)."""
    parts = [torch.load(shard_file(path, r, world_size), map_location="cpu")[key] for r in range(world_size)]
    return torch.cat(parts, dim=0)
