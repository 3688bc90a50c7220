"""Dataset generation: the fine-grid explicit Burgers solver of the reference's
``src/operators.py`` ``__main__`` block (lines 403-492), sharded by sample across
GPUs with no communication.

    torchrun --nproc-per-node 8 -m ppnn_b200.datagen --out burgers.pt --mesh 1025 \
        --para 8 --repeat 64 --steps 20001 --every 200

Per rank: samples ``shard_bounds(P, rank, world)``; initial conditions from per-sample
RNG streams keyed by the GLOBAL sample index (the union over ranks equals the
1-GPU run bit for bit); one fused kernel per time step replayed from a CUDA graph
of ``every`` steps; snapshots are staged on the device and copied to pinned host
memory on a side stream while the solver keeps running.  Output is the reference's
wire format ``{'u': float32 [P, T, 2, F, F]}`` (what train2D.py:86 loads), one file
per rank (``dist.shard_file``) unless ``--gather``.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import queue
import threading
import time
from typing import Optional

import torch

from . import dist as D
from .fused import FineBurgers


def fourier_initial_condition(coef: torch.Tensor, n: int, device) -> torch.Tensor:
    """Random 81-mode periodic field, min-max normalised to [0.1, 1.1]
    (operators.py:423-445) -- one evaluation kernel + one normalisation kernel
    (``ppnn_fourier_ic``) instead of the reference's 324 full-size element-wise launches.
    ``coef`` [B, (2m+1)^2, 2] holds the (sin, cos) N(0,1) coefficients in the reference's
    (k, l) loop order; x = linspace(0, 2*pi, n) includes the duplicated endpoint.
    Returns [B, 1, n, n] on ``device``."""
    import ctypes as C
    from . import _lib as L
    B, m2, _ = coef.shape
    modes = (int(round(math.sqrt(m2))) - 1) // 2
    device = torch.device(device)
    c = coef.to(device=device, dtype=torch.float32).contiguous()
    out = torch.empty(B, 1, n, n, dtype=torch.float32, device=device)
    scratch = torch.empty(B, 2, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        L.check(L.load().ppnn_fourier_ic(L.dev(c, "coefficients").data_ptr(), out.data_ptr(), scratch.data_ptr(),
                                         B, n, modes, C.c_float(0.1), L.stream()), "ppnn_fourier_ic")
    return out


class SnapshotWriter:
    """Side-stream D2H of (u, v) snapshots into one host tensor ``[B, T, 2, H, W]`` -- hides the only D2H
    of the generator (operators.py:479-482).

    Pinning the whole result would cost more than the solve (page pinning runs at ~0.6 s/GB: 3 s for the
    reference's 5.3 GB data set, 30 s for a 54 GB shard), so only two snapshot-sized staging buffers are
    pinned: a snapshot goes device -> device stage -> pinned buffer on the side stream, and a drain thread
    moves it into the (pageable, first touched there) result while the solver keeps running."""

    NBUF = 2

    def __init__(self, B: int, T: int, H: int, W: int, device, keep: bool = True):
        self.keep = keep                                   # False: snapshots are dropped once they have landed (throughput runs)
        self.host = torch.empty((B, T if keep else 0, 2, H, W), dtype=torch.float32)
        self.stage = torch.empty((B, 2, H, W), dtype=torch.float32, device=device)
        self.pin = [torch.empty((B, 2, H, W), dtype=torch.float32, pin_memory=True) for _ in range(self.NBUF)]
        self.free = [threading.Event() for _ in range(self.NBUF)]
        for e in self.free:
            e.set()
        self.side = torch.cuda.Stream(device=device)
        self.done = torch.cuda.Event()
        self.done.record(self.side)
        self.count = 0
        self.error: Optional[BaseException] = None
        self.jobs: "queue.Queue" = queue.Queue()
        self.thread = threading.Thread(target=self._drain, daemon=True)
        self.thread.start()

    def _drain(self) -> None:
        try:
            while True:
                job = self.jobs.get()
                if job is None:
                    return
                k, c, landed = job
                landed.synchronize()                       # the D2H into pin[k] has completed
                if self.keep:
                    self.host[:, c].copy_(self.pin[k])
                self.free[k].set()
        except BaseException as exc:                       # surfaced by finish()
            self.error = exc
            for e in self.free:
                e.set()

    def __call__(self, i: int, u: torch.Tensor, v: torch.Tensor) -> None:
        cur = torch.cuda.current_stream()
        cur.wait_event(self.done)                          # previous D2H has drained the staging buffer
        self.stage[:, 0].copy_(u.reshape(self.stage[:, 0].shape))
        self.stage[:, 1].copy_(v.reshape(self.stage[:, 1].shape))
        self.side.wait_stream(cur)
        k = self.count % self.NBUF
        self.free[k].wait()                                # the drain thread is done with pin[k]
        self.free[k].clear()
        with torch.cuda.stream(self.side):
            self.pin[k].copy_(self.stage, non_blocking=True)
            self.done.record(self.side)
            landed = torch.cuda.Event()
            landed.record(self.side)
        self.jobs.put((k, self.count, landed))
        self.count += 1

    def finish(self) -> torch.Tensor:
        self.jobs.put(None)
        self.thread.join()
        if self.error is not None:
            raise self.error
        return self.host[:, :self.count] if self.keep else self.host


class DeviceSink:
    """Snapshots stay in HBM as ``[B, T, 2, H, W]`` (cfg 4: 64 samples x 101 snapshots x 1025^2 = 54 GB of the
    180 GB) -- the form ``PdeDataset`` trains from; no host traffic while the solver runs."""

    def __init__(self, B: int, T: int, H: int, W: int, device):
        self.data = torch.empty((B, T, 2, H, W), dtype=torch.float32, device=device)
        self.count = 0

    def __call__(self, i: int, u: torch.Tensor, v: torch.Tensor) -> None:
        B, _, _, H, W = self.data.shape
        self.data[:, self.count, 0].copy_(u.reshape(B, H, W))
        self.data[:, self.count, 1].copy_(v.reshape(B, H, W))
        self.count += 1

    def finish(self) -> torch.Tensor:
        return self.data[:, :self.count]


def generate(mesh: int = 257, para: int = 5, repeat: int = 20, mu_lo: float = 0.025, mu_hi: float = 0.065,
             steps: int = 20001, every: int = 200, seed: int = 10, mode: str = "as_written",
             endpoint: bool = True, dt: float = 1e-4, dx: float = 0.0125, device: Optional[torch.device] = None,
             rank: Optional[int] = None, world_size: Optional[int] = None, graph: bool = True,
             timings: Optional[dict] = None, keep: bool = True, on_device: bool = False) -> torch.Tensor:
    """This rank's shard of the dataset as a host tensor [b, T, 2, mesh, mesh] (``on_device``: a CUDA tensor,
    snapshots never leave HBM).
    Defaults are the reference's (5 nu x 20 ICs, 257^2, 20 001 steps, snapshot every 200)."""
    t_begin = time.perf_counter()
    if rank is None or world_size is None:
        rank, world_size = D.world()
    device = device or torch.device("cuda", torch.cuda.current_device())
    P = para * repeat
    lo, hi = D.shard_bounds(P, rank, world_size)
    b = hi - lo
    # mu = linspace(lo, hi, para).repeat(repeat)  (operators.py:411-414): sample p has mu[p % para]
    mu_all = torch.linspace(mu_lo, mu_hi, para).repeat(repeat)
    mu = mu_all[lo:hi].to(device)
    coef = D.fourier_coefficients(seed, lo, hi)                     # [b, 2, 81, 2]
    u0 = fourier_initial_condition(coef[:, 0], mesh, device)
    v0 = fourier_initial_condition(coef[:, 1], mesh, device)
    T = len(range(0, steps, every))
    writer = DeviceSink(b, T, mesh, mesh, device) if on_device else SnapshotWriter(b, T, mesh, mesh, device, keep=keep)
    solver = FineBurgers(dt=dt, dx=dx, dy=dx, mode=mode, endpoint=endpoint)
    if timings is not None:
        torch.cuda.synchronize(device)
        timings["setup_seconds"] = time.perf_counter() - t_begin       # ICs + pinned staging buffers
        t_begin = time.perf_counter()
    solver.run(u0, v0, mu, steps, every=every, graph=graph, sink=writer)
    out = writer.finish()
    if timings is not None:
        torch.cuda.synchronize(device)
        timings["solve_seconds"] = time.perf_counter() - t_begin       # time stepping + snapshot D2H
    return out


def main(argv=None) -> None:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--out", required=True)
    ap.add_argument("--mesh", type=int, default=257)
    ap.add_argument("--para", type=int, default=5)
    ap.add_argument("--repeat", type=int, default=20)
    ap.add_argument("--mu-lo", type=float, default=0.025)
    ap.add_argument("--mu-hi", type=float, default=0.065)
    ap.add_argument("--steps", type=int, default=20001)
    ap.add_argument("--every", type=int, default=200)
    ap.add_argument("--seed", type=int, default=10)
    ap.add_argument("--mode", default="as_written", choices=["as_written", "intended"])
    ap.add_argument("--true-periodic", action="store_true", help="grid without the duplicated endpoint")
    ap.add_argument("--gather", action="store_true", help="gather all shards into one file on rank 0")
    ap.add_argument("--device-sink", action="store_true",
                    help="keep the snapshots in HBM while the solver runs (one device-to-host copy before saving)")
    ap.add_argument("--no-save", action="store_true", help="generate and time only (throughput runs)")
    ap.add_argument("--discard", action="store_true",
                    help="drop every snapshot once it has landed in pinned host memory (implies --no-save): "
                         "throughput of solver + D2H without the host-side accumulation")
    a = ap.parse_args(argv)
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, ws = D.world()
    # the drain thread's copies share the host cores with the other ranks of the node
    local_ws = int(os.environ.get("LOCAL_WORLD_SIZE", str(ws)))
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // max(1, local_ws)))
    if ws > 1:
        torch.distributed.barrier(device_ids=[local])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tm = {}
    data = generate(a.mesh, a.para, a.repeat, a.mu_lo, a.mu_hi, a.steps, a.every, a.seed, a.mode,
                    not a.true_periodic, timings=tm, keep=not a.discard, on_device=a.device_sink)
    torch.cuda.synchronize()
    sec = torch.tensor([time.perf_counter() - t0, tm["setup_seconds"], tm["solve_seconds"]], dtype=torch.float64, device="cuda")
    if ws > 1:
        torch.distributed.all_reduce(sec, op=torch.distributed.ReduceOp.MAX)      # slowest rank
    if rank == 0:
        P = a.para * a.repeat
        sec, setup, solve = sec.tolist()
        print(json.dumps({"datagen": f"{P} samples x {a.mesh}^2 x {a.steps} steps, snapshot every {a.every} ({a.mode})",
                          "n_gpus": ws, "seconds": sec, "setup_seconds": setup, "solve_seconds": solve,
                          "samples_per_s": P / sec,
                          "gpoint_updates_per_s": P * 2 * a.mesh * a.mesh * a.steps / sec / 1e9,
                          "solve_gpoint_updates_per_s": P * 2 * a.mesh * a.mesh * a.steps / solve / 1e9,
                          "snapshots": "kept in HBM" if a.device_sink else "discarded after D2H" if a.discard else "kept in host memory",
                          "host_threads_per_rank": torch.get_num_threads(),
                          "note": "max over ranks; setup = initial conditions + pinned staging buffers, "
                                  "solve = time stepping, snapshot D2H on a side stream and the drain into the host tensor; file write excluded"}),
              flush=True)
    if a.no_save or a.discard:
        pass
    elif a.device_sink and not (a.gather and ws > 1):
        torch.save({"u": data.cpu()}, D.shard_file(a.out, rank, ws))
    elif a.gather and ws > 1:
        full = D.gather_shards(data.cuda(), a.para * a.repeat)
        if rank == 0:
            torch.save({"u": full.cpu()}, a.out)
    else:
        torch.save({"u": data.clone()}, D.shard_file(a.out, rank, ws))
    if ws > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
