"""N > 1 host logic on CPU: world-size-2 gloo processes exercise the sharding,
the per-sample RNG streams (union of shards == single-process run), the global
normalisation statistics (train2D.py:115-151 over sharded data) and the shard
gather.  The compute stand-in is the oracle (this is test infrastructure)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

from ppnn_b200 import dist as D  # noqa: E402


def test_shard_bounds_partition_every_batch():
    for n in (0, 1, 7, 64, 512, 1921):
        for ws in (1, 2, 3, 8):
            edges = [D.shard_bounds(n, r, ws) for r in range(ws)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_bounds(8, 2, 2)
    assert D.shard_file("a/b.pt", 3, 8) == "a/b.rank003-of-008.pt" and D.shard_file("a/b.pt", 0, 1) == "a/b.pt"


def test_fourier_coefficients_depend_on_global_index_only():
    full = D.fourier_coefficients(10, 0, 11)
    parts = torch.cat([D.fourier_coefficients(10, *D.shard_bounds(11, r, 3)) for r in range(3)])
    assert torch.equal(full, parts)
    assert not torch.equal(full[0], full[1])


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ppnn_oracle as O
        P = 7                                                # ragged: 4 + 3 samples
        lo, hi = D.shard_bounds(P, rank, world)
        coef = D.fourier_coefficients(10, lo, hi)
        u = torch.cat((O.fourier_initial_condition(coef[:, 0], 33) - 0.5,
                       O.fourier_initial_condition(coef[:, 1], 33) - 0.5), dim=1)
        mu = torch.linspace(0.02, 0.07, P)[lo:hi].reshape(-1, 1, 1, 1)
        step = O.CoarseStep(33, 9, "burgers2Dpu", "burgers2Dpv", 0.02, 0.8)
        pdeu = step.pde_du(u, mu)
        mean, std = D.channel_stats(pdeu[:, :, :-1, :-1])
        full = D.gather_shards(pdeu, P)
        torch.save({"u": u.unsqueeze(1)}, D.shard_file(os.path.join(tmp, "d.pt"), rank, world))
        dist.barrier()
        if rank == 0:
            # single-process reference of the same job
            coef = D.fourier_coefficients(10, 0, P)
            ua = torch.cat((O.fourier_initial_condition(coef[:, 0], 33) - 0.5,
                            O.fourier_initial_condition(coef[:, 1], 33) - 0.5), dim=1)
            ref = step.pde_du(ua, torch.linspace(0.02, 0.07, P).reshape(-1, 1, 1, 1))
            assert torch.equal(full, ref), "union of shards differs from the 1-process run"
            m, s = ref[:, :, :-1, :-1].mean(dim=(0, 2, 3), keepdim=True), ref[:, :, :-1, :-1].std(dim=(0, 2, 3), keepdim=True)
            assert torch.allclose(mean, m, rtol=1e-5, atol=1e-9) and torch.allclose(std, s, rtol=1e-5, atol=1e-9)
            merged = D.merge_shard_files(os.path.join(tmp, "d.pt"), world)
            assert torch.equal(merged[:, 0], ua)
            open(os.path.join(tmp, "ok"), "w").write("ok")
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


def _stats_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ppnn_b200.dataset import finalize_stats
        gen = torch.Generator().manual_seed(5)
        x = torch.randn(3, 7, 2, 400, generator=gen) * torch.tensor([1.0, 3.0, 0.2]).reshape(3, 1, 1, 1) + 0.7
        lo, hi = D.shard_bounds(7, rank, world)                 # ragged shards: 4 + 3 samples
        xs = x[:, lo:hi].double()
        sums = torch.stack((xs.sum(dim=(1, 3)), (xs * xs).sum(dim=(1, 3))), dim=-1)   # [3 stats, 2 ch, 2]
        sums = sums.permute(1, 0, 2).reshape(2, 6)               # the kernel's [channel][6] layout
        counts = torch.full((3,), float((hi - lo) * 400), dtype=torch.float64)
        mean, std = finalize_stats(sums, counts)
        ref_m = x.mean(dim=(1, 3))
        ref_s = x.std(dim=(1, 3))
        assert torch.allclose(mean, ref_m, rtol=1e-6, atol=1e-7) and torch.allclose(std, ref_s, rtol=1e-6)
        # the parameter statistics of train2D.py:149-150 over RAGGED shards with different mu subsets
        # (each rank alone would get another mean / std): every rank must see the global values
        mus = torch.linspace(0.02, 0.07, 7).reshape(-1, 1).repeat(1, 5).reshape(-1, 1, 1, 1) + 100.0   # |mean| >> std
        mm, ms = D.scalar_stats(mus[lo * 5:hi * 5])
        assert torch.allclose(mm, mus.mean(), rtol=1e-7) and torch.allclose(ms, mus.std(), rtol=1e-5), (mm, ms)
        assert not torch.allclose(mus[lo * 5:hi * 5].std(), mus.std(), rtol=1e-3)
        if rank == 0:
            open(os.path.join(tmp, "ok2"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_dataset_statistics_all_reduce_gloo(tmp_path):
    """Sharded training set: per-rank partial sums of ppnn_dataset_stats -> global mean / std."""
    port = 29900 + os.getpid() % 90
    mp.spawn(_stats_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok2").read_text() == "ok"


def test_world_size_2_gloo(tmp_path):
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").read_text() == "ok"
