"""Parity of the on-chip multi-step solvers (csrc/onchip_solvers.cu: Kuramoto-Sivashinsky RK4 in 2-D and 1-D,
src/operators.py:750-918; the Bv/ heat step with Dirichlet / Neumann ghost cells, Bv/data.py) against the
reference's own outputs (tests/golden/solvers.npz) and the CPU oracle.  Metric: SURVEY.md 8(c), rel-L2 and
max-abs/||ref||_inf <= 1e-5 per step, fp32."""
import json
import math
import os

import numpy as np
import pytest
import torch

import ppnn_oracle as O
from conftest import ROOT, assert_parity, load_golden, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-5
PI8 = math.pi / 8


@pytest.fixture(scope="module")
def S():
    import ppnn_b200.solvers as solvers
    from ppnn_b200 import _lib
    _lib.load()
    return solvers


# ------------------------------------------------------------------------------------------------- 2-D KS
def test_ks2d_against_reference_golden(S):
    g = load_golden("solvers.npz")
    ks = S.KS2D(dt=1e-3, dx=PI8)
    u = g["ks2d/u0"].cuda()
    # the right-hand side itself: with dt = 1e-3 a step moves u by ~1e-3 |f|, so parity of u alone would not pin f
    assert_parity(ks.f(u).cpu(), g["ks2d/f0"], TOL, "ks2d f (fused kernel)")
    for i in range(3):                                   # one launch per step
        u = ks.rk4(u)
        assert_parity(u.cpu(), g[f"ks2d/u{i + 1}"], TOL, f"ks2d step {i + 1}")
    # <= 37 samples run one sample per thread-block cluster (row slabs + DSMEM ghost rows), more run one per CTA:
    # the same golden vectors through the other kernel (scaled copies behind them so that the samples differ), and
    # the two kernels against each other bit for bit
    scale = torch.cat((torch.ones(1), torch.linspace(0.5, 1.5, 19))).reshape(-1, 1, 1, 1, 1)
    many = (g["ks2d/u0"].unsqueeze(0) * scale).reshape(40, 1, 65, 65).cuda()
    fm, um = ks.f(many), ks.rk4(many)
    assert_parity(fm[:2].cpu(), g["ks2d/f0"], TOL, "ks2d f, one CTA per sample")
    assert_parity(um[:2].cpu(), g["ks2d/u1"], TOL, "ks2d step, one CTA per sample")
    few_f = torch.cat([ks.f(many[k:k + 2]) for k in range(0, 40, 2)])
    few_u = torch.cat([ks.run(many[k:k + 2], 2) for k in range(0, 40, 2)])
    assert torch.equal(fm, few_f), "f: cluster kernel and single-CTA kernel differ"
    assert torch.equal(ks.run(many, 2), few_u), "2 steps: cluster kernel and single-CTA kernel differ"
    u3 = ks.run(g["ks2d/u0"].cuda(), 3)                  # three steps inside one launch
    assert torch.equal(u3, u), "multi-step launch differs from step-by-step launches"
    assert_parity(u3.cpu(), g["ks2d/u3"], TOL, "ks2d 3 steps in one launch")
    # the Tier-A composition of the same closure (operator kernels), the path taken by oversize meshes
    w = g["ks2d/u0"].cuda()
    assert_parity(ks.f_composed(w).cpu(), g["ks2d/f0"], TOL, "ks2d f from Tier-A operators")
    assert_parity(ks.rk4_composed(w).cpu(), g["ks2d/u1"], TOL, "ks2d rk4 from Tier-A operators")


@pytest.mark.parametrize("h,w,endpoint,acc", [(64, 64, False, (3, 6, 4)), (33, 47, True, (3, 6, 4)), (40, 40, True, (2, 4, 2)),
                                              (97, 97, True, (3, 6, 4)), (114, 113, True, (3, 6, 4)), (31, 30, False, (2, 4, 2))])
def test_ks2d_sizes_against_oracle(S, h, w, endpoint, acc):
    """every points-per-thread instantiation (1 .. 24), both periodic conventions, both stencil families"""
    torch.manual_seed(h * 1000 + w)
    x = torch.linspace(0, 8 * math.pi, h).reshape(1, 1, h, 1)
    y = torch.linspace(0, 8 * math.pi, w).reshape(1, 1, 1, w)
    u = torch.cos(x / 4 + torch.randn(3, 1, 1, 1)) * (1 + torch.sin(y / 4 + torch.randn(3, 1, 1, 1)))
    u[1, 0, 5, 7] = 0.0
    kw = dict(adv_acc=acc[0], d2_acc=acc[1], d4_acc=acc[2], period_endpoint=endpoint)
    ref = O.ks2d_rk4_step(O.ks2d_rk4_step(u, 1e-3, PI8, **kw), 1e-3, PI8, **kw)
    ks = S.KS2D(dt=1e-3, dx=PI8, adv_accuracy=acc[0], d2_accuracy=acc[1], d4_accuracy=acc[2], endpoint=endpoint)
    assert_parity(ks.f(u.cuda()).cpu(), O.ks2d_rhs(u, PI8, PI8, **kw), TOL, f"ks2d f {h}x{w}")
    assert_parity(ks.run(u.cuda(), 2).cpu(), ref, TOL, f"ks2d {h}x{w}")
    # a step large enough for the increment to matter (a few times the production dt): pins the RK4 staging, not just u + O(dt)
    big = 0.004
    kb = S.KS2D(dt=big, dx=PI8, adv_accuracy=acc[0], d2_accuracy=acc[1], d4_accuracy=acc[2], endpoint=endpoint)
    assert_parity(kb.rk4(u.cuda()).cpu(), O.ks2d_rk4_step(u, big, PI8, **kw), TOL, f"ks2d {h}x{w}, dt={big}")


def test_ks2d_oversize_mesh_composes_tier_a(S):
    torch.manual_seed(3)
    n = 129                                              # 4 * 129^2 floats do not fit in shared memory
    x = torch.linspace(0, 16 * math.pi, n)
    x, y = torch.meshgrid(x, x, indexing="ij")
    u = (torch.cos(x / 4) * (1 + torch.sin(y / 4))).reshape(1, 1, n, n).repeat(2, 1, 1, 1)
    u[1] *= 0.7
    ks = S.KS2D(dt=1e-3, dx=PI8)
    assert_parity(ks.rk4(u.cuda()).cpu(), O.ks2d_rk4_step(u, 1e-3, PI8), TOL, "ks2d 129^2 (composed)")


# ------------------------------------------------------------------------------------------------- 1-D KS
def test_ks1d_against_reference_golden(S):
    g = load_golden("solvers.npz")
    ks = S.KS1D(dt=1e-4, dx=PI8, source=g["ks1d/source"].cuda())
    u = g["ks1d/u0"].cuda()
    assert_parity(ks.f(u).cpu(), g["ks1d/f0"], TOL, "ks1d f (fused kernel)")
    assert_parity(S.KS1D(dt=1e-4, dx=PI8).f(u).cpu(), g["ks1d/rhs_ks1d"], TOL, "ks1d f without source = rhs.ks1d")
    for i in range(3):
        u = ks.rk4(u)
        assert_parity(u.cpu(), g[f"ks1d/u{i + 1}"], TOL, f"ks1d step {i + 1}")
    assert torch.equal(ks.run(g["ks1d/u0"].cuda(), 3), u)
    # rhs.ks1d through the Tier-A operator classes, called as operators.py:883-884 calls it
    import ppnn_b200.operators as ops
    import ppnn_b200.rhs as rhs
    u0 = g["ks1d/u0"].cuda()
    pad = torch.cat((u0[:, :, -4:-1], u0, u0[:, :, 1:4]), dim=2)
    r = rhs.ks1d(pad, None, ops.dudx_1D('Upwind', accuracy=3, device="cuda"), ops.diffusion1D('Central', accuracy=6, device="cuda"),
                 ops.diffusion1D('Central_4th', accuracy=4, device="cuda"), pad, PI8, PI8 ** 2, PI8 ** 4)
    assert_parity(r.cpu(), g["ks1d/rhs_ks1d"], TOL, "rhs.ks1d (Tier A)")


@pytest.mark.parametrize("n,endpoint,acc", [(129, True, (3, 6, 4)), (128, False, (3, 6, 4)), (1000, True, (2, 4, 2)),
                                            (2049, True, (3, 6, 4)), (8192, False, (3, 6, 4)), (31, True, (3, 8, 6))])
def test_ks1d_sizes_against_oracle(S, n, endpoint, acc):
    torch.manual_seed(n)
    x = torch.linspace(0, 8 * math.pi * n / 129, n).reshape(1, 1, n)
    u = torch.cos(x / 16 + torch.randn(5, 1, 1)) * (1 + torch.sin(x / 16 + torch.randn(5, 1, 1)))
    u[2, 0, n // 3] = 0.0
    src = 0.1 * torch.exp(-(x - 4 * math.pi) ** 2 / 2)
    if acc == (3, 8, 6):        # stencils of different halos: restate f with one pad per operator
        def f(w):
            def pad(v, h):
                return torch.cat((v[:, :, -h - 1:-1], v, v[:, :, 1:h + 1]), dim=2)
            return (-O.stencil1d(pad(w, 4), "Central", 8) / PI8 ** 2 - O.stencil1d(pad(w, 4), "Central_4th", 6) / PI8 ** 4
                    - 0.5 * O.stencil1d(pad(w * w, 3), "Upwind", 3) / PI8 + src)
        ref = O.rk4_step(f, O.rk4_step(f, u, 1e-4), 1e-4)
        fref, big_ref = f(u), O.rk4_step(f, u, 0.003)
    else:
        kw = dict(adv_acc=acc[0], d2_acc=acc[1], d4_acc=acc[2], period_endpoint=endpoint)
        ref = O.ks1d_rk4_step(O.ks1d_rk4_step(u, 1e-4, PI8, src, **kw), 1e-4, PI8, src, **kw)
        fref, big_ref = O.ks1d_rhs(u, PI8, src, **kw), O.ks1d_rk4_step(u, 0.003, PI8, src, **kw)
    ks = S.KS1D(dt=1e-4, dx=PI8, source=src.cuda(), adv_accuracy=acc[0], d2_accuracy=acc[1], d4_accuracy=acc[2],
                endpoint=endpoint)
    assert_parity(ks.f(u.cuda()).cpu(), fref, TOL, f"ks1d f n={n}")
    assert_parity(ks.run(u.cuda(), 2).cpu(), ref, TOL, f"ks1d n={n}")
    kb = S.KS1D(dt=0.003, dx=PI8, source=src.cuda(), adv_accuracy=acc[0], d2_accuracy=acc[1], d4_accuracy=acc[2],
                endpoint=endpoint)
    assert_parity(kb.rk4(u.cuda()).cpu(), big_ref, TOL, f"ks1d n={n}, dt=0.003")


def test_ks_batch_independence_and_generate(S):
    """size-independent properties at a batch the oracle would not finish: a sample's result does not depend on
    its position in the batch (bit-exact under a flip), and generate() snapshots equal a straight run"""
    torch.manual_seed(11)
    B, n = 4096, 129
    x = torch.linspace(0, 8 * math.pi, n, device="cuda").reshape(1, 1, n)
    u = torch.cos(x / 16 + torch.randn(B, 1, 1, device="cuda")) * (1 + torch.sin(x / 16 + torch.randn(B, 1, 1, device="cuda")))
    ks = S.KS1D(dt=1e-4, dx=PI8)
    a = ks.run(u, 50)
    b = ks.run(u.flip(0).contiguous(), 50).flip(0)
    assert torch.equal(a, b)
    assert torch.isfinite(a).all()
    snaps = ks.generate(u[:8], 25, every=10)             # snapshots after steps 1, 11, 21
    assert snaps.shape == (8, 3, 1, n)
    assert torch.equal(snaps[:, 0], ks.run(u[:8], 1))
    assert torch.equal(snaps[:, 2], ks.run(u[:8], 21))
    k2 = S.KS2D(dt=1e-3, dx=PI8)
    x2 = torch.linspace(0, 8 * math.pi, 65, device="cuda")
    w = (torch.cos(x2.reshape(-1, 1) / 4) * (1 + torch.sin(x2.reshape(1, -1) / 4))).reshape(1, 1, 65, 65)
    w = w * torch.linspace(0.5, 1.5, 300, device="cuda").reshape(-1, 1, 1, 1)
    a2 = k2.run(w, 10)
    assert torch.equal(a2, k2.run(w.flip(0).contiguous(), 10).flip(0))


def test_ks_drift_100_steps(S):
    """100-step drift of both KS solvers against the fp32 and the fp64 oracle (north_star: drift reported)."""
    g = load_golden("solvers.npz")
    out = {}
    for name in ("ks2d", "ks1d"):
        u0 = g[f"{name}/u0"]
        if name == "ks2d":
            solver = S.KS2D(dt=1e-3, dx=PI8)
            step32 = lambda w: O.ks2d_rk4_step(w, 1e-3, PI8)
            step64 = step32
        else:
            src = g["ks1d/source"].reshape(1, 1, -1)
            solver = S.KS1D(dt=1e-4, dx=PI8, source=src.cuda())
            step32 = lambda w: O.ks1d_rk4_step(w, 1e-4, PI8, src)
            step64 = lambda w: O.ks1d_rk4_step(w, 1e-4, PI8, src.double())
        ours, r32, r64 = u0.cuda(), u0.clone(), u0.double()
        curve = []
        for i in range(100):
            ours, r32, r64 = solver.rk4(ours), step32(r32), step64(r64)
            curve.append((rel_l2(ours.cpu(), r32), rel_l2(ours.cpu(), r64), rel_l2(r32, r64)))
        out[name] = {"steps": 100, "ours_vs_fp32_oracle": [c[0] for c in curve],
                     "ours_vs_fp64_oracle": [c[1] for c in curve], "fp32_vs_fp64_oracle": [c[2] for c in curve]}
        assert curve[0][0] <= TOL
        # after 100 steps we must sit at the reference's own rounding drift, not above it by more than 10x
        assert curve[-1][1] <= 10 * max(curve[-1][2], 1e-7), (name, curve[-1])
        assert rel_l2(solver.run(u0.cuda(), 100).cpu(), ours.cpu()) == 0.0
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "drift_solvers.json"), "w") as f:
        json.dump(out, f)


# ------------------------------------------------------------------------------------------------- Bv heat
def test_bv_heat_against_reference_golden(S):
    import ppnn_b200.bv as bv
    import ppnn_b200.operators as ops
    g = load_golden("solvers.npz")
    cls, bvr, dx, dt = g["bv/cls"].cuda(), g["bv/BVr"].cuda(), float(g["bv/dx"]), float(g["bv/dt"])
    save = int(g["bv/saveevery"])
    assert torch.equal(bv.denormBV(g["bv/BV"].cuda(), cls).cpu(), g["bv/BVr"])
    heat = S.HeatBC1D(dt, dx)
    u0 = g["bv/u0"].cuda()
    assert_parity(heat.pad(u0, cls, bvr).cpu(), g["bv/pad0"], TOL, "padder (fused class)")
    assert_parity(bv.padder(dx)(u0, cls, bvr).cpu(), g["bv/pad0"], TOL, "padder (Bv mirror)")
    res = heat.generate(u0, cls, bvr, 151, save)
    assert_parity(res.cpu(), g["bv/results"], TOL, "generator snapshots")
    assert_parity(heat.run(u0, cls, bvr, 151).cpu(), g["bv/u_final"], TOL, "final state, one launch")
    # the same loop as Bv/data.py:92-97 writes it, on the Tier-A classes
    pad, diffuer, u = bv.padder(dx), ops.diffusion1D(accuracy=2, device="cuda"), u0
    for t in range(50):
        up = pad(u, cls, bvr)
        u = u + diffuer(up) / heat.dx2 * dt
    assert_parity(pad(u, cls, bvr).cpu(), g["bv/results"][:, 1], TOL, "Tier-A loop, 50 steps")
    scale = float(g["bv/scale"])
    inc = S.HeatCoarseIncrement(scale, dt * save, dx / scale)
    fwd = bv.genforward(diffuer, lambda w: torch.nn.functional.interpolate(w, scale_factor=scale, mode="linear", align_corners=True),
                        lambda w: torch.nn.functional.interpolate(w, scale_factor=1 / scale, mode="linear", align_corners=True),
                        bv.padder(dx / scale), dt * save, inc.cdx2)
    for t in range(g["bv/results"].shape[1]):
        r = g["bv/results"][:, t].cuda()
        assert_parity(inc(r, cls, bvr).cpu(), g["bv/pdeu"][:, t], TOL, f"genforward fused t={t}")
        assert_parity(fwd(r, cls, bvr).cpu(), g["bv/pdeu"][:, t], TOL, f"genforward composed t={t}")


def test_bv_rollout_update(S):
    """Bv/train.py:162: u <- fpad(u[...,1:-1] + nnu + pdeu[...,1:-1], cls, bv), with and without the PDE increment"""
    g = load_golden("solvers.npz")
    cls, bvr, dx = g["bv/cls"], g["bv/BVr"], g["bv/dx"]
    u = g["bv/results"][:, 1]
    pdeu = g["bv/pdeu"][:, 1]
    torch.manual_seed(5)
    nn = 0.05 * torch.randn(u.shape[0], 1, u.shape[-1] - 2)
    heat = S.HeatBC1D(float(g["bv/dt"]), float(dx))
    ref = O.bv_pad(u[..., 1:-1] + nn + pdeu[..., 1:-1], cls, bvr, dx)
    assert_parity(heat.step_update(u.cuda(), nn.cuda(), pdeu.cuda(), cls.cuda(), bvr.cuda()).cpu(), ref, TOL, "bv update")
    ref0 = O.bv_pad(u[..., 1:-1] + nn, cls, bvr, dx)
    assert_parity(heat.step_update(u.cuda(), nn.cuda(), None, cls.cuda(), bvr.cuda()).cpu(), ref0, TOL, "bv update, no pde")


@pytest.mark.parametrize("n,b", [(126, 256), (30, 3), (5000, 4), (1, 2)])
def test_bv_heat_sizes_against_oracle(S, n, b):
    torch.manual_seed(n + b)
    dx, dt = torch.tensor(0.4), 0.01
    u = torch.randn(b, 1, n).cumsum(-1) * 0.2
    cls = torch.randint(0, 2, (b, 1, 2))
    bv_ = torch.rand(b, 1, 2) * 2 - 1
    uf, res = O.bv_heat_run(u, cls, bv_, 21, 10, dt, dx, dx ** 2)
    heat = S.HeatBC1D(dt, float(dx))
    assert_parity(heat.generate(u.cuda(), cls.cuda(), bv_.cuda(), 21, 10).cpu(), res, TOL, f"bv n={n}")
    assert_parity(heat.run(u.cuda(), cls.cuda(), bv_.cuda(), 21).cpu(), uf, TOL, f"bv final n={n}")
    if n >= 30:
        up = O.bv_pad(u, cls, bv_, dx)
        nf = n + 2
        if nf % 4 == 0:
            ref = O.bv_coarse_increment(up, cls, bv_, 0.25, 0.5, dx * 4, (dx * 4) ** 2)
            inc = S.HeatCoarseIncrement(0.25, 0.5, float(dx * 4))
            assert_parity(inc(up.cuda(), cls.cuda(), bv_.cuda()).cpu(), ref, TOL, f"genforward n={n}")


def test_solvers_reject_cpu_tensors(S):
    with pytest.raises(RuntimeError):
        S.KS1D().rk4(torch.zeros(2, 1, 129))
    with pytest.raises(RuntimeError):
        S.KS2D().rk4(torch.zeros(2, 1, 65, 65))
    with pytest.raises(RuntimeError):
        S.HeatBC1D(0.01, 0.4).run(torch.zeros(2, 1, 30), torch.zeros(2, 1, 2), torch.zeros(2, 1, 2), 3)


def test_solver_throughput_report(S):
    """Timing at the reference's own generator shapes, three ways on the same box: the on-chip kernel (all steps in
    one launch), the oracle's ATen op sequence on CUDA tensors (what the reference does on a GPU: one launch per
    op), and the same on the host cores.  Written to gpurun_out/solvers_perf.json; asserts only sanity."""
    import time
    dev = torch.device("cuda")

    def gpu_ms(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        best = float("inf")
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best

    def cpu_ms(fn, reps=2):
        fn()
        best = float("inf")
        for _ in range(reps):
            t = time.perf_counter()
            fn()
            best = min(best, (time.perf_counter() - t) * 1e3)
        return best

    report = {}
    # 1-D KS: 256 samples x 129 points (operators.py:849,855), dt 1e-4
    x = torch.linspace(0, 8 * math.pi, 129).reshape(1, 1, -1)
    torch.manual_seed(42)
    u = torch.cos(x / 16 + torch.randn(256, 1, 1)) * (1 + torch.sin(x / 16 + torch.randn(256, 1, 1))) + torch.exp(-25 * (x / 8 / math.pi - 0.5) ** 2)
    src = 0.1 * torch.exp(-(x - 4 * math.pi) ** 2 / 2)
    ks1 = S.KS1D(dt=1e-4, dx=PI8, source=src.to(dev))
    ud, sd = u.to(dev), src.to(dev)
    n_fused = 20000
    t_fused = gpu_ms(lambda: ks1.run(ud, n_fused)) / n_fused
    t_aten = gpu_ms(lambda: [O.ks1d_rk4_step(ud, 1e-4, PI8, sd) for _ in range(10)]) / 10
    t_cpu = cpu_ms(lambda: [O.ks1d_rk4_step(u, 1e-4, PI8, src) for _ in range(10)]) / 10
    report["ks1d_256x129"] = {"fused_us_per_step": t_fused * 1e3, "aten_cuda_us_per_step": t_aten * 1e3,
                              "aten_cpu_us_per_step": t_cpu * 1e3, "steps_in_one_launch": n_fused,
                              "gpt_updates_per_s": 256 * 129 / (t_fused * 1e-3) / 1e9}
    # 2-D KS: the reference runs ONE 65^2 sample (operators.py:754-757); also an ensemble of 296
    x2 = torch.linspace(0, 8 * math.pi, 65)
    w = (torch.cos(x2.reshape(-1, 1) / 4) * (1 + torch.sin(x2.reshape(1, -1) / 4))).reshape(1, 1, 65, 65)
    ks2 = S.KS2D(dt=1e-3, dx=PI8)
    for b in (1, 296):
        wb = (w * torch.linspace(0.9, 1.1, b).reshape(-1, 1, 1, 1)).contiguous()
        wd = wb.to(dev)
        n_fused = 2000
        t_fused = gpu_ms(lambda: ks2.run(wd, n_fused)) / n_fused
        t_aten = gpu_ms(lambda: [O.ks2d_rk4_step(wd, 1e-3, PI8) for _ in range(5)]) / 5
        t_cpu = cpu_ms(lambda: [O.ks2d_rk4_step(wb, 1e-3, PI8) for _ in range(2)]) / 2
        report[f"ks2d_{b}x65x65"] = {"fused_us_per_step": t_fused * 1e3, "aten_cuda_us_per_step": t_aten * 1e3,
                                     "aten_cpu_us_per_step": t_cpu * 1e3, "steps_in_one_launch": n_fused,
                                     "gpt_updates_per_s": b * 65 * 65 / (t_fused * 1e-3) / 1e9}
    # Bv heat: 256 samples x 126 points, 10001 steps, snapshot every 50 (Bv/data.py:56-63,92-98)
    torch.manual_seed(10)
    ub = torch.randn(256, 1, 126).cumsum(-1) * 0.1
    cls, bvv = torch.randint(0, 2, (256, 1, 2)), torch.rand(256, 1, 2) * 2 - 1
    dxb = torch.tensor(16 * math.pi / 127)
    heat = S.HeatBC1D(0.01, float(dxb))
    ubd, cd, bd = ub.to(dev), cls.to(dev), bvv.to(dev)
    t_fused = gpu_ms(lambda: heat.generate(ubd, cd, bd, 10001, 50)) / 10001
    t_aten = gpu_ms(lambda: O.bv_heat_run(ubd, cd, bd, 100, 50, 0.01, dxb.to(dev), dxb.to(dev) ** 2)) / 100
    t_cpu = cpu_ms(lambda: O.bv_heat_run(ub, cls, bvv, 100, 50, 0.01, dxb, dxb ** 2)) / 100
    report["bv_heat_256x126"] = {"fused_us_per_step": t_fused * 1e3, "aten_cuda_us_per_step": t_aten * 1e3,
                                 "aten_cpu_us_per_step": t_cpu * 1e3, "steps": 10001, "snapshots": 201,
                                 "gpt_updates_per_s": 256 * 126 / (t_fused * 1e-3) / 1e9}
    report["host_threads"] = torch.get_num_threads()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "solvers_perf.json"), "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report))
    for k, v in report.items():
        if isinstance(v, dict):
            assert v["fused_us_per_step"] < v["aten_cuda_us_per_step"], (k, v)
