#!/usr/bin/env python
"""bench.py -- throughput of the fused PDE-preserving step on B200.

    python bench.py --gpus N --steps K --warmup W            (our arm)
    python bench.py --impl reference --gpus N --steps K --warmup W   (CPU reference arm)
    python bench.py --sweep [quick|full] --gpus N            (BASELINE configs[4]: roofline sweep table)

Workload (BASELINE.json configs[1], cases/RD_pPDE.yaml): one "step" is ONE launch
of the fused coarse-grid step kernel over a batch of B samples per GPU

    u_next = padBC_rd(nn_out*outstd + outmean) + u + pde_du(u, mu)      (train2D.py:235-238)

with F = 257, C = 49, rhs = rdpu/rdpv, dt = 0.002, length = 6.4, mu = linspace(.2,1.3,12)
repeated, synthetic periodic fields (Fourier fields of operators.py:436-443, shifted to
mixed sign) and a random stand-in for the network output.  Unit: point-updates/s, one
point-update = one fp32 value of one field advanced one step.

* value      -- inputs resident in HBM, CUDA-event timed, max over ranks.
* e2e        -- the same step through the public API with HOST (pinned) buffers:
                H2D of u and nn_out, kernel, D2H of u_next, inside the timed region.
* roofline   -- HBM bound: 12 B per point-update (read u, read nn_out, write u_next)
                / measured launch time, against MEASURED_PEAKS.json.
* cpu_baseline -- the reference's own modules (operators.py / rhs.py / utility/utils.py staged
                byte-for-byte in oracle/_ref by the build container; kind "reference") on a bounded
                sample of the same workload on this host's cores; the oracle port
                (oracle/ppnn_oracle.py, kind "port") only if nothing was staged.
* parity     -- after the timed loop (outside it) 4 samples of the B-sample output are recomputed
                by that CPU reference and compared (rel-L2, max-abs / max|ref|).
Multi-GPU: samples are independent, so ranks own disjoint shards of the batch and there
is no data-path collective (weak scaling: B per GPU fixed).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CASE = dict(fine=257, coarse=49, rhsu="rdpu", rhsv="rdpv", dt=0.002, length=6.4,
            paralow=0.2, parahigh=1.3, num_para=12)
B_PER_GPU = 1920            # 12 mu x 16 ICs x 10 time levels: 1/10 of the train2D.py:121 precompute batch
BYTES_PER_POINT = 12        # nominal: read u + read nn_out + write u_next (nn_out is (F-1)^2: see algorithmic_bytes)
METRIC = "fused_pde_step_point_updates_per_s"
UNIT = "Gpoint-updates/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="samples per GPU")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary kernel timings")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-legs", action="store_true", help="extras: kernels only (no training / datagen / rollout legs); for ncu runs")
    ap.add_argument("--cpu-batch", type=int, default=240)
    ap.add_argument("--sweep", nargs="?", const="full", choices=["quick", "full"],
                    help="BASELINE configs[4]: stencil roofline sweep (markdown table on stdout) instead of the bench line")
    return ap.parse_args()


# ----------------------------------------------------------------------------- inputs
def synth_fields(B, F, device, seed):
    """Smooth periodic [B,2,F,F] state: random 81-mode Fourier sums like the reference
    generator (operators.py:436-443), min-max normalised and shifted to mixed sign."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.linspace(0, 2 * torch.pi, F, device=device)
    kk = torch.arange(-4, 5, device=device, dtype=torch.float32)
    sx, cx = torch.sin(kk[:, None] * x[None]), torch.cos(kk[:, None] * x[None])      # [9,F]
    out = torch.empty(B, 2, F, F, device=device)
    chunk = 64
    for b0 in range(0, B, chunk):
        nb = min(chunk, B - b0)
        a = torch.randn(nb * 2, 9, 9, device=device, generator=g)
        c = torch.randn(nb * 2, 9, 9, device=device, generator=g)
        # sum_kl a_kl sin(kx+ly) + c_kl cos(kx+ly), expanded with the angle-sum identities
        f = (torch.einsum("nkl,ki,lj->nij", a, sx, cx) + torch.einsum("nkl,ki,lj->nij", a, cx, sx)
             + torch.einsum("nkl,ki,lj->nij", c, cx, cx) - torch.einsum("nkl,ki,lj->nij", c, sx, sx))
        lo = f.amin(dim=(1, 2), keepdim=True)
        hi = f.amax(dim=(1, 2), keepdim=True)
        out[b0:b0 + nb] = ((f - lo) / (hi - lo) + 0.1 - 0.6).reshape(nb, 2, F, F)
    return out


def make_mu(B, device, offset=0):
    import torch
    base = torch.linspace(CASE["paralow"], CASE["parahigh"], CASE["num_para"], device=device)
    idx = (torch.arange(B, device=device) + offset) % CASE["num_para"]
    return base[idx].reshape(B, 1, 1, 1).contiguous()


STATS = dict(outmean=[1.0e-4, -2.0e-4], outstd=[3.0e-3, 2.0e-3])


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,utilization.gpu,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        rows = []
        for line in open(self.tmp.name):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) == 7 and parts[0].isdigit():
                rows.append(parts)
        os.unlink(self.tmp.name)
        if not rows:
            return out
        busy = [r for r in rows if r[2].isdigit() and int(r[2]) > 0] or rows
        clk = sorted(int(r[0]) for r in busy)
        out["sm_mhz"] = clk[len(clk) // 2]
        out["sm_max_mhz"] = int(rows[0][1])
        out["samples"] = len(busy)
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        out["reasons"] = [n for k, n in enumerate(names) if any(r[3 + k] == "Active" for r in busy)]
        return out


def algorithmic_bytes(B):
    """Compulsory HBM traffic of one update launch: read u [B,2,F,F], read nn_out [B,2,F-1,F-1], write u_next."""
    F = CASE["fine"]
    return 4 * (2 * F * F + (F - 1) * (F - 1)) * 2 * B


# ----------------------------------------------------------------------------- CPU legs
def cpu_update_fn():
    """``(update(u, mu, nn, outmean, outstd) -> u_next, kind, description)`` on the CPU: the staged
    reference modules bound as train2D.py:66-79,235-238 does, else the oracle port."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    torch.set_num_threads(os.cpu_count() or 1)
    args = (CASE["fine"], CASE["coarse"], CASE["rhsu"], CASE["rhsv"], CASE["dt"], CASE["length"])
    import ref_closures as R
    if R.available():
        step = R.RefCoarseStep(*args)

        def update(u, mu, nn, om, os_):
            with torch.no_grad():
                return step.update(u, mu, nn, om, os_)
        return update, "reference", ("the reference's own operators.py / rhs.py / utility/utils.py (oracle/_ref, unmodified), "
                                     "bound as train2D.py:66-79,235-238")
    import ppnn_oracle as O
    step = O.CoarseStep(*args)

    def update(u, mu, nn, om, os_):
        with torch.no_grad():
            return O.ppnn_update(u, step.pde_du(u, mu), nn, om, os_)
    return update, "port", "oracle/ppnn_oracle.py (the reference's ATen op sequence: interpolate/conv2d/cat)"


def cpu_oracle_step_fn(B):
    """Bounded sample of the workload on the host cores."""
    import torch
    update, kind, desc = cpu_update_fn()
    F = CASE["fine"]
    u = synth_fields(B, F, torch.device("cpu"), 10)
    mu = make_mu(B, torch.device("cpu"))
    nn = torch.randn(B, 2, F - 1, F - 1, generator=torch.Generator().manual_seed(11))
    om = torch.tensor(STATS["outmean"]).reshape(1, 2, 1, 1)
    os_ = torch.tensor(STATS["outstd"]).reshape(1, 2, 1, 1)

    def run():
        return update(u, mu, nn, om, os_)
    return run, B * 2 * F * F, kind, desc


def time_cpu(B, min_seconds=12.0, max_reps=400):
    run, pts, kind, desc = cpu_oracle_step_fn(B)
    run()                                                   # warm-up
    t0 = time.perf_counter()
    reps = 0
    while reps < max_reps and (time.perf_counter() - t0 < min_seconds or reps < 2):
        run()
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    return pts / dt / 1e9, reps, dt, kind, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    B = args.cpu_batch
    run, pts, kind, desc = cpu_oracle_step_fn(B)
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    val = pts / dt / 1e9
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.batch, max(args.gpus, 1)),
            "cpu_sample_per_step": B,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{B} of {args.batch} samples per step (the metric is per point-update, so the "
                                       f"slice is comparable), {desc}, on {cores} torch threads"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(B, n):
    c = {"workload": "cases/RD_pPDE.yaml fused step: u_next = padBC_rd(nn*outstd+outmean) + u + pde_du(u,mu)",
         "fine": CASE["fine"], "coarse": CASE["coarse"], "rhs": [CASE["rhsu"], CASE["rhsv"]],
         "dt": CASE["dt"], "length": CASE["length"], "batch_per_gpu": B, "global_batch": B * n,
         "parallelism": f"batch-sharded x{n}, no data-path collective",
         "l2": "inputs (3.0 GB per GPU) exceed the 126 MB L2; no flush needed"}
    return c


# ----------------------------------------------------------------------------- config 5: roofline sweep
def cpu_sweep_fn():
    """CPU-reference column of the sweep: the staged reference's own operator / rhs objects (oracle/_ref) on the
    host cores, per (op, N) on a bounded sample (about 4 M grid points), G units/s."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_closures as R
    if not R.available():
        return None
    m = R.modules()
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    o = m.operators
    op6 = (o.dudx_2D("Upwind1", 3), o.dudy_2D("Upwind1", 3), o.d2udx2_2D("Central2", 6), o.d2udy2_2D("Central2", 6))
    d2x2, dy1 = o.d2udx2_2D("Central2", 2), o.dudy_2D("Upwind1", 1)

    def padx(w, h):
        return torch.cat((w[:, :, -h:], w, w[:, :, :h]), dim=2)

    def pady(w, h):
        return torch.cat((w[:, :, :, -h:], w, w[:, :, :, :h]), dim=3)

    def cpu(name, N):
        # bounded sample: B samples of N x N while they stay under 4 M grid points, else a strip of Hc rows x N columns
        # of one sample (every op is local, so the cost per point is that of the full mesh)
        B = max(1, (1 << 22) // (N * N))
        Hc = N if N * N <= (1 << 22) else max(64, (1 << 22) // N)
        g = torch.Generator().manual_seed(1)
        u = torch.rand(B, 1, Hc, N, generator=g) - 0.4
        v = torch.rand(B, 1, Hc, N, generator=g) - 0.4
        mu = torch.full((B, 1, 1, 1), 0.04)
        dx = 0.0125
        ux, uy, ux1, uy1 = padx(u, 3), pady(u, 3), padx(u, 1), pady(u, 1)

        def gen_step():                                  # operators.py:466-475 (true-periodic pads on these meshes)
            ax, ay = padx(u, 3), pady(u, 3)
            padx(v, 3), pady(v, 3)                       # :469-470: computed by the reference, unused as written
            u * u + v * v                                # :472
            un = u + 1e-4 * m.rhs.burgers2Dpu(u, v, mu, *op6, ax, ay, dx, dx, dx * dx, dx * dx)
            return un, v + 1e-4 * m.rhs.burgers2Dpv(un, v, mu, *op6, ax, ay, dx, dx, dx * dx, dx * dx)
        fns = {"d2udx2_2D Central2 acc 6 (x)": lambda: op6[2](ux), "d2udy2_2D Central2 acc 6 (y)": lambda: op6[3](uy),
               "dudx_2D Upwind1 acc 3 (x)": lambda: op6[0](ux), "dudy_2D Upwind1 acc 3 (y)": lambda: op6[1](uy),
               "d2udx2_2D Central2 acc 2 (x)": lambda: d2x2(ux1), "dudy_2D Upwind1 acc 1 (y)": lambda: dy1(uy1),
               "rhs.burgers2Dpu fused (acc 3/6)": lambda: m.rhs.burgers2Dpu(u, v, mu, *op6, ux, uy, dx, dx, dx * dx, dx * dx),
               "fused Burgers step (u,v)": gen_step}
        fn = fns[name]
        units = 2 if name.startswith("fused Burgers step") else 1
        with torch.no_grad():
            fn()
            t0, reps = time.perf_counter(), 0
            while reps < 2 or (time.perf_counter() - t0 < 0.25 and reps < 50):
                fn()
                reps += 1
            dt = (time.perf_counter() - t0) / reps
        return units * B * Hc * N / dt / 1e9, cores
    return cpu


def run_sweep(args):
    """`bench.py --sweep [quick|full] [--gpus N]`: tools/sweep.py on N GPUs (batch-sharded, no communication) with
    the reference's CPU implementation of every op timed beside it on rank 0."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --sweep: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sweep as SW
    cpu = None if args.no_cpu else cpu_sweep_fn()
    if rank == 0:
        print(f"# Stencil roofline sweep (BASELINE configs[4]), {world} x B200\n\n"
              f"`bench.py --sweep {args.sweep} --gpus {world}`: fp32, CUDA events, max over ranks, 6-20 repetitions after 3 "
              f"warm-ups; N x N periodic grids x GLOBAL batch B sharded over the GPUs (cells with B < GPUs skipped).  "
              f"CPU reference = the reference's own operator / rhs modules (oracle/_ref) on the host's "
              f"{os.cpu_count()} hardware threads, per (op, N) on a bounded sample (about 4 M grid points: whole samples, or a strip of rows of one sample at 4096^2 and beyond); units = output "
              f"points (operators, rhs) or point-updates (step).\n", flush=True)
    SW.sweep(dev, world, rank, quick=args.sweep == "quick", cpu=cpu, barrier=barrier, max_over_ranks=max_over_ranks)
    if world > 1:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- config 3 / config 4 legs
def train_leg(torch, dist, dev, world, rank, local, timed, L):
    """BASELINE configs[2]: one training step of cases/burgers_fPDE.yaml THROUGH the PDE-preserving step -- a
    4-step unrolled rollout u <- padBC_rd(net(norm u, norm mu)*outstd + outmean) + u + pde_du(u, mu) with the
    fused kernel forward and the adjoint kernel backward, a cnn2d-shaped network (ignores pdeu, models.py:95-96),
    100 samples per GPU (= the YAML's batchsize 800 on 8 GPUs), Adam; with N > 1 the network is wrapped in
    DistributedDataParallel: one NCCL all-reduce of its fp32 gradients per step over NVLink."""
    import ctypes as C
    from ppnn_b200 import CoarseStep
    from ppnn_b200.nets import ConvResNet
    Bt, unroll, F = 100, 4, 257
    step = CoarseStep(F, 49, "burgers2DFu", "burgers2DFv", 0.02, 3.2)
    torch.manual_seed(10)
    net = ConvResNet(uses_pde=False).to(dev)
    nparams = sum(p.numel() for p in net.parameters())
    model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local]) if world > 1 else net
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    u0 = synth_fields(Bt, F, dev, 300 + rank)
    target = synth_fields(Bt, F, dev, 400 + rank)
    mu = torch.linspace(0.02, 0.07, 6, device=dev)[torch.arange(Bt, device=dev) % 6].reshape(Bt, 1, 1, 1)
    mun = (mu - 0.045) / 0.017
    inmean, instd = torch.zeros(1, 2, 1, 1, device=dev), torch.full((1, 2, 1, 1), 0.3, device=dev)
    outmean, outstd = [0.0, 0.0], [1e-2, 1e-2]
    losses = []

    def train_step(m):
        u = u0
        for _ in range(unroll):
            nn_out = m((u[:, :, :-1, :-1] - inmean) / instd, mun)
            u = step.step_update(u, mu, nn_out, outmean, outstd)
        loss = torch.nn.functional.mse_loss(u, target)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(loss.detach())

    t_full = timed(lambda: train_step(model), 5, 2)
    out = {"ms_per_train_step": t_full, "unroll": unroll, "batch_per_gpu": Bt, "global_batch": Bt * world,
           "traj_steps_per_s": world * Bt * unroll / (t_full * 1e-3), "params": nparams,
           "allreduce_bytes": 4 * nparams if world > 1 else 0,
           "loss_first": float(losses[0]), "loss_last": float(losses[-1])}
    if world > 1:
        def local_step():
            with model.no_sync():                                # same step, gradients NOT all-reduced
                train_step(model)
        t_local = timed(local_step, 5, 2)
        out["ms_without_allreduce"] = t_local
        out["allreduce_exposed_ms"] = t_full - t_local
    # the PDE-preserving part alone: `unroll` x (update kernel saving the coarse state + its one-launch backward)
    lib = L.load()
    muv = mu.reshape(-1).contiguous()
    nn = torch.randn(Bt, 2, F - 1, F - 1, device=dev)
    o, gu, gnn = torch.empty_like(u0), torch.empty_like(u0), torch.empty_like(nn)
    gmu = torch.empty(Bt, device=dev)
    coarse = torch.empty(Bt, 2, 48, 48, device=dev)
    om, os_ = L.f2(outmean), L.f2(outstd)

    def pde_part():
        for _ in range(unroll):
            L.check(lib.ppnn_step_update_save(C.byref(step.cfg), u0.data_ptr(), muv.data_ptr(), nn.data_ptr(), om, os_,
                                              o.data_ptr(), coarse.data_ptr(), Bt, L.stream()), "update")
            L.check(lib.ppnn_step_update_bwd_saved(C.byref(step.cfg), coarse.data_ptr(), muv.data_ptr(), target.data_ptr(), os_,
                                                   gu.data_ptr(), gmu.data_ptr(), gnn.data_ptr(), Bt, L.stream()), "bwd")
    t_pde = timed(pde_part, 5, 2)

    def net_part():
        for _ in range(unroll):
            y = net((u0[:, :, :-1, :-1] - inmean) / instd, mun)
            y.backward(nn)
        net.zero_grad(set_to_none=True)
    t_net = timed(net_part, 3, 1)
    out.update(ms_pde_kernels=t_pde, ms_network_fwd_bwd=t_net,
               share_pde_kernels=t_pde / t_full, share_network=t_net / t_full,
               limited_by="the PyTorch network (cuDNN fp32/TF32 convolutions + LayerNorm), which north_star leaves in "
                          "PyTorch; the fused step + adjoint are a few per cent of the step and the all-reduce overlaps "
                          "with the backward pass",
               note="CUDA events, max over ranks; 2 warm-up + 5 timed train steps; synthetic periodic fields")
    return out


def tier_a_leg(torch, dev, timed, fused_step):
    """pde_du three ways at the rollout batch (12) and a precompute-sized batch (2160): the closure of
    train2D.py:38-79 bound to (a) ppnn_b200's drop-in operators / rhs / mesh_convertor (Tier A: one kernel per
    reference call, the reference's own torch.cat pads), (b) the same ATen ops the reference runs on a GPU today
    (true fp32: cuDNN TF32 off), (c) the fused kernel (Tier B)."""
    import ppnn_b200.operators as ops
    import ppnn_b200.rhs as rhs
    from ppnn_b200.utility.utils import mesh_convertor
    F, C = CASE["fine"], CASE["coarse"]
    dx = CASE["length"] / (C - 1)
    dy, dx2, dy2, dt = dx, dx * dx, dx * dx, CASE["dt"]
    rhsu, rhsv = getattr(rhs, CASE["rhsu"]), getattr(rhs, CASE["rhsv"])
    d2udx2, d2udy2 = ops.d2udx2_2D(accuracy=2, device=dev), ops.d2udy2_2D(accuracy=2, device=dev)
    dudx, dudy = ops.dudx_2D(accuracy=1, device=dev), ops.dudy_2D(accuracy=1, device=dev)
    mcvter = mesh_convertor(F, C, dim=2)

    def padbcx(t):
        return torch.cat((t[:, :, -1:], t, t[:, :, :1]), dim=2)

    def padbcy(t):
        return torch.cat((t[:, :, :, -1:], t, t[:, :, :, :1]), dim=3)

    def padBC_rd(t):
        tmp = torch.cat((t, t[:, :, :, :1]), dim=3)
        return torch.cat((tmp, tmp[:, :, :1]), dim=2)

    def pde_du(u, mu):
        u1 = mcvter.down(u[:, :1])[:, :, :-1, :-1]
        v1 = mcvter.down(u[:, 1:])[:, :, :-1, :-1]
        ux, uy, vx, vy = padbcx(u1), padbcy(u1), padbcx(v1), padbcy(v1)
        return torch.cat((mcvter.up(padBC_rd(dt * rhsu(u1, v1, mu, dudx, dudy, d2udx2, d2udy2, ux, uy, dx, dy, dx2, dy2))),
                          mcvter.up(padBC_rd(dt * rhsv(u1, v1, mu, dudx, dudy, d2udx2, d2udy2, vx, vy, dx, dy, dx2, dy2)))),
                         dim=1)

    out = {"note": "ms per pde_du call; tier_a = drop-in modules behind the unchanged closure, tier_b = fused kernel"}
    with torch.no_grad():
        for B in (12, 2160):
            u = synth_fields(B, F, dev, 21)
            mu = make_mu(B, dev).reshape(-1, 1, 1, 1)
            ya, yb = pde_du(u, mu), fused_step.pde_du(u, mu)
            err = float((ya - yb).norm() / yb.norm())
            out[f"B{B}"] = {"tier_a_ms": timed(lambda: pde_du(u, mu), 10, 3), "tier_b_ms": timed(lambda: fused_step.pde_du(u, mu), 10, 3),
                            "tier_a_vs_tier_b_rel_l2": err}
            del u, ya, yb
    return out


def datagen_leg(torch, dist, dev, world, rank, local, max_over_ranks, barrier, link=None):
    """BASELINE configs[3]: the fine-grid Burgers generator (operators.py:403-492) at the reference's LENGTH
    (20 001 steps, snapshot every 200 => 101 snapshots), 1025^2 (the reference's duplicated-endpoint convention at
    config 4's resolution), 8 viscosities x 8 initial conditions = 64 samples PER GPU, sharded with no communication.
    End to end including set-up (initial conditions, buffers, graph capture): wall clock, max over ranks."""
    from ppnn_b200.datagen import generate
    out = {}
    for tag, kw in (("device_sink", dict(on_device=True)), ("host_sink_discard", dict(keep=False))):
        barrier()
        t0 = time.perf_counter()
        tm = {}
        data = generate(mesh=1025, para=8, repeat=8 * world, steps=20001, every=200, device=dev, rank=rank,
                        world_size=world, timings=tm, **kw)
        torch.cuda.synchronize()
        sec = max_over_ranks(time.perf_counter() - t0)
        finite = bool(torch.isfinite(data[:, -1]).all()) if data.numel() else None
        del data
        P = 64 * world
        out[tag] = {"seconds": sec, "samples_per_s": P / sec, "setup_seconds": max_over_ranks(tm["setup_seconds"]),
                    "solve_seconds": max_over_ranks(tm["solve_seconds"]),
                    "gpoint_updates_per_s": P * 2 * 1025 * 1025 * 20001 / sec / 1e9,
                    "snapshot_gb_per_gpu": 101 * 64 * 2 * 1025 * 1025 * 4 / 1e9, "finite": finite}
        torch.cuda.empty_cache()
    if link:
        # the host sink moves every snapshot over the host link while the solver runs: its floor is the slower of the
        # solver alone (the device-sink solve time is an upper bound of it) and the D2H at the measured link rate
        h = out["host_sink_discard"]
        gb = h["snapshot_gb_per_gpu"]
        h["d2h_gbs_per_gpu"] = gb / h["solve_seconds"]
        h["d2h_ceiling_gbs_per_gpu"] = link["d2h_gbs_per_gpu"]
        link_floor = gb / link["d2h_gbs_per_gpu"]
        h["link_floor_seconds"] = link_floor
        h["bound"] = "host link (D2H)" if link_floor > 0.9 * out["device_sink"]["solve_seconds"] else "solver (the D2H is hidden behind it)"
        h["frac_of_floor"] = max(link_floor, min(out["device_sink"]["solve_seconds"], h["solve_seconds"])) / h["solve_seconds"]
    out["note"] = ("device_sink: the 101 snapshots stay in HBM (54 GB per GPU; the form PdeDataset trains from); "
                   "host_sink_discard: every snapshot is copied to pinned host memory on a side stream and dropped "
                   "(the box cannot hold 8 x 54 GB): solver + the D2H the reference does at operators.py:479-482")
    return out


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from ppnn_b200 import CoarseStep, FineBurgers, HostStepPipeline, _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from ppnn_b200.dist import bind_to_gpu_numa
    numa_cores = bind_to_gpu_numa(local)              # pinned staging buffers next to this GPU's PCIe root
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.load()

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    F, B = CASE["fine"], args.batch
    step = CoarseStep(F, CASE["coarse"], CASE["rhsu"], CASE["rhsv"], CASE["dt"], CASE["length"])
    u = synth_fields(B, F, dev, 10 + rank)
    mu = make_mu(B, dev, offset=rank * B)
    muv = mu.reshape(-1).contiguous()
    nn = torch.randn(B, 2, F - 1, F - 1, device=dev, generator=torch.Generator(device=dev).manual_seed(11 + rank))
    out = torch.empty_like(u)
    om, os_ = L.f2(STATS["outmean"]), L.f2(STATS["outstd"])
    pts = B * 2 * F * F

    def kernel_step():
        step.step_update_(u, muv, nn, om, os_, out)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps

    sampler = ClockSampler(local) if rank == 0 else None

    # ---- value: resident inputs, one launch per step ---------------------------------
    ms = timed(kernel_step, args.steps, max(args.warmup, 3))
    value = world * pts / (ms * 1e-3) / 1e9

    # ---- parity of what was just timed (outside the timed region): 4 samples spread over the batch,
    # i.e. over first / middle / last CTA waves, recomputed by the CPU reference
    parity = None
    if rank == 0:
        idx = sorted({0, B // 3, (2 * B) // 3, B - 1})
        sel = torch.tensor(idx, device=dev)
        update, pkind, _ = cpu_update_fn()
        ref = update(u[sel].cpu(), mu[sel].cpu(), nn[sel].cpu(), torch.tensor(STATS["outmean"]).reshape(1, 2, 1, 1),
                     torch.tensor(STATS["outstd"]).reshape(1, 2, 1, 1)).double()
        got = out[sel].cpu().double()
        parity = {"rel_l2": ((got - ref).norm() / ref.norm()).item(),
                  "rel_max": ((got - ref).abs().max() / ref.abs().max()).item(),
                  "samples": len(idx), "sample_indices": idx, "against": pkind, "tolerance": 1e-5}
        parity["ok"] = bool(parity["rel_l2"] <= 1e-5 and parity["rel_max"] <= 1e-5)

    # ---- host link ceiling: plain pinned-memory copies, every rank at once (one cudaMemcpyAsync per call) ----
    def link_ceiling():
        n = 64 << 20                                        # 256 MB per copy
        h_in = torch.empty(n, dtype=torch.float32, pin_memory=True).fill_(1.0)
        h_out = torch.empty(n, dtype=torch.float32, pin_memory=True)
        d_in, d_out = torch.empty(n, device=dev), torch.ones(n, device=dev)
        s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        reps, nbytes = 6, n * 4

        def run(h2d, d2h):
            """`h2d` / `d2h` copies per repetition on their own streams, concurrently"""
            cur = torch.cuda.current_stream()
            s1.wait_stream(cur); s2.wait_stream(cur)
            for _ in range(reps):
                with torch.cuda.stream(s1):
                    for _k in range(h2d):
                        d_in.copy_(h_in, non_blocking=True)
                with torch.cuda.stream(s2):
                    for _k in range(d2h):
                        h_out.copy_(d_out, non_blocking=True)
            cur.wait_stream(s1); cur.wait_stream(s2)
        res = {}
        # "mix" = the traffic pattern of the e2e step itself: two bytes in (u, nn_out) for every byte out (u_next)
        for tag, h2d, d2h in (("h2d", 1, 0), ("d2h", 0, 1), ("bidir", 1, 1), ("mix_2to1", 2, 1)):
            # ms per run, max over ranks; best of 3 trials (a ceiling, so the fastest counts)
            t = min(timed(lambda: run(h2d, d2h), 2, 1) for _ in range(3))
            per_dir = reps * nbytes / t / 1e6
            res[tag + "_gbs_per_gpu"] = per_dir * (h2d + d2h) if tag == "mix_2to1" else per_dir
        res["aggregate_bidir_gbs"] = world * 2 * res["bidir_gbs_per_gpu"]
        res["aggregate_mix_gbs"] = world * res["mix_2to1_gbs_per_gpu"]
        res["how"] = ("256 MB pinned<->device copies, 6 repetitions, all ranks at once, CUDA events, max over ranks, best of 3 "
                      "trials; h2d / d2h: one direction alone; bidir: one copy each way at once (per-direction rate); "
                      "mix_2to1: two copies in for one copy out, the e2e step's own traffic (total rate, both directions)")
        return res
    link = link_ceiling()

    # ---- e2e: pinned host buffers -> HostStepPipeline (H2D / kernel / D2H on three streams) ----
    nchunk = 16
    u_h = torch.empty(u.shape, dtype=torch.float32, pin_memory=True).copy_(u)
    nn_h = torch.empty(nn.shape, dtype=torch.float32, pin_memory=True).copy_(nn)
    out_h = torch.empty(u.shape, dtype=torch.float32, pin_memory=True)
    pipe = HostStepPipeline(step, B, chunks=nchunk, depth=3, device=dev)
    dbuf = pipe.bufs

    def e2e_step():
        pipe.run(u_h, muv, nn_h, om, os_, out_h)

    e2e_steps = max(2, min(args.steps, 10))
    ms_e2e = timed(e2e_step, e2e_steps, 1)
    e2e_value = world * pts / (ms_e2e * 1e-3) / 1e9
    torch.cuda.synchronize()
    e2e_ok = bool(torch.equal(out_h, out.cpu()))             # the WHOLE batch: every chunk incl. the ragged last one
    h2d = (u.numel() + nn.numel()) * 4
    d2h = u.numel() * 4
    # the step cannot beat either direction at its own measured rate, nor its total traffic at the rate measured for the same 2:1 mix
    e2e_floor_ms = max(h2d / link["h2d_gbs_per_gpu"], d2h / link["d2h_gbs_per_gpu"],
                       (h2d + d2h) / link["mix_2to1_gbs_per_gpu"]) / 1e6

    # ---- secondary kernels (same case), device resident ------------------------------
    extras = {}
    if not args.no_extras:
        reps = max(3, min(args.steps, 10))
        stats = dict(inmean=L.f2([0.0, 0.0]), instd=L.f2([0.3, 0.3]), pdemean=L.f2([0.0, 0.0]), pdestd=L.f2([1e-3, 1e-3]))
        t = timed(lambda: step.step_inputs(u, muv, stats["inmean"], stats["instd"], stats["pdemean"], stats["pdestd"]), reps, 2)
        extras["step_inputs_K1"] = {"ms": t, "gpts": world * pts / t / 1e6, "bytes_per_pt": 12,
                                    "hbm_gbs": 12 * pts / t / 1e6}
        t = timed(lambda: step.step_update_(u, muv, None, None, None, out), reps, 2)
        extras["pde_only_update"] = {"ms": t, "gpts": world * pts / t / 1e6, "bytes_per_pt": 8, "hbm_gbs": 8 * pts / t / 1e6}
        lib = L.load()
        import ctypes as C
        gu = torch.empty_like(u)
        gmu = torch.empty(B, device=dev)

        def adj():
            L.check(lib.ppnn_pde_du_bwd(C.byref(step.cfg), u.data_ptr(), muv.data_ptr(), out.data_ptr(), gu.data_ptr(),
                                        gmu.data_ptr(), B, 1, L.stream()), "bwd")
        t = timed(adj, reps, 2)
        extras["step_adjoint"] = {"ms": t, "gpts": world * pts / t / 1e6, "bytes_per_pt": 8, "hbm_gbs": 8 * pts / t / 1e6}
        # backward of the whole update in one launch: grad_u, grad_mu and the network-output gradient
        gnn = torch.empty_like(nn)

        def upd_bwd():
            L.check(lib.ppnn_step_update_bwd(C.byref(step.cfg), u.data_ptr(), muv.data_ptr(), out.data_ptr(), os_,
                                             gu.data_ptr(), gmu.data_ptr(), gnn.data_ptr(), B, L.stream()), "update bwd")
        t = timed(upd_bwd, reps, 2)
        extras["step_update_bwd"] = {"ms": t, "gpts": world * pts / t / 1e6, "bytes_per_pt": 12, "hbm_gbs": 12 * pts / t / 1e6}
        # the same two from the coarse state the forward kernel saved (what autograd through step_update does)
        coarse = torch.empty(B, 2, CASE["coarse"] - 1, CASE["coarse"] - 1, device=dev)
        L.check(lib.ppnn_step_update_save(C.byref(step.cfg), u.data_ptr(), muv.data_ptr(), nn.data_ptr(), om, os_,
                                          gu.data_ptr(), coarse.data_ptr(), B, L.stream()), "update (saving)")
        t = timed(lambda: L.check(lib.ppnn_step_update_save(C.byref(step.cfg), u.data_ptr(), muv.data_ptr(), nn.data_ptr(), om, os_,
                                                            gu.data_ptr(), coarse.data_ptr(), B, L.stream()), "update (saving)"), reps, 2)
        extras["step_update_saving_coarse_state"] = {"ms": t, "gpts": world * pts / t / 1e6, "bytes_per_pt": 12,
                                                     "hbm_gbs": 12 * pts / t / 1e6}
        t = timed(lambda: L.check(lib.ppnn_pde_du_bwd_saved(C.byref(step.cfg), coarse.data_ptr(), muv.data_ptr(), out.data_ptr(),
                                                            gu.data_ptr(), gmu.data_ptr(), B, 1, L.stream()), "bwd saved"), reps, 2)
        extras["step_adjoint_saved_state"] = {"ms": t, "gpts": world * pts / t / 1e6, "bytes_per_pt": 8, "hbm_gbs": 8 * pts / t / 1e6}
        t = timed(lambda: L.check(lib.ppnn_step_update_bwd_saved(C.byref(step.cfg), coarse.data_ptr(), muv.data_ptr(), out.data_ptr(),
                                                                 os_, gu.data_ptr(), gmu.data_ptr(), gnn.data_ptr(), B, L.stream()),
                                  "update bwd saved"), reps, 2)
        extras["step_update_bwd_saved_state"] = {"ms": t, "gpts": world * pts / t / 1e6, "bytes_per_pt": 12,
                                                 "hbm_gbs": 12 * pts / t / 1e6}
        del gnn, coarse
        # training-set precompute (train2D.py:109-151): statistics pass and emit pass over (u, u_next) pairs
        part = torch.empty(B, 2, 6, dtype=torch.float64, device=dev)
        t = timed(lambda: L.check(lib.ppnn_dataset_stats(C.byref(step.cfg), u.data_ptr(), out.data_ptr(), muv.data_ptr(),
                                                         B, 0, part.data_ptr(), L.stream()), "stats"), reps, 2)
        extras["dataset_stats_pass"] = {"ms": t, "gpts": world * pts / t / 1e6, "bytes_per_pt": 8, "hbm_gbs": 8 * pts / t / 1e6}
        m6, s6 = (C.c_float * 6)(0, 0, 0, 0, 0, 0), (C.c_float * 6)(0.3, 0.3, 1e-3, 1e-3, 1e-3, 1e-3)
        e1, e2 = torch.empty_like(nn), torch.empty_like(nn)
        t = timed(lambda: L.check(lib.ppnn_dataset_emit(C.byref(step.cfg), u.data_ptr(), out.data_ptr(), muv.data_ptr(), B, 0,
                                                        m6, s6, gu.data_ptr(), e1.data_ptr(), e2.data_ptr(), L.stream()),
                                  "emit"), reps, 2)
        extras["dataset_emit_pass"] = {"ms": t, "gpts": world * pts / t / 1e6, "bytes_per_pt": 20, "hbm_gbs": 20 * pts / t / 1e6}
        del gu, u_h, nn_h, out_h, dbuf, part, e1, e2
        # datagen (BASELINE config 4): fused Burgers step, 64 samples x 1025^2 per GPU
        fb = FineBurgers()
        n, bf = 1025, 64
        fu = synth_fields(bf, n, dev, 99)            # [bf,2,n,n]
        a, b_ = fu[:, 0].contiguous(), fu[:, 1].contiguous()
        a2, b2 = torch.empty_like(a), torch.empty_like(b_)
        mf = make_mu(bf, dev).reshape(-1) * 0.05
        fpts = bf * 2 * n * n
        t = timed(lambda: fb.step_(a, b_, mf, a2, b2), reps * 2, 3)
        extras["fine_burgers_step_1025x64"] = {"ms": t, "gpts": world * fpts / t / 1e6, "bytes_per_pt": 8,
                                               "hbm_gbs": 8 * fpts / t / 1e6}
        # the same on all-positive fields like the generator's own (initial conditions in [0.1, 1.1], operators.py:436-445)
        ap, bp = a + 0.6, b_ + 0.6
        t = timed(lambda: fb.step_(ap, bp, mf, a2, b2), reps * 2, 3)
        extras["fine_burgers_step_1025x64_positive_fields"] = {"ms": t, "gpts": world * fpts / t / 1e6, "bytes_per_pt": 8,
                                                               "hbm_gbs": 8 * fpts / t / 1e6}
        del ap, bp
        fbi = FineBurgers(mode="intended")
        t = timed(lambda: fbi.step_(a, b_, mf, a2, b2), reps * 2, 3)
        extras["fine_burgers_step_intended_1025x64"] = {"ms": t, "gpts": world * fpts / t / 1e6, "bytes_per_pt": 8,
                                                        "hbm_gbs": 8 * fpts / t / 1e6}

        del fu, a, b_, a2, b2
        n = 1024                                     # BASELINE config 4 as named: 1024^2 true periodic
        fu = synth_fields(bf, n, dev, 98)
        a, b_ = fu[:, 0].contiguous(), fu[:, 1].contiguous()
        a2, b2 = torch.empty_like(a), torch.empty_like(b_)
        fbp = FineBurgers(endpoint=False)
        fpts = bf * 2 * n * n
        t = timed(lambda: fbp.step_(a, b_, mf, a2, b2), reps * 2, 3)
        extras["fine_burgers_step_1024x64_true_periodic"] = {"ms": t, "gpts": world * fpts / t / 1e6, "bytes_per_pt": 8,
                                                             "hbm_gbs": 8 * fpts / t / 1e6}
        del fu, a, b_, a2, b2

        # small explicit solvers with the state on chip: all steps in ONE launch
        try:
            import math
            from ppnn_b200 import KS1D, KS2D, HeatBC1D
            xs = torch.linspace(0, 8 * math.pi, 129, device=dev).reshape(1, 1, -1)
            ku = torch.cos(xs / 16 + torch.randn(256, 1, 1, device=dev)) * (1 + torch.sin(xs / 16 + torch.randn(256, 1, 1, device=dev)))
            ks1 = KS1D(dt=1e-4, dx=math.pi / 8, source=0.1 * torch.exp(-(xs - 4 * math.pi) ** 2 / 2))
            t = timed(lambda: ks1.run(ku, 20000), 2, 1)
            extras["ks1d_rk4_256x129"] = {"us_per_step": t * 1e3 / 20000, "steps_per_launch": 20000,
                                          "gpts": world * 256 * 129 * 20000 / t / 1e6}
            x2 = torch.linspace(0, 8 * math.pi, 65, device=dev)
            kw = (torch.cos(x2.reshape(-1, 1) / 4) * (1 + torch.sin(x2.reshape(1, -1) / 4))).reshape(1, 1, 65, 65)
            kw = (kw * torch.linspace(0.9, 1.1, 296, device=dev).reshape(-1, 1, 1, 1)).contiguous()
            ks2 = KS2D(dt=1e-3, dx=math.pi / 8)
            t = timed(lambda: ks2.run(kw, 1000), 2, 1)
            extras["ks2d_rk4_296x65x65"] = {"us_per_step": t * 1e3 / 1000, "steps_per_launch": 1000,
                                            "gpts": world * 296 * 65 * 65 * 1000 / t / 1e6}
            hb = torch.randn(256, 1, 126, device=dev).cumsum(-1) * 0.1
            hc, hv = torch.randint(0, 2, (256, 1, 2), device=dev), torch.rand(256, 1, 2, device=dev) * 2 - 1
            heat = HeatBC1D(0.01, 16 * math.pi / 127)
            t = timed(lambda: heat.generate(hb, hc, hv, 10001, 50), 2, 1)
            extras["bv_heat_256x126_10001steps"] = {"us_per_step": t * 1e3 / 10001, "launches": 201,
                                                    "gpts": world * 256 * 126 * 10001 / t / 1e6}
            del ku, kw, hb
        except Exception as exc:
            extras["onchip_solvers_error"] = repr(exc)[:300]

        # Tier A: the pde_du closure of train2D.py:38-79 re-typed around the drop-in modules -- what an UNCHANGED
        # reference script executes -- next to the same ATen op sequence on the GPU and the fused kernel
        try:
            extras["tier_a_pde_du_chain_RD_pPDE"] = tier_a_leg(torch, dev, timed, step)
        except Exception as exc:
            extras["tier_a_error"] = repr(exc)[:300]

        try:
            if not args.no_legs:
                extras["train_unrolled_burgers_fPDE"] = train_leg(torch, dist, dev, world, rank, local, timed, L)
        except Exception as exc:
            extras["train_unrolled_error"] = repr(exc)[:300]
        try:
            if not args.no_legs:
                extras["datagen_1025x64_per_gpu"] = datagen_leg(torch, dist, dev, world, rank, local, max_over_ranks, barrier, link)
        except Exception as exc:
            extras["datagen_error"] = repr(exc)[:300]

        # rollout (train2D.py:230-245): K1 -> network -> K2 per step, CUDA-graph replayed,
        # around a random-weight network with the layer shapes of cnn2dRich (ppnn_b200/nets.py)
        try:
            if args.no_legs:
                raise StopIteration
            from ppnn_b200.nets import ConvResNet
            torch.manual_seed(1)
            net = ConvResNet(True).to(dev).eval()
            rstats = dict(inmean=torch.zeros(1, 2, 1, 1), instd=torch.full((1, 2, 1, 1), 0.3),
                          pdemean=torch.zeros(1, 2, 1, 1), pdestd=torch.full((1, 2, 1, 1), 1e-3),
                          outmean=torch.tensor(STATS["outmean"]).reshape(1, 2, 1, 1),
                          outstd=torch.tensor(STATS["outstd"]).reshape(1, 2, 1, 1),
                          mumean=torch.tensor(0.75, device=dev), mustd=torch.tensor(0.35, device=dev))
            T = 100
            for nb, tag in ((CASE["num_para"], "rollout_100steps_B12"), (96, "rollout_100steps_B96")):
                u0 = u[:nb].clone()
                m0 = mu[:nb].clone()
                plan = step.prepare_rollout(net, nb, m0, rstats, uses_pde=True, graph=True, device=dev)   # capture: once
                res = torch.empty(nb, T, 2, F, F, device=dev)
                plan.run(u0, 3, out=res[:, :3])                                       # warm-up replays
                barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                plan.run(u0, T, out=res)
                torch.cuda.synchronize()
                dt_roll = time.perf_counter() - t0
                extras[tag] = {"seconds": dt_roll, "traj_per_s": world * nb / dt_roll, "steps": T,
                               "finite": bool(torch.isfinite(res).all()),
                               "note": "replay of a captured step graph (RolloutPlan.run): K1 + ConvResNet(1.55M params, "
                                       "fp32/cudnn) + K2 + snapshot copy per step"}
                del plan
                del res
        except StopIteration:
            pass
        except Exception as exc:   # the rollout extra must never take the headline down
            extras["rollout_error"] = repr(exc)[:300]

    clocks = sampler.stop() if sampler else None

    if world > 1:
        dist.barrier(device_ids=[local])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = None, "fallback"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        peak = 6650.0
    achieved = algorithmic_bytes(B) / (ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("coarse_step_update_bytes_per_launch")
    except Exception:
        pass
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(B, world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e, "steps": e2e_steps, "matches_resident_result": e2e_ok,
                    "link_ceiling": link, "link_floor_ms_per_step": e2e_floor_ms,
                    "frac_of_link_ceiling": e2e_floor_ms / ms_e2e, "numa_bound_cores": len(numa_cores) if numa_cores else None,
                    "api": "ppnn_b200.HostStepPipeline.run -> CoarseStep.step_update_ -> ppnn_step_update (C ABI), pinned "
                           "host buffers, 16 chunks through separate H2D / kernel / D2H streams (ring of 3 device buffer sets)"},
            "gpu_launches": args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": "static: ncu --set full capture of this kernel at this shape "
                                                               "(profiles/traffic.json), not measured in this run",
                         "peak_source": peak_src, "kernel": "coarse_step_staged_kernel<MODE_UPDATE, 257, HAS_NN>",
                         "algorithmic_bytes_per_launch": algorithmic_bytes(B),
                         "bytes_per_point_update": algorithmic_bytes(B) / pts},
            "parity": parity,
            "extras": extras}
    if world == 1 and not args.no_cpu:
        v, reps, dt, kind, desc = time_cpu(args.cpu_batch)
        import torch as _t
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": _t.get_num_threads(), "kind": kind,
                                "sample": f"{args.cpu_batch} of {B} samples x {reps} reps ({dt:.2f} s each), "
                                          f"{desc}, on all host threads"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.sweep:
        run_sweep(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
