"""Tier B: the fully fused PDE-preserving path.

``CoarseStep`` is the ``pde_du`` closure that seven reference scripts re-type
(src/train2D.py:49-79 and the post_process/test_*.py copies) as one object whose
every method is one kernel launch:

* ``pde_du(u, mu)``            train2D.py:66-79 -- differentiable (adjoint kernel)
* ``step_inputs(...)``         train2D.py:235-237 -- normalised network inputs
* ``step_update(...)``         train2D.py:235-238 -- padBC_rd(net*std+mean) + u + pde_du(u)
* ``pde_only_rollout(...)``    post_process/test_predict_2Dp.py:116
* ``rollout(net, ...)``        train2D.py:230-245 (optionally CUDA-graph captured)
* ``coarse_increment(...)``    post_process/test_2D_csolver.py:56-66

``FineBurgers`` is the explicit fine-grid generator loop of
src/operators.py:466-484 with one fused kernel per time step.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from .rhs import RHS_IDS


def _mu_vec(mu, B: int, device) -> torch.Tensor:
    if isinstance(mu, torch.Tensor):
        m = mu.to(device=device, dtype=torch.float32).reshape(-1)
        if m.numel() == 1 and B != 1:
            m = m.expand(B)
        if m.numel() != B:
            raise RuntimeError(f"mu has {m.numel()} entries for a batch of {B}")
        return m.contiguous()
    return torch.full((B,), float(mu), dtype=torch.float32, device=device)


class _PdeDuFn(torch.autograd.Function):
    """pde_du with the hand-written adjoint.  When a gradient will be needed the forward kernel also
    stores the restricted coarse state (18 KB per sample on the 257/49 mesh) and the backward pass
    runs from that (``ppnn_pde_du_bwd_saved``): pde_du depends on ``u`` only through it, so neither is
    the 528 KB state kept alive for the backward pass nor re-gathered from HBM by the adjoint kernel."""

    @staticmethod
    def forward(ctx, u, mu, step: "CoarseStep"):
        lib = L.load()
        x = L.dev(u, "pde_du input")
        B = x.shape[0]
        muv = _mu_vec(mu, B, x.device)
        out = torch.empty_like(x)
        want = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        coarse = torch.empty(B, 2, step.CH - 1, step.CW - 1, dtype=torch.float32, device=x.device) if want else None
        with torch.cuda.device(x.device):
            L.check(lib.ppnn_pde_du_fwd_save(C.byref(step.cfg), x.data_ptr(), muv.data_ptr(), out.data_ptr(),
                                             L.ptr(coarse), B, L.stream()), "ppnn_pde_du_fwd")
        ctx.step = step
        ctx.mu_shape = mu.shape if isinstance(mu, torch.Tensor) else None
        ctx.state_shape = x.shape
        if want:
            ctx.save_for_backward(coarse, muv)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = L.load()
        coarse, muv = ctx.saved_tensors
        g = L.dev(g, "pde_du grad")
        B = g.shape[0]
        gu = torch.empty(ctx.state_shape, dtype=torch.float32, device=g.device)
        want_mu = ctx.mu_shape is not None and ctx.needs_input_grad[1]
        gmu = torch.empty(B, dtype=torch.float32, device=g.device) if want_mu else None
        with torch.cuda.device(g.device):
            L.check(lib.ppnn_pde_du_bwd_saved(C.byref(ctx.step.cfg), coarse.data_ptr(), muv.data_ptr(), g.data_ptr(),
                                              gu.data_ptr(), L.ptr(gmu), B, 0, L.stream()), "ppnn_pde_du_bwd")
        if want_mu:
            gmu = gmu.sum().reshape(ctx.mu_shape) if ctx.mu_shape.numel() == 1 and B != 1 else gmu.reshape(ctx.mu_shape)
        return gu, gmu, None


class _StepUpdateFn(torch.autograd.Function):
    """u_next = padBC_rd(nn*outstd + outmean) + u + pde_du(u, mu), one kernel each way; the backward
    pass runs from the coarse state saved by the forward kernel (see ``_PdeDuFn``)."""

    @staticmethod
    def forward(ctx, u, mu, nn_out, step: "CoarseStep", outmean, outstd):
        lib = L.load()
        x = L.dev(u, "step_update state")
        B = x.shape[0]
        muv = _mu_vec(mu, B, x.device)
        n = L.dev(nn_out, "network output") if nn_out is not None else None
        if n is not None and tuple(n.shape) != (B, 2, step.FH - 1, step.FW - 1):
            raise RuntimeError(f"network output has shape {tuple(n.shape)}, expected {(B, 2, step.FH - 1, step.FW - 1)}")
        out = torch.empty_like(x)
        m, s = (L.f2(outmean), L.f2(outstd)) if n is not None else (None, None)
        want = any(ctx.needs_input_grad[:3])
        coarse = torch.empty(B, 2, step.CH - 1, step.CW - 1, dtype=torch.float32, device=x.device) if want else None
        with torch.cuda.device(x.device):
            L.check(lib.ppnn_step_update_save(C.byref(step.cfg), x.data_ptr(), muv.data_ptr(), L.ptr(n), m, s,
                                              out.data_ptr(), L.ptr(coarse), B, L.stream()), "ppnn_step_update")
        ctx.step = step
        ctx.mu_shape = mu.shape if isinstance(mu, torch.Tensor) else None
        ctx.has_nn = n is not None
        ctx.outstd = [float(v) for v in s] if n is not None else None
        ctx.state_shape = x.shape
        if want:
            ctx.save_for_backward(coarse, muv)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = L.load()
        coarse, muv = ctx.saved_tensors
        g = L.dev(g, "step_update grad")
        B = g.shape[0]
        gu = gmu = gnn = None
        want_nn = ctx.has_nn and ctx.needs_input_grad[2]
        want_mu = ctx.mu_shape is not None and ctx.needs_input_grad[1]
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1] or want_nn:
            gu = torch.empty(ctx.state_shape, dtype=torch.float32, device=g.device)
            gmu = torch.empty(B, dtype=torch.float32, device=g.device) if want_mu else None
            st = ctx.step
            gnn = torch.empty(B, 2, st.FH - 1, st.FW - 1, dtype=torch.float32, device=g.device) if want_nn else None
            with torch.cuda.device(g.device):
                # one launch: grad_u, grad_mu and the gradient of the network output
                status = lib.ppnn_step_update_bwd_saved(C.byref(st.cfg), coarse.data_ptr(), muv.data_ptr(), g.data_ptr(),
                                                        (C.c_float * 2)(*ctx.outstd) if want_nn else None,
                                                        gu.data_ptr(), L.ptr(gmu), L.ptr(gnn), B, L.stream())
                if status == L.E_SHAPE and want_nn:
                    # meshes outside the reference family: adjoint kernel + padBC_rd^T composed on the device
                    L.check(lib.ppnn_pde_du_bwd_saved(C.byref(st.cfg), coarse.data_ptr(), muv.data_ptr(), g.data_ptr(),
                                                      gu.data_ptr(), L.ptr(gmu), B, 1, L.stream()), "ppnn_pde_du_bwd")
                    gnn = g[:, :, :-1, :-1].clone()
                    gnn[:, :, 0, :] += g[:, :, -1, :-1]
                    gnn[:, :, :, 0] += g[:, :, :-1, -1]
                    gnn[:, :, 0, 0] += g[:, :, -1, -1]
                    gnn = gnn * torch.tensor(ctx.outstd, device=g.device).reshape(1, 2, 1, 1)
                else:
                    L.check(status, "ppnn_step_update_bwd")
            if want_mu:
                gmu = gmu.sum().reshape(ctx.mu_shape) if ctx.mu_shape.numel() == 1 and B != 1 else gmu.reshape(ctx.mu_shape)
        return gu, gmu, gnn, None, None, None


class CoarseStep:
    """The PDE-preserving coarse step of one case (train2D.py:49-64 builds the same
    quantities from the YAML): ``fine`` / ``coarse`` are mesh sizes INCLUDING the
    duplicated periodic endpoint, ``dx = length / (coarse - 1)`` and ``dy = dx``."""

    def __init__(self, fine, coarse, rhsu: str, rhsv: str, dt: float, length: float,
                 adv_accuracy: int = 1, diff_accuracy: int = 2,
                 dx: Optional[float] = None, dy: Optional[float] = None):
        fh, fw = (fine, fine) if isinstance(fine, int) else tuple(fine)
        ch, cw = (coarse, coarse) if isinstance(coarse, int) else tuple(coarse)
        self.FH, self.FW, self.CH, self.CW = int(fh), int(fw), int(ch), int(cw)
        self.rhsu, self.rhsv = rhsu, rhsv
        self.dt = float(dt)
        self.dx = float(dx) if dx is not None else length / (self.CH - 1)      # train2D.py:58
        self.dy = float(dy) if dy is not None else self.dx                     # train2D.py:61
        self.dx2, self.dy2 = self.dx ** 2, self.dy ** 2
        self.cfg = L.CoarseCfg(self.FH, self.FW, self.CH, self.CW, RHS_IDS[rhsu], RHS_IDS[rhsv],
                               int(adv_accuracy), int(diff_accuracy),
                               self.dt, self.dx, self.dy, self.dx2, self.dy2)

    @classmethod
    def from_params(cls, params) -> "CoarseStep":
        """From the SimpleNamespace train2D.py:19 builds out of a cases/*.yaml file."""
        return cls(params.finemeshsize, params.coarsemeshsize, params.rhsu, params.rhsv,
                   params.dt, params.length)

    # ------------------------------------------------------------------ kernels
    def _check_state(self, u):
        if u.dim() != 4 or u.shape[1] != 2 or u.shape[2] != self.FH or u.shape[3] != self.FW:
            raise RuntimeError(f"state has shape {tuple(u.shape)}, expected [B,2,{self.FH},{self.FW}]")

    def pde_du(self, u: torch.Tensor, mu) -> torch.Tensor:
        """[B,2,F,F], mu [B,1,1,1] -> [B,2,F,F]; autograd flows to ``u`` and ``mu``."""
        self._check_state(u)
        return _PdeDuFn.apply(u, mu, self)

    __call__ = pde_du

    @torch.no_grad()
    def pde_du_iter(self, u: torch.Tensor, mu, inner: int = 1) -> torch.Tensor:
        """The ``pde_du`` of post_process/test2D2.py:70-91: ``inner`` sequential explicit sub-steps of the restricted
        coarse state (v sees the updated u) before the increment is evaluated and prolonged.  Inference only."""
        self._check_state(u)
        u = L.dev(u, "u")
        muv = _mu_vec(mu, u.shape[0], u.device)
        out = torch.empty_like(u)
        with L.on_device(u):
            L.check(L.load().ppnn_pde_du_iter_fwd(C.byref(self.cfg), u.data_ptr(), muv.data_ptr(), out.data_ptr(),
                                                  u.shape[0], int(inner), L.stream()), "ppnn_pde_du_iter_fwd")
        return out

    def step_inputs(self, u, mu, inmean, instd, pdemean=None, pdestd=None, want_pdeu=False):
        """Normalised network inputs of train2D.py:235-237 (no autograd):
        ``(u_norm, pdeu_norm | None, pdeu | None)``.  Statistics may be tensors
        ([1,2,1,1]), sequences, or pre-converted ``float[2]`` arrays (``_lib.f2``)."""
        self._check_state(u)
        lib = L.load()
        x = L.dev(u, "state")
        B = x.shape[0]
        muv = _mu_vec(mu, B, x.device)
        shp = (B, 2, self.FH - 1, self.FW - 1)
        un = torch.empty(shp, dtype=torch.float32, device=x.device)
        pn = torch.empty(shp, dtype=torch.float32, device=x.device) if pdemean is not None else None
        pd = torch.empty_like(x) if want_pdeu else None
        as2 = lambda a: a if isinstance(a, C.Array) else L.f2(a)
        with torch.cuda.device(x.device):
            L.check(lib.ppnn_step_inputs(C.byref(self.cfg), x.data_ptr(), muv.data_ptr(),
                                         as2(inmean), as2(instd),
                                         as2(pdemean) if pn is not None else None,
                                         as2(pdestd) if pn is not None else None,
                                         un.data_ptr(), L.ptr(pn), L.ptr(pd), B, L.stream()), "ppnn_step_inputs")
        return un, pn, pd

    def step_update(self, u, mu, nn_out=None, outmean=None, outstd=None) -> torch.Tensor:
        """``padBC_rd(nn_out*outstd + outmean) + u + pde_du(u, mu)``; with
        ``nn_out=None`` the PDE-only update ``u + pde_du(u, mu)``.  Differentiable."""
        self._check_state(u)
        return _StepUpdateFn.apply(u, mu, nn_out, self, outmean, outstd)

    def step_update_(self, u, mu, nn_out, outmean2, outstd2, out) -> torch.Tensor:
        """Allocation-free, autograd-free update into ``out`` (must not overlap ``u``); the
        statistics are pre-converted ``float[2]`` arrays (``_lib.f2``) -- for graphs."""
        lib = L.load()
        B = u.shape[0]
        L.check(lib.ppnn_step_update(C.byref(self.cfg), u.data_ptr(), mu.data_ptr(), L.ptr(nn_out),
                                     outmean2, outstd2, out.data_ptr(), B, L.stream()), "ppnn_step_update")
        return out

    def coarse_increment(self, uc, mu, add_state: bool = False) -> torch.Tensor:
        """Coarse-only solver (test_2D_csolver.py:56-66): uc [B,2,C,C]."""
        lib = L.load()
        x = L.dev(uc, "coarse state")
        if x.dim() != 4 or x.shape[1:] != (2, self.CH, self.CW):
            raise RuntimeError(f"coarse state has shape {tuple(x.shape)}, expected [B,2,{self.CH},{self.CW}]")
        B = x.shape[0]
        muv = _mu_vec(mu, B, x.device)
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            L.check(lib.ppnn_coarse_increment(C.byref(self.cfg), x.data_ptr(), muv.data_ptr(), out.data_ptr(),
                                              B, int(add_state), L.stream()), "ppnn_coarse_increment")
        return out

    # ------------------------------------------------------------------ rollouts
    @torch.no_grad()
    def pde_only_rollout(self, u, mu, nsteps: int, keep: bool = True):
        """``u <- u + pde_du(u, mu)`` (test_predict_2Dp.py:116); ping-pong buffers."""
        self._check_state(u)
        x = L.dev(u, "state").clone()
        muv = _mu_vec(mu, x.shape[0], x.device)
        y = torch.empty_like(x)
        out = []
        with torch.cuda.device(x.device):
            for _ in range(nsteps):
                self.step_update_(x, muv, None, None, None, y)
                x, y = y, x
                if keep:
                    out.append(x.clone())
        return out if keep else x

    @torch.no_grad()
    def rollout(self, net: Callable, u, mu, stats: Dict[str, torch.Tensor], nsteps: int,
                uses_pde: bool = True, graph: bool = False) -> torch.Tensor:
        """The autoregressive test rollout of train2D.py:230-245: returns
        [B, nsteps, 2, F, F].  ``stats`` holds inmean/instd/outmean/outstd/mumean/mustd
        (+ pdemean/pdestd).  ``graph=True`` captures one whole step (K1 -> net -> K2)
        in a CUDA graph and replays it, which is what makes B = 6..12 rollouts
        launch-latency free; to reuse the captured graph over many rollouts of the same
        batch size keep the :class:`RolloutPlan` (``prepare_rollout``)."""
        return self.prepare_rollout(net, u.shape[0], mu, stats, uses_pde, graph, device=u.device).run(u, nsteps)

    def prepare_rollout(self, net: Callable, batch: int, mu, stats: Dict[str, torch.Tensor],
                        uses_pde: bool = True, graph: bool = True, device="cuda") -> "RolloutPlan":
        return RolloutPlan(self, net, batch, mu, stats, uses_pde, graph, device)


class RolloutPlan:
    """One rollout configuration (network, batch size, parameters, statistics) with its static
    buffers and, with ``graph=True``, the CUDA graph of one step  K1 -> net -> K2 -> state copy.
    ``run(u0, nsteps)`` only replays; the capture (cuDNN autotune, allocator warm-up, graph
    instantiation) is paid once in the constructor."""

    @torch.no_grad()
    def __init__(self, step: CoarseStep, net: Callable, batch: int, mu, stats: Dict[str, torch.Tensor],
                 uses_pde: bool = True, graph: bool = True, device="cuda"):
        # inference only (train2D.py:227 `model.eval()`): never record autograd -- neither into the captured
        # graph's pool nor across the ping-pong state buffers.  Training through the step uses `step_update`.
        self.step, self.net, self.uses_pde = step, net, uses_pde
        dev = torch.device(device)
        B = int(batch)
        self.x = torch.zeros(B, 2, step.FH, step.FW, dtype=torch.float32, device=dev)
        self.nxt = torch.empty_like(self.x)
        self.muv = _mu_vec(mu, B, dev)
        self.mun = ((self.muv - stats["mumean"].to(dev)) / stats["mustd"].to(dev)).reshape(B, 1, 1, 1)
        self.om, self.os_ = L.f2(stats["outmean"]), L.f2(stats["outstd"])
        self.im, self.is_ = L.f2(stats["inmean"]), L.f2(stats["instd"])
        self.pm, self.ps = (L.f2(stats["pdemean"]), L.f2(stats["pdestd"])) if uses_pde else (None, None)
        self.ostd_t, self.omean_t = stats["outstd"].to(dev), stats["outmean"].to(dev)
        self.graph = None
        if graph:
            with torch.cuda.device(dev):
                stream = torch.cuda.Stream()
                stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(stream):
                    for _ in range(2):                      # warm-up outside capture (cudnn autotune, allocs)
                        self._one_step(self.x, self.nxt)
                torch.cuda.current_stream().wait_stream(stream)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._one_step(self.x, self.nxt)
                    self.x.copy_(self.nxt)
                self.graph = g

    @torch.no_grad()
    def _one_step(self, src, dst):
        st = self.step
        if self.uses_pde:
            un, pn, _ = st.step_inputs(src, self.muv, self.im, self.is_, self.pm, self.ps)
            nn_out = self.net(un, self.mun, pn)
            st.step_update_(src, self.muv, nn_out.contiguous(), self.om, self.os_, dst)
        else:       # train2D.py:240-242: no PDE increment at all
            un, _, _ = st.step_inputs(src, self.muv, self.im, self.is_)
            nn_out = self.net(un, self.mun)
            dst.copy_(torch.nn.functional.pad(nn_out * self.ostd_t + self.omean_t, (0, 1, 0, 1), mode="circular") + src)

    @torch.no_grad()
    def run(self, u: torch.Tensor, nsteps: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[B,2,F,F] -> [B, nsteps, 2, F, F] (``out`` may be a preallocated result tensor)."""
        self.step._check_state(u)
        x = self.x
        result = out if out is not None else torch.empty((x.shape[0], nsteps) + tuple(x.shape[1:]), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            x.copy_(L.dev(u, "state"))
            if self.graph is not None:
                for t in range(nsteps):
                    self.graph.replay()
                    result[:, t] = x
            else:
                nxt = self.nxt
                for t in range(nsteps):
                    self._one_step(x, nxt)
                    x, nxt = nxt, x
                    result[:, t] = x
        return result


class HostStepPipeline:
    """``u_next = padBC_rd(nn*outstd+outmean) + u + pde_du(u, mu)`` for states that live in
    (pinned) HOST memory -- where the reference keeps them between steps (train2D.py:230-245
    moves every batch ``.to(device)`` and back).

    The batch is cut into chunks that flow through three streams -- H2D copies, the fused kernel,
    D2H copies -- with a ring of device buffers and events between the stages, so the two copy
    engines and the SMs all stay busy; the step is PCIe-bound (12 bytes per point-update cross the
    link: 8 in, 4 out)."""

    def __init__(self, step: "CoarseStep", batch: int, chunks: int = 8, depth: int = 3, device="cuda"):
        self.step = step
        self.B = int(batch)
        self.cb = (self.B + chunks - 1) // chunks
        self.nchunk = (self.B + self.cb - 1) // self.cb
        dev = torch.device(device)
        F = (2, step.FH, step.FW)
        P = (2, step.FH - 1, step.FW - 1)
        self.depth = max(2, min(depth, self.nchunk))
        self.bufs = [(torch.empty((self.cb,) + F, device=dev), torch.empty((self.cb,) + P, device=dev),
                      torch.empty((self.cb,) + F, device=dev)) for _ in range(self.depth)]
        self.s_in, self.s_k, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        self.ev_in = [torch.cuda.Event() for _ in range(self.nchunk)]
        self.ev_k = [torch.cuda.Event() for _ in range(self.nchunk)]
        self.ev_out = [torch.cuda.Event() for _ in range(self.nchunk)]

    def run(self, u_host: torch.Tensor, mu_dev: torch.Tensor, nn_host: Optional[torch.Tensor],
            outmean, outstd, out_host: torch.Tensor) -> torch.Tensor:
        """One step; ``mu_dev`` [B] stays on the device.  Returns ``out_host`` (valid once the
        current stream has been synchronised)."""
        cur = torch.cuda.current_stream()
        for s in (self.s_in, self.s_k, self.s_out):
            s.wait_stream(cur)
        om = outmean if isinstance(outmean, C.Array) or outmean is None else L.f2(outmean)
        osd = outstd if isinstance(outstd, C.Array) or outstd is None else L.f2(outstd)
        for c in range(self.nchunk):
            b0, b1 = c * self.cb, min((c + 1) * self.cb, self.B)
            n = b1 - b0
            du, dn, do = self.bufs[c % self.depth]
            with torch.cuda.stream(self.s_in):
                if c >= self.depth:
                    self.s_in.wait_event(self.ev_k[c - self.depth])       # the kernel that read this buffer set
                du[:n].copy_(u_host[b0:b1], non_blocking=True)
                if nn_host is not None:
                    dn[:n].copy_(nn_host[b0:b1], non_blocking=True)
                self.ev_in[c].record(self.s_in)
            with torch.cuda.stream(self.s_k):
                self.s_k.wait_event(self.ev_in[c])
                if c >= self.depth:
                    self.s_k.wait_event(self.ev_out[c - self.depth])      # its output buffer has been drained
                self.step.step_update_(du[:n], mu_dev[b0:b1], dn[:n] if nn_host is not None else None, om, osd, do[:n])
                self.ev_k[c].record(self.s_k)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_k[c])
                out_host[b0:b1].copy_(do[:n], non_blocking=True)
                self.ev_out[c].record(self.s_out)
        for s in (self.s_in, self.s_k, self.s_out):
            cur.wait_stream(s)
        return out_host


class FineBurgers:
    """Fine-grid explicit Burgers generator (operators.py:403-492).

    ``mode='as_written'`` reproduces the live lines 474-475 (v is updated with the
    NEW u and with u's halos -- SURVEY quirk Q3); ``mode='intended'`` is the
    commented formula on the same lines.  ``endpoint=True`` is the reference grid
    (257 points, period 256); ``endpoint=False`` a true periodic grid."""

    def __init__(self, dt: float = 1e-4, dx: float = 0.0125, dy: float = 0.0125, mode: str = "as_written",
                 adv_accuracy: int = 3, diff_accuracy: int = 6, endpoint: bool = True):
        if mode not in ("as_written", "intended"):
            raise ValueError(mode)
        self.dt, self.dx, self.dy = float(dt), float(dx), float(dy)
        self.mode = L.FINE_AS_WRITTEN if mode == "as_written" else L.FINE_INTENDED
        self.adv_accuracy, self.diff_accuracy = int(adv_accuracy), int(diff_accuracy)
        self.endpoint = bool(endpoint)

    def step_(self, u, v, muv, uo, vo) -> None:
        """One Euler step u,v -> uo,vo ([B,1,H,W] or [B,H,W] fp32 CUDA, no aliasing)."""
        lib = L.load()
        B, H, W = u.shape[0], u.shape[-2], u.shape[-1]
        L.check(lib.ppnn_fine_burgers_step(u.data_ptr(), v.data_ptr(), muv.data_ptr(), uo.data_ptr(), vo.data_ptr(),
                                           B, H, W, int(self.endpoint), self.mode,
                                           self.adv_accuracy, self.diff_accuracy,
                                           self.dt, self.dx, self.dy, L.stream()), "ppnn_fine_burgers_step")

    @torch.no_grad()
    def step(self, u, v, mu) -> Tuple[torch.Tensor, torch.Tensor]:
        u, v = L.dev(u, "u"), L.dev(v, "v")
        if u.shape != v.shape:
            raise RuntimeError("u and v must have the same shape")
        muv = _mu_vec(mu, u.shape[0], u.device)
        uo, vo = torch.empty_like(u), torch.empty_like(v)
        with torch.cuda.device(u.device):
            self.step_(u, v, muv, uo, vo)
        return uo, vo

    @torch.no_grad()
    def run(self, u, v, mu, nsteps: int, every: int = 200, graph: bool = True,
            sink: Optional[Callable[[int, torch.Tensor, torch.Tensor], None]] = None):
        """The generator loop of operators.py:466-484: a snapshot is taken when
        ``i % every == 0`` (so the first one is the state after ONE step).  Snapshots
        go to ``sink(i, u, v)`` if given, else they are returned as lists.  With
        ``graph=True`` blocks of steps are replayed from one CUDA graph."""
        u, v = L.dev(u, "u").clone(), L.dev(v, "v").clone()
        muv = _mu_vec(mu, u.shape[0], u.device)
        u2, v2 = torch.empty_like(u), torch.empty_like(v)
        su: List[torch.Tensor] = []
        sv: List[torch.Tensor] = []

        def snap(i, a, b):
            if sink is not None:
                sink(i, a, b)
            else:
                su.append(a.clone())
                sv.append(b.clone())

        with torch.cuda.device(u.device):
            use_graph = graph and every >= 2 and every % 2 == 0 and nsteps > every
            if not use_graph:
                for i in range(nsteps):
                    self.step_(u, v, muv, u2, v2)
                    u, u2, v, v2 = u2, u, v2, v
                    if i % every == 0:
                        snap(i, u, v)
                return u, v, su, sv

            # graph path: the state lives in (u, v); one graph = `every` steps
            # (even, so the ping-pong lands back in u, v), replayed from a step index
            # i = 1 (mod every) so that its last step is a snapshot step.
            def one(i):
                self.step_(u, v, muv, u2, v2)
                u.copy_(u2)
                v.copy_(v2)
                if i % every == 0:
                    snap(i, u, v)

            one(0)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    for _ in range(every // 2):
                        self.step_(u, v, muv, u2, v2)
                        self.step_(u2, v2, muv, u, v)
            torch.cuda.current_stream().wait_stream(side)
            i = 1
            while i < nsteps:
                if i % every == 1 % every and i + every <= nsteps:
                    g.replay()
                    i += every
                    snap(i - 1, u, v)
                else:
                    one(i)
                    i += 1
        return u, v, su, sv


class FineRhsSolver:
    """Generic fused explicit Euler step on the fine grid: ``u, v <- u + dt*rhsu(u, v), v + dt*rhsv(u, v)``
    for any pair of ``rhs.py`` names (both right-hand sides see the OLD state, like ``pde_du``),
    with Upwind1(``adv_accuracy``) / Central2(``diff_accuracy``) stencils sharing one halo -- the kernel
    family of the commented reaction-diffusion / diffusion / convection generators of
    operators.py:494-748 (``ppnn_fine_rhs_step``)."""

    def __init__(self, rhsu: str, rhsv: str, dt: float, dx: float, dy: Optional[float] = None,
                 adv_accuracy: int = 3, diff_accuracy: int = 6, endpoint: bool = True):
        self.rhs_u, self.rhs_v = RHS_IDS[rhsu], RHS_IDS[rhsv]
        self.dt, self.dx, self.dy = float(dt), float(dx), float(dx if dy is None else dy)
        self.adv_accuracy, self.diff_accuracy = int(adv_accuracy), int(diff_accuracy)
        self.endpoint = bool(endpoint)

    def step_(self, u, v, muv, uo, vo) -> None:
        lib = L.load()
        B, H, W = u.shape[0], u.shape[-2], u.shape[-1]
        L.check(lib.ppnn_fine_rhs_step(u.data_ptr(), v.data_ptr(), muv.data_ptr(), uo.data_ptr(), vo.data_ptr(),
                                       B, H, W, int(self.endpoint), self.rhs_u, self.rhs_v,
                                       self.adv_accuracy, self.diff_accuracy,
                                       self.dt, self.dx, self.dy, L.stream()), "ppnn_fine_rhs_step")

    @torch.no_grad()
    def step(self, u, v, mu) -> Tuple[torch.Tensor, torch.Tensor]:
        u, v = L.dev(u, "u"), L.dev(v, "v")
        if u.shape != v.shape:
            raise RuntimeError("u and v must have the same shape")
        muv = _mu_vec(mu, u.shape[0], u.device)
        uo, vo = torch.empty_like(u), torch.empty_like(v)
        with torch.cuda.device(u.device):
            self.step_(u, v, muv, uo, vo)
        return uo, vo
