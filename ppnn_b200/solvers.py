"""Tier B: the reference's small explicit solvers, many time steps per kernel launch with the state on chip
(``csrc/onchip_solvers.cu``).

* :class:`KS2D` -- the 2-D Kuramoto-Sivashinsky generator with classic RK4, ``src/operators.py:750-843``
  (``rk4`` / ``f`` closures at ``:805-815``);
* :class:`KS1D` -- the 1-D Kuramoto-Sivashinsky generator, ``src/operators.py:845-918`` (``f`` at ``:883-884`` =
  ``rhs.ks1d``, ``src/rhs.py:49-50``, plus a source term);
* :class:`HeatBC1D` -- the 1-D heat equation with Dirichlet / Neumann ghost cells of ``Bv/data.py``: ``padder``
  (``:9-28``), the generator loop (``:92-97``) and ``genforward`` (``:39-51``).

All three keep the reference's closure names (``rk4``, ``pad``, ``forward``) for one step and add ``run`` for the
whole loop.  There is no CPU fallback; a mesh that does not fit on chip composes the step from the Tier-A operator
kernels (:mod:`ppnn_b200.operators`), still on the GPU.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch

from . import _lib as L
from . import operators as ops


def _as_batch(u: torch.Tensor, spatial: int) -> Tuple[torch.Tensor, torch.Size]:
    """[B,1,*spatial] or [B,*spatial] -> contiguous tensor + the caller's shape."""
    shape = u.shape
    if u.dim() == spatial + 2:
        if u.shape[1] != 1:
            raise RuntimeError(f"expected a single-channel field, got shape {tuple(shape)}")
    elif u.dim() != spatial + 1:
        raise RuntimeError(f"expected [B,1,...] or [B,...] with {spatial} spatial dims, got shape {tuple(shape)}")
    return L.dev(u, "u"), shape


class KS1D:
    """1-D Kuramoto-Sivashinsky, classic RK4 (``src/operators.py:845-918``).

    ``f(u) = -d2udx2(pad(u))/dx2 - d4udx4(pad(u))/dx4 - 0.5*dudx(pad(u*u))/dx + source`` with
    ``diffusion1D('Central', 6)``, ``diffusion1D('Central_4th', 4)``, ``dudx_1D('Upwind', 3)`` and the
    duplicated-endpoint pad ``cat(u[-4:-1], u, u[1:4])``.  ``source`` is the reference's
    ``0.1*exp(-(x - 4*pi)**2/2)`` ([n], [1,1,n] or None)."""

    def __init__(self, dt: float = 1e-4, dx: float = math.pi / 8, source: Optional[torch.Tensor] = None,
                 adv_accuracy: int = 3, d2_accuracy: int = 6, d4_accuracy: int = 4, endpoint: bool = True):
        self.dt, self.dx = float(dt), float(dx)
        self.dx2 = self.dx ** 2
        self.dx4 = self.dx2 ** 2
        self.acc = (int(adv_accuracy), int(d2_accuracy), int(d4_accuracy))
        self.endpoint = bool(endpoint)
        self.source = None if source is None else L.dev(source.reshape(-1), "source")

    @torch.no_grad()
    def run(self, u: torch.Tensor, nsteps: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``nsteps`` RK4 steps in ONE launch."""
        u, shape = _as_batch(u, 1)
        B, n = u.shape[0], u.shape[-1]
        if self.source is not None and self.source.numel() != n:
            raise RuntimeError(f"source has {self.source.numel()} points, the state {n}")
        out = torch.empty_like(u) if out is None else out
        with L.on_device(u):
            L.check(L.load().ppnn_ks1d_rk4(u.data_ptr(), L.ptr(self.source), out.data_ptr(), B, n,
                                           int(self.endpoint), int(nsteps), *self.acc,
                                           self.dt, self.dx, self.dx2, self.dx4, L.stream()), "ppnn_ks1d_rk4")
        return out.reshape(shape)

    def rk4(self, u: torch.Tensor) -> torch.Tensor:
        """One step (the reference's ``rk4`` closure, ``operators.py:882-890``)."""
        return self.run(u, 1)

    @torch.no_grad()
    def f(self, u: torch.Tensor) -> torch.Tensor:
        """The right-hand side alone (the reference's ``f`` closure, ``operators.py:883-884``)."""
        u, shape = _as_batch(u, 1)
        B, n = u.shape[0], u.shape[-1]
        out = torch.empty_like(u)
        with L.on_device(u):
            L.check(L.load().ppnn_ks1d_rhs(u.data_ptr(), L.ptr(self.source), out.data_ptr(), B, n, int(self.endpoint),
                                           *self.acc, self.dx, self.dx2, self.dx4, L.stream()), "ppnn_ks1d_rhs")
        return out.reshape(shape)

    @torch.no_grad()
    def generate(self, u: torch.Tensor, nsteps: int, every: int = 500) -> torch.Tensor:
        """The loop of ``operators.py:899-907``: snapshot when ``i % every == 0`` (after step i) ->
        ``[B, T, ...]`` stacked on dim 1 like ``torch.stack(resultu, dim=1)``."""
        snaps: List[torch.Tensor] = []
        done = 0
        for i in range(0, nsteps, every):          # snapshot steps: i = 0, every, 2*every, ...
            u = self.run(u, i + 1 - done)
            done = i + 1
            snaps.append(u)
        if done < nsteps:
            u = self.run(u, nsteps - done)
        return torch.stack(snaps, dim=1)


class KS2D:
    """2-D Kuramoto-Sivashinsky, classic RK4 (``src/operators.py:750-843``).

    ``f(u)`` as written at ``:806-809`` (``c2 = 0.1`` in front of the second derivatives; the mixed term
    ``d2udy2(pad(d2udx2(pad u)))`` is not divided by ``dx2*dy2`` there: ``mixed_scale = 1``), ``Central2`` acc 6,
    ``Central4`` acc 4, ``Upwind1`` acc 3 on ``u*u``, pads ``cat(u[-4:-1], u, u[1:4])``."""

    def __init__(self, dt: float = 1e-3, dx: float = math.pi / 8, dy: Optional[float] = None, c2: float = 0.1,
                 mixed_scale: float = 1.0, adv_accuracy: int = 3, d2_accuracy: int = 6, d4_accuracy: int = 4,
                 endpoint: bool = True):
        self.dt, self.dx = float(dt), float(dx)
        self.dy = self.dx if dy is None else float(dy)
        self.dx2, self.dy2 = self.dx ** 2, self.dy ** 2
        self.dx4, self.dy4 = self.dx2 ** 2, self.dy2 ** 2
        self.c2, self.mixed_scale = float(c2), float(mixed_scale)
        self.acc = (int(adv_accuracy), int(d2_accuracy), int(d4_accuracy))
        self.endpoint = bool(endpoint)
        self._tier_a = None

    # -- Tier-A composition (meshes beyond the on-chip limit; also the cross-check of the fused kernel) --------
    def _operators(self, device):
        if self._tier_a is None:
            a, d2, d4 = self.acc
            self._tier_a = (ops.dudx_2D('Upwind1', accuracy=a, device=device), ops.dudy_2D('Upwind1', accuracy=a, device=device),
                            ops.d2udx2_2D('Central2', accuracy=d2, device=device), ops.d2udy2_2D('Central2', accuracy=d2, device=device),
                            ops.d2udx2_2D('Central4', accuracy=d4, device=device), ops.d2udy2_2D('Central4', accuracy=d4, device=device))
        return self._tier_a

    def f_composed(self, u: torch.Tensor) -> torch.Tensor:
        """``f`` from the Tier-A operator kernels, written as the reference writes it (``operators.py:806-809``).
        All three stencil families must share one halo (the reference pads once)."""
        dudx, dudy, d2udx2, d2udy2, d4udx4, d4udy4 = self._operators(u.device)
        h = max(self.acc[0], self.acc[1] // 2, self.acc[2] // 2 + 1)
        if not (self.acc[0] == self.acc[1] // 2 == self.acc[2] // 2 + 1):
            raise RuntimeError("the composed KS right-hand side needs stencils of one halo (e.g. accuracies 3 / 6 / 4)")
        if self.endpoint:
            def padx(w): return torch.cat((w[:, :, -(h + 1):-1], w, w[:, :, 1:h + 1]), dim=2)
            def pady(w): return torch.cat((w[:, :, :, -(h + 1):-1], w, w[:, :, :, 1:h + 1]), dim=3)
        else:
            def padx(w): return torch.cat((w[:, :, -h:], w, w[:, :, :h]), dim=2)
            def pady(w): return torch.cat((w[:, :, :, -h:], w, w[:, :, :, :h]), dim=3)
        c2, m = self.c2, self.mixed_scale
        return (- c2 * d2udx2(padx(u)) / self.dx2 - c2 * d2udy2(pady(u)) / self.dy2 - d4udx4(padx(u)) / self.dx4
                - d4udy4(pady(u)) / self.dy4 - m * d2udy2(pady(d2udx2(padx(u)))) - m * d2udx2(padx(d2udy2(pady(u))))
                - 0.5 * dudx(padx(u * u)) / self.dx - 0.5 * dudy(pady(u * u)) / self.dy)

    def rk4_composed(self, u: torch.Tensor) -> torch.Tensor:
        f, dt = self.f_composed, self.dt
        k1 = f(u)
        k2 = f(u + dt * k1 / 2)
        k3 = f(u + dt * k2 / 2)
        k4 = f(u + dt * k3)
        return u + dt * (k1 + 2 * k2 + 2 * k3 + k4) / 6

    # -- fused ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def run(self, u: torch.Tensor, nsteps: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``nsteps`` RK4 steps in ONE launch (state on chip); composed from Tier A when the mesh is too big."""
        u, shape = _as_batch(u, 2)
        B, H, W = u.shape[0], u.shape[-2], u.shape[-1]
        out = torch.empty_like(u) if out is None else out
        with L.on_device(u):
            st = L.load().ppnn_ks2d_rk4(u.data_ptr(), out.data_ptr(), B, H, W, int(self.endpoint), int(nsteps),
                                        *self.acc, self.c2, self.mixed_scale, self.dt, self.dx, self.dy,
                                        self.dx2, self.dy2, self.dx4, self.dy4, L.stream())
            if st == L.E_TOOBIG:
                w = u.reshape(B, 1, H, W)
                for _ in range(int(nsteps)):
                    w = self.rk4_composed(w)
                out.copy_(w.reshape(out.shape))
            else:
                L.check(st, "ppnn_ks2d_rk4")
        return out.reshape(shape)

    def rk4(self, u: torch.Tensor) -> torch.Tensor:
        """One step (the reference's ``rk4`` closure, ``operators.py:805-815``)."""
        return self.run(u, 1)

    @torch.no_grad()
    def f(self, u: torch.Tensor) -> torch.Tensor:
        """The right-hand side alone (the reference's ``f`` closure, ``operators.py:806-809``)."""
        u, shape = _as_batch(u, 2)
        B, H, W = u.shape[0], u.shape[-2], u.shape[-1]
        out = torch.empty_like(u)
        with L.on_device(u):
            st = L.load().ppnn_ks2d_rhs(u.data_ptr(), out.data_ptr(), B, H, W, int(self.endpoint), *self.acc,
                                        self.c2, self.mixed_scale, self.dx, self.dy, self.dx2, self.dy2,
                                        self.dx4, self.dy4, L.stream())
            if st == L.E_TOOBIG:
                return self.f_composed(u.reshape(B, 1, H, W)).reshape(shape)
            L.check(st, "ppnn_ks2d_rhs")
        return out.reshape(shape)


class HeatBC1D:
    """1-D heat equation with per-sample Dirichlet (``cls == 0``) / Neumann (``cls == 1``) ghost cells
    (``Bv/data.py``).  ``cls`` and ``bv`` are ``[B,1,2]`` (left, right) as in the reference; ``bv`` is the
    de-normalised boundary value (``denormBV``)."""

    def __init__(self, dt: float, dx: float):
        self.dt, self.dx = float(dt), float(dx)
        self.dx2 = float(torch.tensor(self.dx, dtype=torch.float32) ** 2)   # dx is an fp32 tensor in the reference

    @staticmethod
    def _bc(cls: torch.Tensor, bv: torch.Tensor, B: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
        c = cls.to(device=device, dtype=torch.int32).reshape(-1).contiguous()
        b = bv.to(device=device, dtype=torch.float32).reshape(-1).contiguous()
        if c.numel() != 2 * B or b.numel() != 2 * B:
            raise RuntimeError(f"cls / bv must hold 2 values per sample (B = {B}), got {tuple(cls.shape)} / {tuple(bv.shape)}")
        return c, b

    @torch.no_grad()
    def pad(self, u: torch.Tensor, cls: torch.Tensor, bv: torch.Tensor) -> torch.Tensor:
        """``padder(dx)(u, cls, bv)`` (``Bv/data.py:25-28``): ``[B,1,n] -> [B,1,n+2]``."""
        u, shape = _as_batch(u, 1)
        B, n = u.shape[0], u.shape[-1]
        c, b = self._bc(cls, bv, B, u.device)
        out = torch.empty(*shape[:-1], n + 2, device=u.device, dtype=torch.float32)
        with L.on_device(u):
            L.check(L.load().ppnn_heat1d_pad(u.data_ptr(), c.data_ptr(), b.data_ptr(), out.data_ptr(),
                                             B, n, self.dx, L.stream()), "ppnn_heat1d_pad")
        return out

    @torch.no_grad()
    def step_update(self, u: torch.Tensor, nn_out: torch.Tensor, pdeu: Optional[torch.Tensor],
                    cls: torch.Tensor, bv: torch.Tensor) -> torch.Tensor:
        """The rollout update of ``Bv/train.py:162`` / ``Bv/test.py``:
        ``fpad(u[...,1:-1] + nnu + pdeu[...,1:-1], cls, bv)`` on the padded state ``u [B,1,n+2]``, the network
        increment ``nn_out [B,1,n]`` and the coarse PDE increment ``pdeu [B,1,n+2]`` (None: network-only model)."""
        u, shape = _as_batch(u, 1)
        B, n = u.shape[0], u.shape[-1] - 2
        nn_out = L.dev(nn_out, "nn_out")
        if nn_out.numel() != B * n:
            raise RuntimeError(f"nn_out has shape {tuple(nn_out.shape)}, expected [B,1,{n}]")
        if pdeu is not None:
            pdeu = L.dev(pdeu, "pdeu")
            if pdeu.shape != u.shape:
                raise RuntimeError(f"pdeu has shape {tuple(pdeu.shape)}, expected {tuple(u.shape)}")
        c, b = self._bc(cls, bv, B, u.device)
        out = torch.empty_like(u)
        with L.on_device(u):
            L.check(L.load().ppnn_heat1d_update(u.data_ptr(), nn_out.data_ptr(), L.ptr(pdeu), c.data_ptr(), b.data_ptr(),
                                                out.data_ptr(), B, n, self.dx, L.stream()), "ppnn_heat1d_update")
        return out.reshape(shape)

    @torch.no_grad()
    def run(self, u: torch.Tensor, cls: torch.Tensor, bv: torch.Tensor, nsteps: int,
            padded_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``nsteps`` steps ``u <- u + diffusion1D(padder(u))/dx2*dt`` in ONE launch.  ``padded_out`` (a
        ``[B, n+2]`` view whose rows are contiguous, e.g. ``results[:, k, 0]``) receives ``padder(u)`` of the
        incoming state."""
        u, shape = _as_batch(u, 1)
        B, n = u.shape[0], u.shape[-1]
        c, b = self._bc(cls, bv, B, u.device)
        out = torch.empty_like(u)
        pptr, pstride = None, 0
        if padded_out is not None:
            if (padded_out.dim() != 2 or padded_out.shape != (B, n + 2) or padded_out.stride(1) != 1
                    or not padded_out.is_cuda or padded_out.dtype != torch.float32):
                raise RuntimeError("padded_out must be a float32 CUDA [B, n+2] view with contiguous rows")
            pptr, pstride = padded_out.data_ptr(), padded_out.stride(0)
        with L.on_device(u):
            L.check(L.load().ppnn_heat1d_bc_steps(u.data_ptr(), c.data_ptr(), b.data_ptr(), out.data_ptr(),
                                                  pptr, pstride, B, n, int(nsteps), self.dt, self.dx, self.dx2,
                                                  L.stream()), "ppnn_heat1d_bc_steps")
        return out.reshape(shape)

    @torch.no_grad()
    def generate(self, u: torch.Tensor, cls: torch.Tensor, bv: torch.Tensor, nsteps: int, saveevery: int) -> torch.Tensor:
        """The loop of ``Bv/data.py:92-98``: ``results[:, k] = padder(u)`` at steps ``t = k*saveevery`` (taken
        BEFORE the update of step t), ``t < nsteps`` -> ``[B, T, 1, n+2]`` on the device."""
        u, _ = _as_batch(u, 1)
        B, n = u.shape[0], u.shape[-1]
        T = (int(nsteps) + saveevery - 1) // saveevery
        results = torch.empty(B, T, 1, n + 2, device=u.device, dtype=torch.float32)
        u = u.reshape(B, n)
        for k in range(T):
            steps = min(saveevery, nsteps - k * saveevery)
            u = self.run(u, cls, bv, steps, padded_out=results[:, k, 0])
        return results


class HeatCoarseIncrement:
    """``genforward(diffusion1D(accuracy=2), down, up, padder(cdx), dt, cdx2)`` of ``Bv/data.py:39-51,113-117``:
    the coarse-grid PDE increment of the PADDED fine state, ``forward(u, cls, bv) -> up(uc1) - u``."""

    def __init__(self, scale: float, dt: float, cdx: float):
        self.scale, self.dt, self.cdx = float(scale), float(dt), float(cdx)
        self.cdx2 = float(torch.tensor(self.cdx, dtype=torch.float32) ** 2)

    @torch.no_grad()
    def forward(self, u: torch.Tensor, cls: torch.Tensor, bv: torch.Tensor) -> torch.Tensor:
        u, shape = _as_batch(u, 1)
        B, nf = u.shape[0], u.shape[-1]
        nc = int(math.floor(nf * self.scale))
        c, b = HeatBC1D._bc(cls, bv, B, u.device)
        out = torch.empty_like(u)
        with L.on_device(u):
            L.check(L.load().ppnn_heat1d_coarse_increment(u.data_ptr(), c.data_ptr(), b.data_ptr(), out.data_ptr(),
                                                          B, nf, nc, self.dt, self.cdx, self.cdx2, L.stream()),
                    "ppnn_heat1d_coarse_increment")
        return out.reshape(shape)

    __call__ = forward
