"""Tier A mirror of ``Bv/data.py`` (the 1-D boundary-value workload): ``padder``, ``denormBV``, ``genforward``
with the reference's names, signatures and call conventions, on CUDA tensors only.

``padder.__call__`` is one kernel (``ppnn_heat1d_pad``); ``genforward`` composes the caller's ``fn`` / ``dfn`` /
``ufn`` / ``pfn`` exactly as ``Bv/data.py:41-49`` does, so it works with any callables.  The fully fused forms
(many steps per launch, restriction + coarse step + ghost cells + prolongation in one kernel) are
:class:`ppnn_b200.solvers.HeatBC1D` and :class:`ppnn_b200.solvers.HeatCoarseIncrement`.
"""
from __future__ import annotations

import torch

from . import _lib as L


class padder(object):
    """``Bv/data.py:9-28``: ghost cells for Dirichlet (``cls == 0``) / Neumann (``cls == 1``) boundaries."""

    def __init__(self, dx):
        self.dx = dx

    def pad2(self, u, bv):
        return u + bv * self.dx

    @torch.no_grad()
    def __call__(self, u, cls, bv):
        u = L.dev(u, "u")
        n = u.shape[-1]
        B = u.numel() // n
        c = cls.to(device=u.device, dtype=torch.int32).reshape(-1).contiguous()
        b = L.dev(bv.to(device=u.device, dtype=torch.float32), "bv").reshape(-1)
        if c.numel() != 2 * B or b.numel() != 2 * B:
            raise RuntimeError(f"cls / bv must hold (left, right) per row of u: {tuple(cls.shape)} / {tuple(bv.shape)} "
                               f"for u {tuple(u.shape)}")
        out = torch.empty(*u.shape[:-1], n + 2, device=u.device, dtype=torch.float32)
        with L.on_device(u):
            L.check(L.load().ppnn_heat1d_pad(u.data_ptr(), c.data_ptr(), b.data_ptr(), out.data_ptr(),
                                             B, n, float(self.dx), L.stream()), "ppnn_heat1d_pad")
        return out


def denormBV(BV, cls):
    """``Bv/data.py:31-32``."""
    return (cls == 1) * BV + (cls == 0) * 5 * BV


bids = [0, 1, 2, -3, -2, -1]


def genforward(fn, dfn, ufn, pfn, dt, cdx2):
    """``Bv/data.py:39-51``: coarse-grid increment of the padded fine state."""

    def forward(u, cls, bv):
        uc = dfn(u)
        cdu = fn(uc) / cdx2 * dt
        uc1 = pfn(uc[..., 1:-1] + cdu, cls, bv)
        duf1 = ufn(uc1) - u
        return duf1

    return forward
