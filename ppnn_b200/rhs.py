"""Drop-in replacement for the reference ``src/rhs.py``.

Every 2-D right-hand side keeps the reference's 13-positional signature
``f(u1, v1, mu, dudx, dudy, d2udx2, d2udy2, ax, ay, dx, dy, dx2, dy2)`` (``ks1d``
keeps its 9), so ``getattr(rhs, params.rhsu)`` in train2D.py:50-51 resolves the
same names.  The sixteen functions are generated from one table: a right-hand
side is  [advection] + [diffusion] + [reaction]  (reference rhs.py:1-47,52-56).

When the four operator arguments are ppnn_b200 operators whose stencils share one
halo (true for every caller in the reference: train2D.py:53-56 uses accuracy 2/1,
operators.py:450-455 uses 6/3), the whole function is ONE fused kernel
(``ppnn_rhs2d_fwd``); otherwise it is composed from the operator calls.
"""
from __future__ import annotations

import torch

from . import _lib as L
from . import operators as _ops

# name -> ABI id  (include/ppnn_b200.h)
RHS_IDS = {
    "burgers2DFu": L.RHS_ADV | L.RHS_DIFF | L.REACT_BFU,      # rhs.py:1-3
    "burgers2DFv": L.RHS_ADV | L.RHS_DIFF | L.REACT_BFV,      # rhs.py:5-7
    "burgers2Dpu": L.RHS_ADV | L.RHS_DIFF,                    # rhs.py:9-11
    "burgers2Dpv": L.RHS_ADV | L.RHS_DIFF,                    # rhs.py:13-15
    "burgers2Ddiffuseu": L.RHS_DIFF,                          # rhs.py:17-19
    "burgers2Ddiffusev": L.RHS_DIFF,                          # rhs.py:21-23
    "rdFu": L.RHS_DIFF | L.REACT_RDU,                         # rhs.py:25-26
    "rdFv": L.RHS_DIFF | L.REACT_RDV,                         # rhs.py:28-29
    "rdpu": L.RHS_DIFF,                                       # rhs.py:31-32
    "rdpv": L.RHS_DIFF,                                       # rhs.py:34-35
    "diffu": L.RHS_DIFF,                                      # rhs.py:37-38
    "diffv": L.RHS_DIFF,                                      # rhs.py:40-41
    "convecu": L.RHS_ADV,                                     # rhs.py:43-44
    "convecv": L.RHS_ADV,                                     # rhs.py:46-47
    "burgers2DWu": L.REACT_RDU,                               # rhs.py:52-53
    "burgers2DWv": L.REACT_RDV,                               # rhs.py:55-56
}


def rhs_id(name: str) -> int:
    return RHS_IDS[name]


def _reaction(react, u1, v1):
    if react == L.REACT_BFU:
        uv = u1 * u1 + v1 * v1
        return (1 - uv) * u1 + uv * v1
    if react == L.REACT_BFV:
        uv = u1 * u1 + v1 * v1
        return -uv * u1 + (1 - uv) * v1
    if react == L.REACT_RDU:
        return u1 - u1 * u1 * u1 - v1 + 0.01
    if react == L.REACT_RDV:
        return 0.25 * (u1 - v1)
    return None


def _reaction_jac(react, u1, v1):
    uv = u1 * u1 + v1 * v1
    if react == L.REACT_BFU:
        return 1 - uv - 2 * u1 * u1 + 2 * u1 * v1, uv - 2 * u1 * v1 + 2 * v1 * v1
    if react == L.REACT_BFV:
        return -uv - 2 * u1 * u1 - 2 * u1 * v1, 1 - uv - 2 * u1 * v1 - 2 * v1 * v1
    if react == L.REACT_RDU:
        return 1 - 3 * u1 * u1, -torch.ones_like(v1)
    if react == L.REACT_RDV:
        return torch.full_like(u1, 0.25), torch.full_like(v1, -0.25)
    return None, None


def _compose(rid, u1, v1, mu, dudx, dudy, d2udx2, d2udy2, ax, ay, dx, dy, dx2, dy2):
    """Term-by-term evaluation through the operator objects (any operator
    implementation, any halo combination)."""
    out = None
    if rid & L.RHS_ADV:
        out = -u1 * dudx(ax) / dx - v1 * dudy(ay) / dy
    if rid & L.RHS_DIFF:
        lap = mu * (d2udx2(ax) / dx2 + d2udy2(ay) / dy2)
        out = lap if out is None else out + lap
    r = _reaction(rid & ~3, u1, v1)
    if r is not None:
        out = r if out is None else out + r
    return out


def _fusable(rid, u1, v1, mu, dudx, dudy, d2udx2, d2udy2, ax, ay):
    if not (isinstance(u1, torch.Tensor) and u1.is_cuda and u1.dim() == 4 and u1.shape[1] == 1):
        return None
    if not (rid & 3):
        return None                                # pure reaction: plain element-wise algebra
    for t in (v1, ax, ay):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32):
            return None
    if u1.dtype != torch.float32 or v1.shape != u1.shape:
        return None
    B, _, H, W = u1.shape
    if ax.shape[0] != B or ay.shape[0] != B or ax.shape[3] != W or ay.shape[2] != H:
        return None
    h2 = ax.shape[2] - H
    if h2 < 0 or h2 % 2 or ay.shape[3] - W != h2:
        return None
    halo = h2 // 2
    k = 2 * halo + 1
    adv = (0, 0)
    dif = (0, 0)
    if rid & L.RHS_ADV:
        if not (isinstance(dudx, _ops.dudx_2D) and isinstance(dudy, _ops.dudy_2D)):
            return None
        if (dudx._scheme_id, dudx.accuracy) != (dudy._scheme_id, dudy.accuracy):
            return None
        if _ops._taps(dudx._scheme_id, dudx.accuracy) != k:
            return None
        adv = (dudx._scheme_id, dudx.accuracy)
    if rid & L.RHS_DIFF:
        if not (isinstance(d2udx2, _ops.d2udx2_2D) and isinstance(d2udy2, _ops.d2udy2_2D)):
            return None
        if (d2udx2._scheme_id, d2udx2.accuracy) != (d2udy2._scheme_id, d2udy2.accuracy):
            return None
        if _ops._taps(d2udx2._scheme_id, d2udx2.accuracy) != k:
            return None
        dif = (d2udx2._scheme_id, d2udx2.accuracy)
    if isinstance(mu, torch.Tensor):
        if mu.numel() not in (1, B) or not mu.is_cuda:
            return None
    elif not isinstance(mu, (int, float)):
        return None
    return halo, adv, dif


def _mu_vector(mu, B, device):
    if isinstance(mu, torch.Tensor):
        return mu.to(torch.float32).reshape(-1).expand(B).contiguous()
    return torch.full((B,), float(mu), dtype=torch.float32, device=device)


def _cropped_view(u1, v1, halo):
    """(plane stride, row pitch) when u1 and v1 are the same kind of row-dense strided view of small planes --
    ``down(u)[:, :, :-1, :-1]`` (train2D.py:68-69) -- that ``ppnn_rhs2d_fwd_strided`` reads in place; else None."""
    if u1.is_contiguous() or u1.stride() != v1.stride():
        return None
    B, _, H, W = u1.shape
    sb, _, sh, sw = u1.stride()
    if sw != 1 or sh < W or sb < (H - 1) * sh + W or W > 128:
        return None
    if 4 * ((H + 2 * halo) * W + H * (W + 2 * halo) + 2 * H * W + 12) > 100 * 1024:     # the whole-plane kernel's shared memory
        return None
    return sb, sh


def _fused_forward(u1, v1, mu, ax, ay, rid, halo, adv, dif, scal):
    lib = L.load()
    B, _, H, W = u1.shape
    axc, ayc = L.dev(ax), L.dev(ay)
    muv = _mu_vector(mu, B, u1.device)
    dx, dy, dx2, dy2 = scal
    view = _cropped_view(u1, v1, halo)
    if view is not None:
        out = torch.empty((B, 1, H, W), dtype=torch.float32, device=u1.device)
        with L.on_device(u1):
            L.check(lib.ppnn_rhs2d_fwd_strided(u1.data_ptr(), v1.data_ptr(), muv.data_ptr(), axc.data_ptr(),
                                               ayc.data_ptr(), out.data_ptr(), B, H, W, halo, rid,
                                               adv[0], adv[1], dif[0], dif[1], dx, dy, dx2, dy2,
                                               view[0], view[1], L.stream()), "ppnn_rhs2d_fwd_strided")
        return out, (u1, v1, muv, axc, ayc)
    u1c, v1c = L.dev(u1), L.dev(v1)
    out = torch.empty_like(u1c)
    with L.on_device(u1c):
        L.check(lib.ppnn_rhs2d_fwd(u1c.data_ptr(), v1c.data_ptr(), muv.data_ptr(), axc.data_ptr(),
                                   ayc.data_ptr(), out.data_ptr(), B, H, W, halo, rid,
                                   adv[0], adv[1], dif[0], dif[1], dx, dy, dx2, dy2, L.stream()),
                "ppnn_rhs2d_fwd")
    return out, (u1c, v1c, muv, axc, ayc)


class _FusedRhs(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u1, v1, mu, ax, ay, rid, halo, adv, dif, scal):
        out, saved = _fused_forward(u1, v1, mu, ax, ay, rid, halo, adv, dif, scal)
        ctx.meta = (rid, halo, adv, dif, scal, isinstance(mu, torch.Tensor) and mu.shape)
        ctx.save_for_backward(*saved)
        return out

    @staticmethod
    def backward(ctx, g):
        # adjoint composed from the operator adjoint kernels + element-wise algebra
        rid, halo, adv, dif, (dx, dy, dx2, dy2), mu_shape = ctx.meta
        u1, v1, muv, ax, ay = ctx.saved_tensors
        B = u1.shape[0]
        mu4 = muv.reshape(B, 1, 1, 1)
        g = g.contiguous()
        gu1 = torch.zeros_like(u1)
        gv1 = torch.zeros_like(v1)
        gax = torch.zeros_like(ax)
        gay = torch.zeros_like(ay)
        gmu = None
        F = _ops._StencilFn
        if rid & L.RHS_ADV:
            with torch.enable_grad():
                axr, ayr = ax.detach().requires_grad_(True), ay.detach().requires_grad_(True)
                ddx, ddy = F.apply(axr, adv[0], adv[1], 2), F.apply(ayr, adv[0], adv[1], 3)
            gu1 = gu1 - ddx.detach() / dx * g
            gv1 = gv1 - ddy.detach() / dy * g
            ga, gb = torch.autograd.grad((ddx, ddy), (axr, ayr), (-u1 / dx * g, -v1 / dy * g))
            gax, gay = gax + ga, gay + gb
        if rid & L.RHS_DIFF:
            with torch.enable_grad():
                axr, ayr = ax.detach().requires_grad_(True), ay.detach().requires_grad_(True)
                lx, ly = F.apply(axr, dif[0], dif[1], 2), F.apply(ayr, dif[0], dif[1], 3)
            gmu = ((lx.detach() / dx2 + ly.detach() / dy2) * g).sum(dim=(1, 2, 3))
            ga, gb = torch.autograd.grad((lx, ly), (axr, ayr), (mu4 / dx2 * g, mu4 / dy2 * g))
            gax, gay = gax + ga, gay + gb
        ju, jv = _reaction_jac(rid & ~3, u1, v1)
        if ju is not None:
            gu1 = gu1 + ju * g
            gv1 = gv1 + jv * g
        gmu_out = None
        if mu_shape is not False and ctx.needs_input_grad[2]:
            if gmu is None:
                gmu = torch.zeros(B, device=u1.device)
            gmu_out = gmu.sum().reshape(mu_shape) if int(torch.Size(mu_shape).numel()) == 1 else gmu.reshape(mu_shape)
        return gu1, gv1, gmu_out, gax, gay, None, None, None, None, None


def _make(name, rid):
    def f(u1, v1, mu, dudx, dudy, d2udx2, d2udy2, ax, ay, dx, dy, dx2, dy2):
        plan = _fusable(rid, u1, v1, mu, dudx, dudy, d2udx2, d2udy2, ax, ay)
        if plan is None:
            return _compose(rid, u1, v1, mu, dudx, dudy, d2udx2, d2udy2, ax, ay, dx, dy, dx2, dy2)
        halo, adv, dif = plan
        scal = (float(dx), float(dy), float(dx2), float(dy2))
        needs_grad = torch.is_grad_enabled() and any(
            isinstance(t, torch.Tensor) and t.requires_grad for t in (u1, v1, mu, ax, ay))
        if not needs_grad:
            return _fused_forward(u1, v1, mu, ax, ay, rid, halo, adv, dif, scal)[0]
        return _FusedRhs.apply(u1, v1, mu, ax, ay, rid, halo, adv, dif, scal)
    f.__name__ = f.__qualname__ = name
    f.__doc__ = f"reference rhs.py `{name}`; ABI id {rid} (include/ppnn_b200.h)."
    return f


for _name, _rid in RHS_IDS.items():
    globals()[_name] = _make(_name, _rid)
del _name, _rid


def ks1d(ux4, mu, dudx, d2udx2, d4udx4, ux, dx, dx2, dx4):
    """1-D Kuramoto-Sivashinsky right-hand side (reference rhs.py:49-50); each
    term is one ppnn_stencil_fwd launch when the operators are ppnn_b200's."""
    return - d2udx2(ux) / dx2 - d4udx4(ux4) / dx4 - 0.5 * dudx(ux * ux) / dx
