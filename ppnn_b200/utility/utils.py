"""Drop-in replacement for the hot-path part of the reference ``src/utility/utils.py``:
``mesh_convertor`` (fine <-> coarse mesh conversion, utils.py:13-31) and
``model_count`` (utils.py:9-10).

``down`` is torch's ``interpolate(mode='bilinear'|'linear', size=cmesh)`` and ``up``
is ``interpolate(mode='bicubic', size=fmesh)`` -- here each is one sm_100a kernel
(``ppnn_resample_fwd``; adjoint ``ppnn_resample_bwd``) with identical index and
weight arithmetic (Keys cubic A = -0.75, clamped borders).  CUDA only.
"""
from __future__ import annotations

from typing import Tuple, Union

import torch

from .. import _lib as L

Size = Union[int, Tuple[int, int]]


def model_count(model):
    return sum(p.numel() for p in model.parameters())


def numpy2string(x, format='(%.6g %.6g %.0g)'):
    """Rows of ``x`` as text in an OpenFOAM vector-list format (utils.py:34-37; host-side helper of the OpenFOAM
    bridge, kept so that ``from utility.utils import numpy2string`` resolves through the drop-in)."""
    import io
    import numpy as np
    buf = io.StringIO()
    np.savetxt(buf, x, fmt=format)
    return buf.getvalue()


def _hw(size: Size) -> Tuple[int, int]:
    if isinstance(size, (tuple, list)):
        if len(size) == 1:
            return int(size[0]), int(size[0])
        return int(size[0]), int(size[1])
    return int(size), int(size)


def _plane_stride(x):
    """Stride (in elements) between the planes of a [B, 1, H, W] channel slice whose planes are dense --
    e.g. ``u[:, :1]`` of a contiguous [B, 2, H, W] state (train2D.py:68-69) -- else None."""
    if (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 1
            and not x.is_contiguous() and x.stride(3) == 1 and x.stride(2) == x.shape[3]
            and x.stride(0) >= x.shape[2] * x.shape[3]):
        return x.stride(0)
    return None


def _resample_forward(x, hout, wout, mode, align):
    lib = L.load()
    stride = _plane_stride(x)
    if stride is None:
        x = L.dev(x, "mesh_convertor input")
    hin, win = x.shape[-2], x.shape[-1]
    planes = x.numel() // (hin * win)
    out = torch.empty(x.shape[:-2] + (hout, wout), dtype=torch.float32, device=x.device)
    with L.on_device(x):
        if stride is None:
            L.check(lib.ppnn_resample_fwd(x.data_ptr(), out.data_ptr(), planes, hin, win, hout, wout,
                                          mode, int(align), L.stream()), "ppnn_resample_fwd")
        else:                       # channel slice: read in place, no contiguous copy
            L.check(lib.ppnn_resample_fwd_strided(x.data_ptr(), out.data_ptr(), planes, stride, hin, win, hout, wout,
                                                  mode, int(align), L.stream()), "ppnn_resample_fwd_strided")
    return out, (planes, hin, win, hout, wout, mode, int(align), x.shape)


def _resample(x, hout, wout, mode, align):
    """mesh conversion: straight to the kernel when no gradient can be needed."""
    if isinstance(x, torch.Tensor) and not (torch.is_grad_enabled() and x.requires_grad):
        return _resample_forward(x, hout, wout, mode, align)[0]
    return _ResampleFn.apply(x, hout, wout, mode, align)


class _ResampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, hout, wout, mode, align):
        out, ctx.meta = _resample_forward(x, hout, wout, mode, align)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = L.load()
        planes, hin, win, hout, wout, mode, align, shape = ctx.meta
        g = L.dev(g, "mesh_convertor grad")
        gin = torch.empty(shape, dtype=torch.float32, device=g.device)
        with L.on_device(g):
            L.check(lib.ppnn_resample_bwd(g.data_ptr(), gin.data_ptr(), planes, hin, win, hout, wout,
                                          mode, align, L.stream()), "ppnn_resample_bwd")
        return gin, None, None, None, None


class mesh_convertor(object):
    '''
        Class for fine/coarse mesh conversion (reference utils.py:13-31):
        ``down`` = (bi)linear to ``cmesh_size``, ``up`` = bicubic to ``fmesh_size``.
    '''

    def __init__(self, fmesh_size: Size, cmesh_size: Size, dim=1, align_corners=True) -> None:
        super().__init__()
        self.align_corners = align_corners
        self.fmesh_size = fmesh_size
        self.cmesh_size = cmesh_size
        self.dim = dim
        self.down = self._down1d if dim == 1 else self._downnd
        self.up = self.up1d if dim == 1 else self.upnd

    # -- down: interpolate(mode='linear' | 'bilinear', size=cmesh) ------------
    def _down1d(self, u: torch.Tensor) -> torch.Tensor:
        if u.dim() != 3:
            raise RuntimeError(f"dim=1 mesh_convertor expects [B,C,L], got {tuple(u.shape)}")
        size = self.cmesh_size[0] if isinstance(self.cmesh_size, (tuple, list)) else self.cmesh_size
        return _resample(u.unsqueeze(2), 1, int(size), L.BILINEAR, self.align_corners).squeeze(2)

    def _downnd(self, u: torch.Tensor) -> torch.Tensor:
        if u.dim() != 4:
            raise RuntimeError(f"dim=2 mesh_convertor expects [B,C,H,W], got {tuple(u.shape)}")
        h, w = _hw(self.cmesh_size)
        return _resample(u, h, w, L.BILINEAR, self.align_corners)

    # -- up: interpolate(mode='bicubic', size=fmesh) ---------------------------
    def up1d(self, u: torch.Tensor) -> torch.Tensor:
        # the reference resamples a fake [B,1,C,L] image to (fmesh, fmesh) and keeps row 0
        # (utils.py:27-28); only the last dimension carries information
        if u.dim() != 3:
            raise RuntimeError(f"dim=1 mesh_convertor expects [B,C,L], got {tuple(u.shape)}")
        _, w = _hw(self.fmesh_size)
        if u.shape[1] != 1:
            # with C > 1 rows the reference interpolates across channels too; keep that behaviour
            h, _ = _hw(self.fmesh_size)
            return _resample(u.unsqueeze(1), h, w, L.BICUBIC, self.align_corners)[:, :, 0]
        return _resample(u.unsqueeze(1), 1, w, L.BICUBIC, self.align_corners)[:, :, 0]

    def upnd(self, u: torch.Tensor) -> torch.Tensor:
        if u.dim() != 4:
            raise RuntimeError(f"dim=2 mesh_convertor expects [B,C,H,W], got {tuple(u.shape)}")
        h, w = _hw(self.fmesh_size)
        return _resample(u, h, w, L.BICUBIC, self.align_corners)
