"""Drop-in replacement for the operator classes of the reference ``src/operators.py``.

Same class names, constructor signatures, defaults, assertions and call
convention (``op(u)`` on a PRE-PADDED single-channel tensor, 'valid' output, no
1/dx scaling -- reference operators.py:74-128, 228-273), but every call is one
hand-written sm_100a kernel behind the C ABI (``ppnn_stencil_fwd`` /
``ppnn_stencil_bwd``) instead of ``conv2d`` + mask arithmetic.  Differentiable
with respect to ``u`` (the masks of the upwind schemes are constants, exactly as
autograd treats the reference's comparisons).

CUDA only: a CPU tensor raises (there is no fallback path).
"""
from __future__ import annotations

import torch

from . import _lib as L

__all__ = ["operator_base1D", "diffusion1D", "advection1d", "dudx_1D",
           "operator_base2D", "d2udx2_2D", "d2udy2_2D", "dudx_2D", "dudy_2D"]

# scheme name -> (ABI id, accuracies the reference tables hold)
_SCHEMES_2D = {
    "Central1": (L.CENTRAL1, (2, 4)),           # operators.py:150-154
    "Central2": (L.CENTRAL2, (2, 4, 6, 8)),     # operators.py:156-168
    "Central4": (L.CENTRAL4, (2, 4, 6)),        # operators.py:169-175
    "Forward1": (L.FORWARD1, (1, 2, 3)),        # operators.py:176-183
    "Backward1": (L.BACKWARD1, (1, 2, 3)),      # operators.py:185-192
    "Upwind1": (L.UPWIND1, (1, 2, 3)),          # operators.py:211-218
}
_SCHEMES_1D = {
    "Central": _SCHEMES_2D["Central2"],          # operators.py:17-30
    "Central_4th": _SCHEMES_2D["Central4"],      # operators.py:32-38
    "Forward": _SCHEMES_2D["Forward1"],
    "Backward": _SCHEMES_2D["Backward1"],
    "Upwind": _SCHEMES_2D["Upwind1"],            # operators.py:102-119
}


def _taps(scheme_id: int, accuracy: int) -> int:
    if scheme_id in (L.CENTRAL1,):
        return accuracy + 1
    if scheme_id == L.CENTRAL2:
        return accuracy + 1
    if scheme_id == L.CENTRAL4:
        return accuracy + 3
    if scheme_id == L.UPWIND1:
        return 2 * accuracy + 1
    return accuracy + 1                          # Forward1 / Backward1


def _stencil_forward(u, scheme_id, accuracy, axis):
    """Launch ``ppnn_stencil_fwd``; returns (output, contiguous input, (planes, hin, win))."""
    lib = L.load()
    x = L.dev(u, "operator input")
    shape = x.shape
    if x.dim() == 3:                         # 1-D operators: [B, 1, L]
        if axis != 3:
            raise RuntimeError("1-D operators act along the last dimension")
        hin, win = 1, shape[-1]
    else:
        hin, win = shape[-2], shape[-1]
    planes = x.numel() // (hin * win)
    k = _taps(scheme_id, accuracy)
    hout, wout = (hin - k + 1, win) if axis == 2 else (hin, win - k + 1)
    if hout < 1 or wout < 1:
        raise RuntimeError(f"input of shape {tuple(shape)} is smaller than the {k}-tap stencil")
    out_shape = list(shape)
    out_shape[-1] = wout
    if axis == 2:
        out_shape[-2] = hout
    out = torch.empty(out_shape, dtype=torch.float32, device=x.device)
    with L.on_device(x):
        L.check(lib.ppnn_stencil_fwd(x.data_ptr(), out.data_ptr(), planes, hin, win, axis,
                                     scheme_id, accuracy, L.stream()), "ppnn_stencil_fwd")
    return out, x, (planes, hin, win)


class _StencilFn(torch.autograd.Function):
    """y = valid cross-correlation of u with the (scheme, accuracy) filter along ``axis``."""

    @staticmethod
    def forward(ctx, u, scheme_id, accuracy, axis):
        out, x, (planes, hin, win) = _stencil_forward(u, scheme_id, accuracy, axis)
        ctx.meta = (scheme_id, accuracy, axis, planes, hin, win)
        ctx.save_for_backward(x if scheme_id == L.UPWIND1 else None)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = L.load()
        scheme_id, accuracy, axis, planes, hin, win = ctx.meta
        (x,) = ctx.saved_tensors
        g = L.dev(g, "operator grad")
        shape = list(g.shape)
        shape[-1] = win
        if axis == 2:
            shape[-2] = hin
        gin = torch.empty(shape, dtype=torch.float32, device=g.device)
        with L.on_device(g):
            L.check(lib.ppnn_stencil_bwd(g.data_ptr(), L.ptr(x), gin.data_ptr(), planes, hin, win, axis,
                                         scheme_id, accuracy, L.stream()), "ppnn_stencil_bwd")
        return gin, None, None, None


def _stencil(u, scheme_id, accuracy, axis):
    """Operator call: straight to the kernel when no gradient can be needed."""
    if isinstance(u, torch.Tensor) and not (torch.is_grad_enabled() and u.requires_grad):
        return _stencil_forward(u, scheme_id, accuracy, axis)[0]
    return _StencilFn.apply(u, scheme_id, accuracy, axis)


def _resolve(table, scheme, accuracy):
    if scheme not in table:
        raise KeyError(scheme)
    sid, accs = table[scheme]
    if accuracy not in accs:
        # the reference indexes a list holding None at the missing accuracies
        raise IndexError(f"scheme {scheme!r} has no filter at accuracy {accuracy}")
    return sid


# --------------------------------------------------------------------------- 1-D
class operator_base1D(object):
    """Base of the 1-D operators (reference operators.py:6-71)."""

    def __init__(self, accuracy, device='cpu') -> None:
        super().__init__()
        self.accuracy: int = accuracy
        self.device: torch.device = torch.device(device)
        self.schemes = dict(_SCHEMES_1D)
        self._scheme_id = None

    def __call__(self, u: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError


class diffusion1D(operator_base1D):
    """2nd / 4th derivative numerator on [B,1,L+2p] (reference operators.py:74-85)."""

    def __init__(self, scheme='Central', accuracy=2, device='cpu') -> None:
        super().__init__(accuracy, device)
        assert accuracy % 2 == 0, 'diffusion operator precision must be even number'
        assert scheme == 'Central' or scheme == 'Central_4th', \
            'scheme must be one of the following: Central, Central_4th'
        self.scheme = scheme
        self._scheme_id = _resolve(_SCHEMES_1D, scheme, accuracy)

    def __call__(self, u: torch.Tensor) -> torch.Tensor:
        return _stencil(u, self._scheme_id, self.accuracy, 3)


class advection1d(operator_base1D):
    """Present in the reference only as a stub (operators.py:89-96)."""

    def __init__(self, scheme='Central', accuracy=2, device='cpu') -> None:
        super().__init__(accuracy, device)
        self.scheme = scheme
        self.filters = self.schemes[scheme]

    def __call__(self, u: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError


class dudx_1D(operator_base1D):
    """Sign-switched one-sided 1st derivative, periodic use only (operators.py:99-128)."""

    def __init__(self, scheme='Upwind', accuracy=2, device='cpu') -> None:
        super().__init__(accuracy, device)
        self.scheme = scheme
        self._scheme_id = _resolve(_SCHEMES_1D, scheme, accuracy)

    def __call__(self, u: torch.Tensor) -> torch.Tensor:
        if self.scheme != 'Upwind':
            # the reference indexes the filter with ['forward'] and fails for other schemes
            raise TypeError("dudx_1D supports scheme='Upwind' only")
        return _stencil(u, self._scheme_id, self.accuracy, 3)


# --------------------------------------------------------------------------- 2-D
class operator_base2D(object):
    """Base of the 2-D operators (reference operators.py:145-225)."""

    def __init__(self, accuracy=2, device='cpu') -> None:
        self.accuracy = accuracy
        self.device = device
        self.xschemes = dict(_SCHEMES_2D)
        self.yschemes = dict(_SCHEMES_2D)
        self.scheme = None
        self._scheme_id = None
        self._axis = None

    def __call__(self, u: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def _apply(self, u: torch.Tensor) -> torch.Tensor:
        if u.dim() != 4 or u.shape[1] != 1:
            raise RuntimeError(f"expected a [B,1,H,W] tensor, got {tuple(u.shape)}")
        return _stencil(u, self._scheme_id, self.accuracy, self._axis)


def _assert_second_derivative(scheme, accuracy):
    assert accuracy % 2 == 0, 'diffusion operator precision must be even number'
    assert scheme == 'Central2' or scheme == 'Central4' or scheme == 'Forward1' or scheme == 'Backward1', \
        'scheme must be one of the following: Central2, Central4, Forward1, Backward1'


class d2udx2_2D(operator_base2D):
    """reference operators.py:228-237 (along dim 2)."""

    def __init__(self, scheme='Central2', accuracy=2, device='cpu') -> None:
        super().__init__(accuracy, device)
        _assert_second_derivative(scheme, accuracy)
        self.scheme, self._axis = scheme, 2
        self._scheme_id = _resolve(_SCHEMES_2D, scheme, accuracy)

    def __call__(self, u: torch.Tensor) -> torch.Tensor:
        return self._apply(u)


class d2udy2_2D(operator_base2D):
    """reference operators.py:239-247 (along dim 3)."""

    def __init__(self, scheme='Central2', accuracy=2, device='cpu') -> None:
        super().__init__(accuracy, device)
        _assert_second_derivative(scheme, accuracy)
        self.scheme, self._axis = scheme, 3
        self._scheme_id = _resolve(_SCHEMES_2D, scheme, accuracy)

    def __call__(self, u: torch.Tensor) -> torch.Tensor:
        return self._apply(u)


class dudx_2D(operator_base2D):
    """reference operators.py:249-260: Upwind1 switches on the sign of ``u`` itself."""

    def __init__(self, scheme='Upwind1', accuracy=1, device='cpu') -> None:
        super().__init__(accuracy, device)
        self.xscheme = self.scheme = scheme
        self._axis = 2
        self._scheme_id = _resolve(_SCHEMES_2D, scheme, accuracy)

    def __call__(self, u: torch.Tensor) -> torch.Tensor:
        return self._apply(u)


class dudy_2D(operator_base2D):
    """reference operators.py:262-273."""

    def __init__(self, scheme='Upwind1', accuracy=1, device='cpu') -> None:
        super().__init__(accuracy, device)
        self.yscheme = self.scheme = scheme
        self._axis = 3
        self._scheme_id = _resolve(_SCHEMES_2D, scheme, accuracy)

    def __call__(self, u: torch.Tensor) -> torch.Tensor:
        return self._apply(u)
