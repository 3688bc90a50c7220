"""ctypes binding of ``libppnn_b200.so`` (the C ABI declared in include/ppnn_b200.h).

Host code stays Python/PyTorch; torch is used for device memory and streams only.
There is NO fallback: if the library is missing, or a tensor is not a contiguous
fp32 CUDA tensor, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PPNN_B200_LIB") or os.path.join(_HERE, "libppnn_b200.so")
ABI_VERSION = 4
E_NULL, E_SHAPE, E_TOOBIG = -1, -2, -5   # status codes of include/ppnn_b200.h used by the host code

# enums of include/ppnn_b200.h
CENTRAL1, CENTRAL2, CENTRAL4, FORWARD1, BACKWARD1, UPWIND1 = range(6)
RHS_ADV, RHS_DIFF = 1, 2
REACT_NONE, REACT_BFU, REACT_BFV, REACT_RDU, REACT_RDV = 0, 4, 8, 12, 16
BILINEAR, BICUBIC = 0, 1
FINE_AS_WRITTEN, FINE_INTENDED = 0, 1


class CoarseCfg(C.Structure):
    """``ppnn_coarse_cfg`` (include/ppnn_b200.h)."""
    _fields_ = [("FH", C.c_int), ("FW", C.c_int), ("CH", C.c_int), ("CW", C.c_int),
                ("rhs_u", C.c_int), ("rhs_v", C.c_int),
                ("adv_accuracy", C.c_int), ("diff_accuracy", C.c_int),
                ("dt", C.c_float), ("dx", C.c_float), ("dy", C.c_float),
                ("dx2", C.c_float), ("dy2", C.c_float)]


_P = C.c_void_p
_SIGNATURES = {
    "ppnn_abi_version": (C.c_int, []),
    "ppnn_status_string": (C.c_char_p, [C.c_int]),
    "ppnn_stencil_fwd": (C.c_int, [_P, _P, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ppnn_stencil_bwd": (C.c_int, [_P, _P, _P, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "ppnn_rhs2d_fwd": (C.c_int, [_P] * 6 + [C.c_int] * 9 + [C.c_float] * 4 + [_P]),
    "ppnn_rhs2d_fwd_strided": (C.c_int, [_P] * 6 + [C.c_int] * 9 + [C.c_float] * 4 + [C.c_long, C.c_int, _P]),
    "ppnn_resample_fwd": (C.c_int, [_P, _P, C.c_long] + [C.c_int] * 6 + [_P]),
    "ppnn_resample_fwd_strided": (C.c_int, [_P, _P, C.c_long, C.c_long] + [C.c_int] * 6 + [_P]),
    "ppnn_resample_bwd": (C.c_int, [_P, _P, C.c_long] + [C.c_int] * 6 + [_P]),
    "ppnn_pde_du_fwd": (C.c_int, [C.POINTER(CoarseCfg), _P, _P, _P, C.c_int, _P]),
    "ppnn_pde_du_iter_fwd": (C.c_int, [C.POINTER(CoarseCfg), _P, _P, _P, C.c_int, C.c_int, _P]),
    "ppnn_pde_du_bwd": (C.c_int, [C.POINTER(CoarseCfg), _P, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "ppnn_step_update_bwd": (C.c_int, [C.POINTER(CoarseCfg), _P, _P, _P, C.POINTER(C.c_float), _P, _P, _P, C.c_int, _P]),
    "ppnn_pde_du_fwd_save": (C.c_int, [C.POINTER(CoarseCfg), _P, _P, _P, _P, C.c_int, _P]),
    "ppnn_step_update_save": (C.c_int, [C.POINTER(CoarseCfg), _P, _P, _P,
                                        C.POINTER(C.c_float), C.POINTER(C.c_float), _P, _P, C.c_int, _P]),
    "ppnn_pde_du_bwd_saved": (C.c_int, [C.POINTER(CoarseCfg), _P, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "ppnn_step_update_bwd_saved": (C.c_int, [C.POINTER(CoarseCfg), _P, _P, _P, C.POINTER(C.c_float), _P, _P, _P, C.c_int, _P]),
    "ppnn_step_inputs": (C.c_int, [C.POINTER(CoarseCfg), _P, _P,
                                   C.POINTER(C.c_float), C.POINTER(C.c_float),
                                   C.POINTER(C.c_float), C.POINTER(C.c_float),
                                   _P, _P, _P, C.c_int, _P]),
    "ppnn_step_update": (C.c_int, [C.POINTER(CoarseCfg), _P, _P, _P,
                                   C.POINTER(C.c_float), C.POINTER(C.c_float), _P, C.c_int, _P]),
    "ppnn_coarse_increment": (C.c_int, [C.POINTER(CoarseCfg), _P, _P, _P, C.c_int, C.c_int, _P]),
    "ppnn_dataset_stats": (C.c_int, [C.POINTER(CoarseCfg), _P, _P, _P, C.c_long, C.c_long, _P, _P]),
    "ppnn_dataset_emit": (C.c_int, [C.POINTER(CoarseCfg), _P, _P, _P, C.c_long, C.c_long,
                                    C.POINTER(C.c_float), C.POINTER(C.c_float), _P, _P, _P, _P]),
    "ppnn_fine_burgers_step": (C.c_int, [_P] * 5 + [C.c_int] * 7 + [C.c_float] * 3 + [_P]),
    "ppnn_fourier_ic": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P]),
    "ppnn_fine_rhs_step": (C.c_int, [_P] * 5 + [C.c_int] * 8 + [C.c_float] * 3 + [_P]),
    "ppnn_ks1d_rk4": (C.c_int, [_P] * 3 + [C.c_int] * 7 + [C.c_float] * 4 + [_P]),
    "ppnn_ks1d_rhs": (C.c_int, [_P] * 3 + [C.c_int] * 6 + [C.c_float] * 3 + [_P]),
    "ppnn_ks2d_rhs": (C.c_int, [_P] * 2 + [C.c_int] * 7 + [C.c_float] * 8 + [_P]),
    "ppnn_ks2d_rk4": (C.c_int, [_P] * 2 + [C.c_int] * 8 + [C.c_float] * 9 + [_P]),
    "ppnn_heat1d_pad": (C.c_int, [_P] * 4 + [C.c_int] * 2 + [C.c_float, _P]),
    "ppnn_heat1d_update": (C.c_int, [_P] * 6 + [C.c_int] * 2 + [C.c_float, _P]),
    "ppnn_heat1d_bc_steps": (C.c_int, [_P] * 5 + [C.c_long] + [C.c_int] * 3 + [C.c_float] * 3 + [_P]),
    "ppnn_heat1d_coarse_increment": (C.c_int, [_P] * 4 + [C.c_int] * 3 + [C.c_float] * 3 + [_P]),
}
EXPORTS = tuple(_SIGNATURES)

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library once and bind every declared entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  ppnn_b200 has no CPU or eager fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.restype, fn.argtypes = res, args
    if lib.ppnn_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI {lib.ppnn_abi_version()} != expected {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().ppnn_status_string(status).decode()
        raise RuntimeError(f"{what}: status {status} ({msg})")


def dev(t: torch.Tensor, what: str = "tensor") -> torch.Tensor:
    """Validate a tensor for the C ABI (contiguous fp32 on a CUDA device)."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{what}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{what}: ppnn_b200 runs on CUDA only (got device {t.device}); there is no CPU fallback")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{what}: expected float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


class _NoCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NOCTX = _NoCtx()


# torch.cuda.current_device() / current_stream() are Python-level and cost ~1.5 / ~14 us per call -- more than
# the launch they precede at rollout batch sizes (tools/profile_tier_host.py); the C bindings behind them do not.
_cuda_get_device = getattr(torch._C, "_cuda_getDevice", None)
_cuda_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _current_device() -> int:
    return _cuda_get_device() if _cuda_get_device is not None else torch.cuda.current_device()


def on_device(t: torch.Tensor):
    """Context that makes ``t``'s device current -- free when it already is."""
    if t.device.index == _current_device():
        return _NOCTX
    return torch.cuda.device(t.device)


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream() -> int:
    """Raw handle of the current stream of the current device (what every C entry point takes)."""
    if _cuda_raw_stream is not None:
        return _cuda_raw_stream(_current_device())
    return torch.cuda.current_stream().cuda_stream


def f2(x) -> "C.Array":
    """Per-channel statistic (tensor [1,2,1,1] / sequence / scalar) -> float[2]."""
    if isinstance(x, torch.Tensor):
        v = x.detach().reshape(-1).tolist()
    elif isinstance(x, (int, float)):
        v = [float(x)]
    else:
        v = [float(a) for a in x]
    if len(v) == 1:
        v = v * 2
    if len(v) != 2:
        raise ValueError(f"expected 2 per-channel values, got {len(v)}")
    return (C.c_float * 2)(*v)
