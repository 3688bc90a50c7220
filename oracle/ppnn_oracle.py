"""CPU oracle for the PPNN PDE-preserving hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, on the CPU and with the same ATen operator sequence
(``conv2d`` / ``conv1d`` / ``interpolate`` / ``cat`` / element-wise), what the
reference computes on the path named in SURVEY.md section 8.  It exists so the CUDA
product path can be checked and so ``bench.py`` has a CPU baseline to time.

Nothing under ``ppnn_b200/`` may import it.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py``.

Parity pinning: the reference ships no tests or golden vectors (SURVEY.md
section 4), so this oracle is pinned against outputs of the *reference code itself*,
generated in the build container by ``tests/golden/make_golden.py`` (which
imports ``/root/reference/src``) and committed as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` compares them.

Reference lines followed (paths relative to the reference checkout):

* stencil coefficient tables ........ src/operators.py:16-56, 102-119, 150-218
* 'valid' cross-correlation ops ...... src/operators.py:85, 126-127, 237, 247
* sign-switched upwind ............... src/operators.py:255-260, 268-273
* right-hand sides ................... src/rhs.py:1-56
* mesh converter (bilinear / bicubic)  src/utility/utils.py:13-31
* coarse periodic pads, endpoint ..... src/train2D.py:38-47
* pde_du ............................. src/train2D.py:66-79
* rollout update ..................... src/train2D.py:230-245
* fine-grid pads and Euler loop ...... src/operators.py:416-420, 466-475
* PDE-only fine rollout .............. post_process/test_predict_2Dp.py:116
* coarse-only solver ................. post_process/test_2D_csolver.py:56-75
* "iter" variant ..................... post_process/test2D2.py:70-91
* 2-D / 1-D Kuramoto-Sivashinsky RK4 . src/operators.py:750-843, 845-918 (commented generators)
* 1-D heat step with ghost cells ..... Bv/data.py:9-51, 92-97

All functions work in whatever dtype their tensor arguments carry (fp32 for
parity, fp64 for the noise-floor reference).
"""
from __future__ import annotations

from fractions import Fraction as Fr
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------
# Coefficient tables (exact rationals; rounded to the working dtype on use).
# --------------------------------------------------------------------------
_CENTRAL = {
    # first derivative                                   operators.py:150-154
    ("Central1", 2): [Fr(-1, 2), Fr(0), Fr(1, 2)],
    ("Central1", 4): [Fr(1, 12), Fr(-2, 3), Fr(0), Fr(2, 3), Fr(-1, 12)],
    # second derivative                                  operators.py:156-168
    ("Central2", 2): [Fr(1), Fr(-2), Fr(1)],
    ("Central2", 4): [Fr(-1, 12), Fr(4, 3), Fr(-5, 2), Fr(4, 3), Fr(-1, 12)],
    ("Central2", 6): [Fr(1, 90), Fr(-3, 20), Fr(3, 2), Fr(-49, 18), Fr(3, 2), Fr(-3, 20), Fr(1, 90)],
    ("Central2", 8): [Fr(-1, 560), Fr(8, 315), Fr(-1, 5), Fr(8, 5), Fr(-205, 72),
                      Fr(8, 5), Fr(-1, 5), Fr(8, 315), Fr(-1, 560)],
    # fourth derivative                                  operators.py:169-175
    ("Central4", 2): [Fr(1), Fr(-4), Fr(6), Fr(-4), Fr(1)],
    ("Central4", 4): [Fr(-1, 6), Fr(2), Fr(-13, 2), Fr(28, 3), Fr(-13, 2), Fr(2), Fr(-1, 6)],
    ("Central4", 6): [Fr(7, 240), Fr(-2, 5), Fr(169, 60), Fr(-122, 15), Fr(91, 8),
                      Fr(-122, 15), Fr(169, 60), Fr(-2, 5), Fr(7, 240)],
}
_ONE_SIDED = {
    # operators.py:176-192
    ("Forward1", 1): [Fr(-1), Fr(1)],
    ("Forward1", 2): [Fr(-3, 2), Fr(2), Fr(-1, 2)],
    ("Forward1", 3): [Fr(-11, 6), Fr(3), Fr(-3, 2), Fr(1, 3)],
    ("Backward1", 1): [Fr(-1), Fr(1)],
    ("Backward1", 2): [Fr(1, 2), Fr(-2), Fr(3, 2)],
    ("Backward1", 3): [Fr(-1, 3), Fr(3, 2), Fr(-3), Fr(11, 6)],
}
# 1-D classes use other scheme names for the same numbers (operators.py:58-61)
_ALIASES_1D = {"Central": "Central2", "Central_4th": "Central4",
               "Forward": "Forward1", "Backward": "Backward1"}


def taps(scheme: str, accuracy: int) -> List[float]:
    """Coefficients of ``scheme`` at ``accuracy`` as Python doubles (the
    reference builds fp32 tensors from Python doubles, operators.py:150-192)."""
    scheme = _ALIASES_1D.get(scheme, scheme)
    key = (scheme, accuracy)
    table = _CENTRAL if key in _CENTRAL else _ONE_SIDED
    if key not in table:
        raise KeyError(f"no {scheme} filter at accuracy {accuracy}")
    return [float(c) for c in table[key]]


def upwind_taps(accuracy: int) -> Tuple[List[float], List[float]]:
    """(forward, backward) filters zero-padded to 2a+1 taps (operators.py:211-218)."""
    fwd = [0.0] * accuracy + taps("Forward1", accuracy)
    bwd = taps("Backward1", accuracy) + [0.0] * accuracy
    return fwd, bwd


# --------------------------------------------------------------------------
# Operators on pre-padded input (Tier A semantics).
# --------------------------------------------------------------------------
_KERNEL_CACHE: Dict[tuple, torch.Tensor] = {}


def _kernel(coeffs: Sequence[float], like: torch.Tensor, axis: int) -> torch.Tensor:
    """Filter tensor on the device / in the dtype of ``like`` (built once per combination,
    as the reference builds its filters once per operator object, operators.py:145-222)."""
    key = (tuple(coeffs), like.dtype, str(like.device), axis)
    k = _KERNEL_CACHE.get(key)
    if k is None:
        k = torch.tensor(list(coeffs), dtype=torch.float32).to(device=like.device, dtype=like.dtype)
        k = k.reshape(1, 1, -1, 1) if axis == 2 else k.reshape(1, 1, 1, -1)
        _KERNEL_CACHE[key] = k
    return k


def correlate2d(u: torch.Tensor, coeffs: Sequence[float], axis: int) -> torch.Tensor:
    """'valid' cross-correlation of [B,1,H,W] along ``axis`` (2 = x, 3 = y)."""
    return F.conv2d(u, _kernel(coeffs, u, axis))


def correlate1d(u: torch.Tensor, coeffs: Sequence[float]) -> torch.Tensor:
    k = torch.tensor(list(coeffs), dtype=torch.float32).to(device=u.device, dtype=u.dtype).reshape(1, 1, -1)
    return F.conv1d(u, k)


def stencil2d(u: torch.Tensor, scheme: str, accuracy: int, axis: int) -> torch.Tensor:
    """d2udx2_2D / d2udy2_2D / non-upwind dudx_2D / dudy_2D (operators.py:228-273)."""
    if scheme == "Upwind1":
        return upwind2d(u, accuracy, axis)
    return correlate2d(u, taps(scheme, accuracy), axis)


def upwind2d(u: torch.Tensor, accuracy: int, axis: int) -> torch.Tensor:
    """Sign-switched one-sided derivative; the switch is the sign of the field
    itself at the output point (operators.py:255-258, 268-271)."""
    a = accuracy
    fwd, bwd = upwind_taps(a)
    centre = u[:, :, a:-a] if axis == 2 else u[:, :, :, a:-a]
    return (centre <= 0) * correlate2d(u, fwd, axis) + (centre > 0) * correlate2d(u, bwd, axis)


def stencil1d(u: torch.Tensor, scheme: str, accuracy: int) -> torch.Tensor:
    """diffusion1D / dudx_1D on [B,1,L+2p] (operators.py:74-128)."""
    if scheme == "Upwind":
        a = accuracy
        fwd, bwd = upwind_taps(a)
        centre = u[:, :, a:-a]
        return (centre <= 0) * correlate1d(u, fwd) + (centre > 0) * correlate1d(u, bwd)
    return correlate1d(u, taps(scheme, accuracy))


# --------------------------------------------------------------------------
# Mesh conversion (utils.py:13-31).
# --------------------------------------------------------------------------
def mesh_down(u: torch.Tensor, size, align_corners: bool = True) -> torch.Tensor:
    mode = "linear" if u.dim() == 3 else "bilinear"
    return F.interpolate(u, size=size, mode=mode, align_corners=align_corners)


def mesh_up(u: torch.Tensor, size, align_corners: bool = True) -> torch.Tensor:
    if u.dim() == 3:                                   # utils.py:27-28
        return F.interpolate(u.unsqueeze(1), size=size, mode="bicubic",
                             align_corners=align_corners)[:, :, 0]
    return F.interpolate(u, size=size, mode="bicubic", align_corners=align_corners)


# --------------------------------------------------------------------------
# Pads.
# --------------------------------------------------------------------------
def wrap_x(u: torch.Tensor, h: int = 1) -> torch.Tensor:
    """True periodic halo of ``h`` rows (train2D.py:38-39 has h = 1)."""
    return torch.cat((u[:, :, -h:], u, u[:, :, :h]), dim=2)


def wrap_y(u: torch.Tensor, h: int = 1) -> torch.Tensor:
    return torch.cat((u[:, :, :, -h:], u, u[:, :, :, :h]), dim=3)


def append_endpoint(u: torch.Tensor) -> torch.Tensor:
    """padBC_rd: re-append first column then first row (train2D.py:45-47)."""
    t = torch.cat((u, u[:, :, :, :1]), dim=3)
    return torch.cat((t, t[:, :, :1]), dim=2)


def fine_wrap_x(u: torch.Tensor, h: int = 3) -> torch.Tensor:
    """Halo for a grid that carries a duplicated endpoint (operators.py:416-417)."""
    return torch.cat((u[:, :, -h - 1:-1], u, u[:, :, 1:h + 1]), dim=2)


def fine_wrap_y(u: torch.Tensor, h: int = 3) -> torch.Tensor:
    return torch.cat((u[:, :, :, -h - 1:-1], u, u[:, :, :, 1:h + 1]), dim=3)


# --------------------------------------------------------------------------
# Right-hand sides (rhs.py).  ``ops`` carries the four operator callables.
# --------------------------------------------------------------------------
class Ops:
    """The four operator objects a right-hand side receives (train2D.py:53-56
    builds them with accuracy 2 / 1; operators.py:450-455 with 6 / 3)."""

    def __init__(self, adv_accuracy: int = 1, diff_accuracy: int = 2,
                 adv_scheme: str = "Upwind1", diff_scheme: str = "Central2"):
        self.adv_accuracy, self.diff_accuracy = adv_accuracy, diff_accuracy
        self.adv_halo = adv_accuracy if adv_scheme == "Upwind1" else adv_accuracy // 2
        self.diff_halo = (len(taps(diff_scheme, diff_accuracy)) - 1) // 2
        self.dudx = lambda a: stencil2d(a, adv_scheme, adv_accuracy, 2)
        self.dudy = lambda a: stencil2d(a, adv_scheme, adv_accuracy, 3)
        self.d2udx2 = lambda a: stencil2d(a, diff_scheme, diff_accuracy, 2)
        self.d2udy2 = lambda a: stencil2d(a, diff_scheme, diff_accuracy, 3)


def _advect(u1, v1, o: Ops, ax, ay, dx, dy):
    return -u1 * o.dudx(ax) / dx - v1 * o.dudy(ay) / dy            # rhs.py:3


def _diffuse(mu, o: Ops, ax, ay, dx2, dy2):
    return mu * (o.d2udx2(ax) / dx2 + o.d2udy2(ay) / dy2)          # rhs.py:19


def _react_u(u1, v1):
    return u1 - u1 * u1 * u1 - v1 + 0.01                           # rhs.py:26


def _react_v(u1, v1):
    return 0.25 * (u1 - v1)                                        # rhs.py:29


def rhs(name: str, u1, v1, mu, o: Ops, ax, ay, dx, dy, dx2, dy2) -> torch.Tensor:
    """Evaluate reference RHS ``name`` (any of the 2-D functions in rhs.py)."""
    if name in ("burgers2Dpu", "burgers2Dpv"):                     # rhs.py:9-15
        return _advect(u1, v1, o, ax, ay, dx, dy) + _diffuse(mu, o, ax, ay, dx2, dy2)
    if name == "burgers2DFu":                                      # rhs.py:1-3
        uv = u1 * u1 + v1 * v1
        return (_advect(u1, v1, o, ax, ay, dx, dy) + _diffuse(mu, o, ax, ay, dx2, dy2)
                + (1 - uv) * u1 + uv * v1)
    if name == "burgers2DFv":                                      # rhs.py:5-7
        uv = u1 * u1 + v1 * v1
        return (_advect(u1, v1, o, ax, ay, dx, dy) + _diffuse(mu, o, ax, ay, dx2, dy2)
                - uv * u1 + (1 - uv) * v1)
    if name in ("burgers2Ddiffuseu", "burgers2Ddiffusev", "rdpu", "rdpv", "diffu", "diffv"):
        return _diffuse(mu, o, ax, ay, dx2, dy2)                   # rhs.py:17-23,31-41
    if name == "rdFu":                                             # rhs.py:25-26
        return _diffuse(mu, o, ax, ay, dx2, dy2) + u1 - u1 * u1 * u1 - v1 + 0.01
    if name == "rdFv":                                             # rhs.py:28-29
        return _diffuse(mu, o, ax, ay, dx2, dy2) + 0.25 * (u1 - v1)
    if name in ("convecu", "convecv"):                             # rhs.py:43-47
        return _advect(u1, v1, o, ax, ay, dx, dy)
    if name == "burgers2DWu":                                      # rhs.py:52-53
        return _react_u(u1, v1)
    if name == "burgers2DWv":                                      # rhs.py:55-56
        return _react_v(u1, v1)
    raise KeyError(name)


RHS_NAMES_2D = ("burgers2DFu", "burgers2DFv", "burgers2Dpu", "burgers2Dpv",
                "burgers2Ddiffuseu", "burgers2Ddiffusev", "rdFu", "rdFv", "rdpu", "rdpv",
                "diffu", "diffv", "convecu", "convecv", "burgers2DWu", "burgers2DWv")


def ks1d(ux4, mu, ux, dx, dx2, dx4, adv_acc: int = 1, d2_acc: int = 2, d4_acc: int = 2):
    """1-D Kuramoto-Sivashinsky RHS (rhs.py:49-50); ``ux4`` carries the wider
    halo of the 4th-derivative stencil, ``ux`` the halo of the other two."""
    return (-stencil1d(ux, "Central", d2_acc) / dx2 - stencil1d(ux4, "Central_4th", d4_acc) / dx4
            - 0.5 * stencil1d(ux * ux, "Upwind", adv_acc) / dx)


# --------------------------------------------------------------------------
# PPNN coarse step (train2D.py:49-79) and its callers.
# --------------------------------------------------------------------------
class CoarseStep:
    """``pde_du`` of one YAML case: F fine points, C coarse points (both with
    the duplicated periodic endpoint), RHS pair, dt, domain length."""

    def __init__(self, fine: int, coarse: int, rhsu: str, rhsv: str, dt: float, length: float):
        self.F, self.C = fine, coarse
        self.rhsu, self.rhsv = rhsu, rhsv
        self.dt = dt
        self.dx = length / (coarse - 1)                            # train2D.py:58
        self.dx2 = self.dx ** 2
        self.dy, self.dy2 = self.dx, self.dx2                      # train2D.py:61
        self.ops = Ops(adv_accuracy=1, diff_accuracy=2)            # train2D.py:53-56

    # coarse state without the endpoint, train2D.py:67-68
    def restrict(self, u: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        u1 = mesh_down(u[:, :1], self.C)[:, :, :-1, :-1]
        v1 = mesh_down(u[:, 1:], self.C)[:, :, :-1, :-1]
        return u1, v1

    def coarse_increment(self, u1, v1, mu) -> Tuple[torch.Tensor, torch.Tensor]:
        """dt * rhs on the N x N periodic coarse grid (train2D.py:69-77)."""
        o = self.ops
        ux, uy, vx, vy = wrap_x(u1), wrap_y(u1), wrap_x(v1), wrap_y(v1)
        ru = self.dt * rhs(self.rhsu, u1, v1, mu, o, ux, uy, self.dx, self.dy, self.dx2, self.dy2)
        rv = self.dt * rhs(self.rhsv, u1, v1, mu, o, vx, vy, self.dx, self.dy, self.dx2, self.dy2)
        return ru, rv

    def pde_du(self, u: torch.Tensor, mu) -> torch.Tensor:
        """[B,2,F,F], mu -> [B,2,F,F] (train2D.py:66-79)."""
        u1, v1 = self.restrict(u)
        ru, rv = self.coarse_increment(u1, v1, mu)
        return torch.cat((mesh_up(append_endpoint(ru), self.F),
                          mesh_up(append_endpoint(rv), self.F)), dim=1)

    def pde_du_iter(self, u: torch.Tensor, mu, inner: int = 1) -> torch.Tensor:
        """test2D2.py:70-91: ``inner`` sequential coarse Euler sub-steps (v sees
        the updated u) before the increment is evaluated."""
        o = self.ops
        u1, v1 = self.restrict(u)
        for _ in range(inner):
            ux, uy, vx, vy = wrap_x(u1), wrap_y(u1), wrap_x(v1), wrap_y(v1)
            u1 = u1 + self.dt * rhs(self.rhsu, u1, v1, mu, o, ux, uy, self.dx, self.dy, self.dx2, self.dy2)
            v1 = v1 + self.dt * rhs(self.rhsv, u1, v1, mu, o, vx, vy, self.dx, self.dy, self.dx2, self.dy2)
        ru, rv = self.coarse_increment(u1, v1, mu)
        return torch.cat((mesh_up(append_endpoint(ru), self.F),
                          mesh_up(append_endpoint(rv), self.F)), dim=1)

    def coarse_only_increment(self, uc: torch.Tensor, mu) -> torch.Tensor:
        """test_2D_csolver.py:56-66: state lives on the C x C grid; the 'down'
        there is an identity resample (C -> C)."""
        u1 = mesh_down(uc[:, :1], self.C)[:, :, :-1, :-1]
        v1 = mesh_down(uc[:, 1:], self.C)[:, :, :-1, :-1]
        ru, rv = self.coarse_increment(u1, v1, mu)
        return torch.cat((append_endpoint(ru), append_endpoint(rv)), dim=1)


def pde_only_rollout(step: CoarseStep, u: torch.Tensor, mu, nsteps: int) -> List[torch.Tensor]:
    """u <- u + pde_du(u) (post_process/test_predict_2Dp.py:116)."""
    out = []
    for _ in range(nsteps):
        u = u + step.pde_du(u, mu)
        out.append(u)
    return out


def ppnn_update(u, pdeuu, nn_out, outmean, outstd) -> torch.Tensor:
    """train2D.py:235-238 with the network output given:
    ``padBC_rd(nn*outstd + outmean) + u + pdeuu``."""
    return append_endpoint(nn_out * outstd + outmean) + u + pdeuu


def ppnn_net_inputs(u, mu, pdeuu, inmean, instd, mumean, mustd, pdemean, pdestd):
    """The three normalised network inputs of train2D.py:235-237."""
    return ((u[:, :, :-1, :-1] - inmean) / instd,
            (mu - mumean) / mustd,
            (pdeuu[:, :, :-1, :-1] - pdemean) / pdestd)


def ppnn_rollout(step: CoarseStep, net: Callable, u, mu, stats: Dict[str, torch.Tensor],
                 nsteps: int, uses_pde: bool = True) -> List[torch.Tensor]:
    """The every-100-epoch test rollout (train2D.py:230-245)."""
    out = []
    for _ in range(nsteps):
        if uses_pde:
            pdeuu = step.pde_du(u, mu)
            a, b, c = ppnn_net_inputs(u, mu, pdeuu, stats["inmean"], stats["instd"],
                                      stats["mumean"], stats["mustd"],
                                      stats["pdemean"], stats["pdestd"])
            u = ppnn_update(u, pdeuu, net(a, b, c), stats["outmean"], stats["outstd"])
        else:
            a = (u[:, :, :-1, :-1] - stats["inmean"]) / stats["instd"]
            b = (mu - stats["mumean"]) / stats["mustd"]
            u = append_endpoint(net(a, b) * stats["outstd"] + stats["outmean"]) + u
        out.append(u)
    return out


def training_set(step: Optional["CoarseStep"], fdata: torch.Tensor, mus: torch.Tensor) -> Dict[str, torch.Tensor]:
    """The data-set precompute of train2D.py:88-90 and class ``myset`` (:109-151).

    ``fdata`` [P, T+1, 2, F, F] trajectories, ``mus`` [P*T, 1, 1, 1] one parameter per
    sample; ``step`` None = the ``params.pde == False`` branch (:134-139).  Returns the
    arrays and statistics the reference keeps as attributes of the data set."""
    F = fdata.shape[-1]
    data_u0 = fdata[:, :-1].reshape(-1, 2, F, F)                                  # :88
    data_du = (fdata[:, 1:] - fdata[:, :-1]).reshape(-1, 2, F, F)                 # :89-90
    out: Dict[str, torch.Tensor] = {}
    out["inmean"] = data_u0[:, :, :-1, :-1].mean(dim=(0, 2, 3), keepdim=True)     # :115
    out["instd"] = data_u0[:, :, :-1, :-1].std(dim=(0, 2, 3), keepdim=True)       # :116
    out["u0_normd"] = (data_u0[:, :, :-1, :-1] - out["inmean"]) / out["instd"]    # :117
    if step is not None:
        pdeu = step.pde_du(data_u0, mus)                                          # :121
        out["pdemean"] = pdeu.mean(dim=(0, 2, 3), keepdim=True)                   # :126 (full plane)
        out["pdestd"] = pdeu.std(dim=(0, 2, 3), keepdim=True)                     # :127
        out["pdeu_normd"] = (pdeu[:, :, :-1, :-1] - out["pdemean"]) / out["pdestd"]   # :128
        du = (data_du - pdeu)[:, :, :-1, :-1]                                     # :129
    else:
        du = data_du[:, :, :-1, :-1]                                              # :137
    out["outmean"] = du.mean(dim=(0, 2, 3), keepdim=True)                         # :144
    out["outstd"] = du.std(dim=(0, 2, 3), keepdim=True)                           # :145
    out["du_normd"] = (du - out["outmean"]) / out["outstd"]                       # :146
    out["mumean"] = mus.mean()                                                    # :149
    out["mustd"] = mus.std()                                                      # :150
    out["mu_normd"] = (mus - mus.mean()) / mus.std()                              # :151
    return out


# --------------------------------------------------------------------------
# Fine-grid explicit solver (operators.py:403-492).
# --------------------------------------------------------------------------
def fine_burgers_step(u, v, mu, dt: float = 1e-4, dx: float = 0.0125, dy: float = 0.0125,
                      mode: str = "as_written", adv_accuracy: int = 3, diff_accuracy: int = 6,
                      period_endpoint: bool = True):
    """One Euler step of the dataset generator on [B,1,F,F] fields.

    ``as_written`` follows operators.py:466-475 literally: the v update
    receives the *updated* u and the halos of the *old u* (SURVEY quirk Q3).
    ``intended`` is the commented formula on the same lines: v's own halos.
    ``period_endpoint=False`` wraps a grid without duplicated endpoint.
    """
    o = Ops(adv_accuracy=adv_accuracy, diff_accuracy=diff_accuracy)
    h = max(o.adv_halo, o.diff_halo)
    assert o.adv_halo == o.diff_halo == h, "the generator pads one halo for both stencils"
    px, py = (fine_wrap_x, fine_wrap_y) if period_endpoint else (wrap_x, wrap_y)
    ux, uy, vx, vy = px(u, h), py(u, h), px(v, h), py(v, h)
    dx2, dy2 = dx ** 2, dy ** 2
    u = u + dt * rhs("burgers2Dpu", u, v, mu, o, ux, uy, dx, dy, dx2, dy2)
    if mode == "as_written":
        v = v + dt * rhs("burgers2Dpv", u, v, mu, o, ux, uy, dx, dy, dx2, dy2)
    elif mode == "intended":
        v = v + dt * rhs("burgers2Dpv", u, v, mu, o, vx, vy, dx, dy, dx2, dy2)
    else:
        raise ValueError(mode)
    return u, v


def fine_burgers_run(u, v, mu, nsteps: int, every: int = 200, **kw):
    """operators.py:466-484: snapshot when ``i % every == 0`` (so the first
    snapshot is the state after ONE step, as in the reference loop)."""
    su, sv = [], []
    for i in range(nsteps):
        u, v = fine_burgers_step(u, v, mu, **kw)
        if i % every == 0:
            su.append(u.clone())
            sv.append(v.clone())
    return u, v, su, sv


# --------------------------------------------------------------------------
# Small explicit solvers: Kuramoto-Sivashinsky with RK4 (operators.py:750-918) and
# the 1-D heat equation with Dirichlet / Neumann ghost cells (Bv/data.py).
# --------------------------------------------------------------------------
def rk4_step(f: Callable, u: torch.Tensor, dt: float) -> torch.Tensor:
    """operators.py:805-815 / :882-890."""
    k1 = f(u)
    k2 = f(u + dt * k1 / 2)
    k3 = f(u + dt * k2 / 2)
    k4 = f(u + dt * k3)
    return u + dt * (k1 + 2 * k2 + 2 * k3 + k4) / 6


def ks2d_rhs(u, dx: float, dy: float, c2: float = 0.1, mixed_scale: float = 1.0,
             adv_acc: int = 3, d2_acc: int = 6, d4_acc: int = 4, period_endpoint: bool = True):
    """``f`` of operators.py:806-809 on [B,1,H,W] (pads :761-765; operators :797-802).  As written there the
    mixed term is not divided by dx2*dy2 (``mixed_scale = 1``) and the second derivatives carry 0.1."""
    h = adv_acc
    assert h == d2_acc // 2 == d4_acc // 2 + 1, "the generator pads one halo for every stencil"
    px, py = (fine_wrap_x, fine_wrap_y) if period_endpoint else (wrap_x, wrap_y)
    dx2, dy2 = dx ** 2, dy ** 2
    dx4, dy4 = dx2 ** 2, dy2 ** 2
    L2x = lambda w: stencil2d(px(w, h), "Central2", d2_acc, 2)
    L2y = lambda w: stencil2d(py(w, h), "Central2", d2_acc, 3)
    m = mixed_scale
    mixed1 = L2y(L2x(u)) if m == 1.0 else m * L2y(L2x(u))
    mixed2 = L2x(L2y(u)) if m == 1.0 else m * L2x(L2y(u))
    return (- c2 * L2x(u) / dx2 - c2 * L2y(u) / dy2 - stencil2d(px(u, h), "Central4", d4_acc, 2) / dx4
            - stencil2d(py(u, h), "Central4", d4_acc, 3) / dy4 - mixed1 - mixed2
            - 0.5 * upwind2d(px(u * u, h), adv_acc, 2) / dx - 0.5 * upwind2d(py(u * u, h), adv_acc, 3) / dy)


def ks2d_rk4_step(u, dt: float = 1e-3, dx: float = 0.39269908169872414, dy: Optional[float] = None, **kw):
    dy = dx if dy is None else dy
    return rk4_step(lambda w: ks2d_rhs(w, dx, dy, **kw), u, dt)


def ks1d_rhs(u, dx: float, source=None, adv_acc: int = 3, d2_acc: int = 6, d4_acc: int = 4,
             period_endpoint: bool = True):
    """``f`` of operators.py:883-884 on [B,1,n]: rhs.ks1d on the pad of :852-853, plus the source term."""
    h = adv_acc
    assert h == d2_acc // 2 == d4_acc // 2 + 1
    if period_endpoint:
        pad = lambda w: torch.cat((w[:, :, -h - 1:-1], w, w[:, :, 1:h + 1]), dim=2)
    else:
        pad = lambda w: torch.cat((w[:, :, -h:], w, w[:, :, :h]), dim=2)
    dx2 = dx ** 2
    dx4 = dx2 ** 2
    r = ks1d(pad(u), None, pad(u), dx, dx2, dx4, adv_acc=adv_acc, d2_acc=d2_acc, d4_acc=d4_acc)
    return r if source is None else r + source


def ks1d_rk4_step(u, dt: float = 1e-4, dx: float = 0.39269908169872414, source=None, **kw):
    return rk4_step(lambda w: ks1d_rhs(w, dx, source, **kw), u, dt)


def bv_pad(u, cls, bv, dx):
    """padder.__call__ (Bv/data.py:25-28)."""
    pad2 = lambda w, b: w + b * dx
    bcl = (cls[..., :1] == 0) * bv[..., :1] + (cls[..., :1] == 1) * pad2(u[..., :1], bv[..., :1])
    bcr = (cls[..., 1:] == 0) * bv[..., 1:] + (cls[..., 1:] == 1) * pad2(u[..., -1:], bv[..., 1:])
    return torch.cat([bcl, u, bcr], dim=-1)


def bv_heat_run(u, cls, bv, nsteps: int, saveevery: int, dt, dx, dx2):
    """The generator loop of Bv/data.py:92-98 -> (final u, stacked padded snapshots [B,T,1,n+2])."""
    results = []
    for t in range(nsteps):
        up = bv_pad(u, cls, bv, dx)
        u = u + stencil1d(up, "Central", 2) / dx2 * dt
        if t % saveevery == 0:
            results.append(up)
    return u, torch.stack(results, dim=1)


def bv_coarse_increment(u, cls, bv, scale: float, dt, cdx, cdx2):
    """genforward(...).forward (Bv/data.py:39-51) with the callables of :113-117."""
    uc = F.interpolate(u, scale_factor=scale, mode="linear", align_corners=True)
    cdu = stencil1d(uc, "Central", 2) / cdx2 * dt
    uc1 = bv_pad(uc[..., 1:-1] + cdu, cls, bv, cdx)
    return F.interpolate(uc1, scale_factor=1 / scale, mode="linear", align_corners=True) - u


def fourier_initial_condition(coef: torch.Tensor, n: int, modes: int = 4) -> torch.Tensor:
    """operators.py:423-445 with the random coefficients given: ``coef`` is
    [B, (2*modes+1)**2, 2] (sin, cos) in the (k, l) loop order of the reference;
    result is min-max normalised to [0.1, 1.1]."""
    from math import pi
    x = torch.linspace(0, 2 * pi, n, dtype=coef.dtype)
    x, y = torch.meshgrid(x, x, indexing="ij")
    x = x[None, None].repeat(coef.shape[0], 1, 1, 1)
    y = y[None, None].repeat(coef.shape[0], 1, 1, 1)
    f = torch.zeros_like(x)
    m = 0
    for k in range(-modes, modes + 1):
        for l in range(-modes, modes + 1):
            f += coef[:, m, 0].reshape(-1, 1, 1, 1) * torch.sin(k * x + l * y) \
                + coef[:, m, 1].reshape(-1, 1, 1, 1) * torch.cos(k * x + l * y)
            m += 1
    lo = f.amin(dim=(1, 2, 3), keepdim=True)
    hi = f.amax(dim=(1, 2, 3), keepdim=True)
    return (f - lo) / (hi - lo) + 0.1


# --------------------------------------------------------------------------
# Closed forms (numpy, float64) used to cross-check the interpolation and
# to document the index arithmetic the CUDA kernels implement.
# --------------------------------------------------------------------------
def bilinear_matrix(n_in: int, n_out: int):
    """Dense [n_out, n_in] matrix of 1-D linear resampling, align_corners=True."""
    import numpy as np
    m = np.zeros((n_out, n_in))
    s = (n_in - 1) / (n_out - 1) if n_out > 1 else 0.0
    for o_ in range(n_out):
        src = s * o_
        i0 = min(int(np.floor(src)), n_in - 1)
        i1 = min(i0 + 1, n_in - 1)
        t = src - i0
        m[o_, i0] += 1 - t
        m[o_, i1] += t
    return m


def bicubic_matrix(n_in: int, n_out: int, a: float = -0.75):
    """Dense [n_out, n_in] matrix of 1-D Keys cubic resampling with clamped
    borders, align_corners=True (what upsample_bicubic2d does per axis)."""
    import numpy as np
    m = np.zeros((n_out, n_in))
    s = (n_in - 1) / (n_out - 1) if n_out > 1 else 0.0

    def w_near(x):
        return ((a + 2) * x - (a + 3)) * x * x + 1

    def w_far(x):
        return ((a * x - 5 * a) * x + 8 * a) * x - 4 * a

    for o_ in range(n_out):
        src = s * o_
        i0 = int(np.floor(src))
        t = src - i0
        ws = (w_far(t + 1), w_near(t), w_near(1 - t), w_far(2 - t))
        for k, w in zip((-1, 0, 1, 2), ws):
            m[o_, min(max(i0 + k, 0), n_in - 1)] += w
    return m
