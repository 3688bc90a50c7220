"""The staged, UNMODIFIED reference modules (``oracle/_ref/src``) as a CPU baseline / checker.

TEST INFRASTRUCTURE ONLY (same rule as ``ppnn_oracle.py``: importable from ``tests/``,
``__graft_entry__.smoke()`` and the CPU legs of ``bench.py`` -- never from ``ppnn_b200/``).

``operators.py``, ``rhs.py`` and ``utility/utils.py`` are importable modules of the
reference and are loaded here exactly as staged (byte-identical copies, see
``stage_ref.py`` and ``_ref/MANIFEST.json``).  What is NOT importable are the closures
that live inside the reference's ``__main__`` blocks; ``RefCoarseStep`` binds those
(src/train2D.py:38-47, 49-79, 235-238) around the loaded modules -- all arithmetic is
executed by the reference's own classes and functions.
"""
from __future__ import annotations

import importlib.util
import os
import sys
from types import SimpleNamespace
from typing import Optional

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.path.join(HERE, "_ref", "src")
_cache: Optional[SimpleNamespace] = None


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_SRC, f)) for f in ("operators.py", "rhs.py", "utility/utils.py"))


def _load(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def modules() -> SimpleNamespace:
    """The reference's ``operators`` / ``rhs`` / ``utility.utils`` under private module names
    (so they can never shadow, or be shadowed by, the ppnn_b200 drop-in modules)."""
    global _cache
    if _cache is None:
        if not available():
            raise ImportError("oracle/_ref is not staged: run `python oracle/stage_ref.py` in the build container")
        _cache = SimpleNamespace(operators=_load("_ppnn_ref_operators", os.path.join(REF_SRC, "operators.py")),
                                 rhs=_load("_ppnn_ref_rhs", os.path.join(REF_SRC, "rhs.py")),
                                 utils=_load("_ppnn_ref_utils", os.path.join(REF_SRC, "utility", "utils.py")))
    return _cache


class RefCoarseStep:
    """src/train2D.py:49-79 with the reference's own operator / rhs / mesh_convertor objects."""

    def __init__(self, fine: int, cmesh: int, rhsu: str, rhsv: str, dt: float, length: float, device="cpu"):
        m = modules()
        self.rhsu, self.rhsv = getattr(m.rhs, rhsu), getattr(m.rhs, rhsv)          # train2D.py:50-51
        self.d2udx2 = m.operators.d2udx2_2D(accuracy=2, device=device)             # :53-56
        self.d2udy2 = m.operators.d2udy2_2D(accuracy=2, device=device)
        self.dudx = m.operators.dudx_2D(accuracy=1, device=device)
        self.dudy = m.operators.dudy_2D(accuracy=1, device=device)
        self.dx = length / (cmesh - 1)                                              # :58
        self.mcvter = m.utils.mesh_convertor(fine, cmesh, dim=2)                    # :59
        self.dx2 = self.dx ** 2
        self.dy, self.dy2 = self.dx, self.dx2
        self.dt = dt

    @staticmethod
    def padbcx(uinner):                                                             # train2D.py:38-39
        return torch.cat((uinner[:, :, -1:], uinner, uinner[:, :, :1]), dim=2)

    @staticmethod
    def padbcy(uinner):                                                             # :41-42
        return torch.cat((uinner[:, :, :, -1:], uinner, uinner[:, :, :, :1]), dim=3)

    @staticmethod
    def padBC_rd(u):                                                                # :45-47
        tmp = torch.cat((u, u[:, :, :, :1]), dim=3)
        return torch.cat((tmp, tmp[:, :, :1]), dim=2)

    def pde_du(self, u, mu):                                                        # :66-79
        s = self
        u1 = s.mcvter.down(u[:, :1])[:, :, :-1, :-1]
        v1 = s.mcvter.down(u[:, 1:])[:, :, :-1, :-1]
        ux, uy, vx, vy = s.padbcx(u1), s.padbcy(u1), s.padbcx(v1), s.padbcy(v1)
        return torch.cat(
            (s.mcvter.up(s.padBC_rd(s.dt * s.rhsu(u1, v1, mu, s.dudx, s.dudy, s.d2udx2, s.d2udy2, ux, uy,
                                                  s.dx, s.dy, s.dx2, s.dy2))),
             s.mcvter.up(s.padBC_rd(s.dt * s.rhsv(u1, v1, mu, s.dudx, s.dudy, s.d2udx2, s.d2udy2, vx, vy,
                                                  s.dx, s.dy, s.dx2, s.dy2)))), dim=1)

    def update(self, u, mu, nn_out, outmean, outstd):                               # :235-238, model(...) := nn_out
        pdeuu = self.pde_du(u, mu)
        return self.padBC_rd(nn_out * outstd + outmean) + u + pdeuu
