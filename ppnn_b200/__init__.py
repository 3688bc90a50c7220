"""ppnn_b200 -- B200-native (sm_100a) implementation of the PDE-preserving hot path of
PPNN (jx-wang-s-group/ppnn): the periodic finite-difference operators, the Burgers /
reaction-diffusion right-hand sides, the fused coarse-grid ``pde_du`` step with its
adjoint, and the batched fine-grid explicit solver.

Tier A (drop-in for the reference's ``src/`` modules; put ``ppnn_b200/dropin`` first
on PYTHONPATH, or import from here):
    ppnn_b200.operators      d2udx2_2D, d2udy2_2D, dudx_2D, dudy_2D, diffusion1D, dudx_1D
    ppnn_b200.rhs            burgers2D{F,p,diffuse,W}{u,v}, rd{F,p}{u,v}, diff*, convec*, ks1d
    ppnn_b200.utility.utils  mesh_convertor, model_count
Tier B (fully fused):
    ppnn_b200.fused          CoarseStep (pde_du, step_inputs, step_update, rollouts), FineBurgers, FineRhsSolver
    ppnn_b200.solvers        KS1D, KS2D (Kuramoto-Sivashinsky RK4), HeatBC1D, HeatCoarseIncrement (Bv/): many steps per launch
    ppnn_b200.bv             padder, denormBV, genforward (mirror of Bv/data.py)
    ppnn_b200.datagen        sharded dataset generation (python -m ppnn_b200.datagen)
    ppnn_b200.dist           sample sharding / global statistics helpers

All device work goes through the C ABI of ``libppnn_b200.so`` (include/ppnn_b200.h).
There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .fused import CoarseStep, FineBurgers, FineRhsSolver, HostStepPipeline, RolloutPlan  # noqa: F401

__all__ = ["CoarseStep", "FineBurgers", "FineRhsSolver", "operators", "rhs", "utility"]
from .dataset import PdeDataset  # noqa: F401
from .solvers import KS1D, KS2D, HeatBC1D, HeatCoarseIncrement  # noqa: F401
