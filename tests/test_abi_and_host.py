"""CPU-side checks: the C-ABI library loads and exports every symbol the header
declares, the Python mirror of the reference interface keeps names, defaults and
assertions, and nothing in the product package touches the oracle."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ppnn_b200.h")


@pytest.fixture(scope="module")
def lib():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as G
    if not os.path.exists(G.LIB):
        G.build()
    from ppnn_b200 import _lib
    return _lib


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ppnn_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 14
    cdll = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(cdll, n), f"{n} declared in include/ppnn_b200.h but not exported"
    assert sorted(lib.EXPORTS) == names, "ppnn_b200/_lib.py must bind exactly the declared entry points"
    l = lib.load()
    assert l.ppnn_abi_version() == lib.ABI_VERSION
    assert l.ppnn_status_string(0) == b"ok"
    assert b"shared memory" in l.ppnn_status_string(-5)


def test_bad_arguments_are_rejected_without_a_gpu(lib):
    """Argument validation happens before any launch, so it can run here."""
    l = lib.load()
    cfg = lib.CoarseCfg(257, 257, 33, 33, 3, 3, 1, 2, 0.02, 0.1, 0.1, 0.01, 0.01)
    assert l.ppnn_pde_du_fwd(ctypes.byref(cfg), None, None, None, 1, None) == -1          # NULL
    assert l.ppnn_stencil_fwd(1, 1, 1, 8, 8, 2, 99, 2, None) == -3                         # scheme
    assert l.ppnn_stencil_fwd(1, 1, 1, 8, 8, 2, lib.CENTRAL2, 3, None) == -3               # accuracy
    assert l.ppnn_stencil_fwd(1, 1, 1, 2, 8, 2, lib.CENTRAL2, 6, None) == -2               # too small
    assert l.ppnn_stencil_fwd(1, 1, 1, 8, 8, 5, lib.CENTRAL2, 2, None) == -6               # axis
    assert l.ppnn_resample_fwd(1, 1, 1, 8, 8, 4, 4, 7, 1, None) == -6                      # mode
    bad = lib.CoarseCfg(257, 257, 33, 33, 3, 64, 1, 2, 0.02, 0.1, 0.1, 0.01, 0.01)
    assert l.ppnn_pde_du_fwd(ctypes.byref(bad), 1, 1, 1, 1, None) == -4                    # rhs id
    big = lib.CoarseCfg(1025, 1025, 513, 513, 3, 3, 1, 2, 0.02, 0.1, 0.1, 0.01, 0.01)
    assert l.ppnn_pde_du_fwd(ctypes.byref(big), 1, 1, 1, 1, None) == -5                    # smem
    assert l.ppnn_fine_burgers_step(1, 2, 3, 4, 5, 1, 33, 33, 1, 0, 3, 2, 1e-4, .01, .01, None) == -7   # halo
    assert l.ppnn_fine_burgers_step(1, 2, 3, 1, 5, 1, 33, 33, 1, 0, 3, 6, 1e-4, .01, .01, None) == -6   # alias
    assert l.ppnn_fine_burgers_step(1, 2, 3, 4, 5, 1, 33, 33, 1, 9, 3, 6, 1e-4, .01, .01, None) == -6   # mode
    assert l.ppnn_pde_du_fwd(ctypes.byref(cfg), 1, 1, 1, 0, None) == 0                     # empty batch
    assert l.ppnn_pde_du_iter_fwd(ctypes.byref(cfg), 1, 1, 1, 1, -1, None) == -2           # negative sub-step count
    assert l.ppnn_pde_du_iter_fwd(ctypes.byref(cfg), 1, None, 1, 1, 1, None) == -1
    # small on-chip solvers
    assert l.ppnn_ks1d_rk4(None, None, 1, 1, 129, 1, 1, 3, 6, 4, 1e-4, .4, .16, .0256, None) == -1      # NULL
    assert l.ppnn_ks1d_rk4(1, None, 1, 1, 129, 1, 1, 3, 5, 4, 1e-4, .4, .16, .0256, None) == -3         # accuracy
    assert l.ppnn_ks1d_rk4(1, None, 1, 1, 6, 1, 1, 3, 6, 4, 1e-4, .4, .16, .0256, None) == -2           # period <= 2*halo
    assert l.ppnn_ks1d_rk4(1, None, 1, 1, 9000, 1, 1, 3, 6, 4, 1e-4, .4, .16, .0256, None) == -5        # too big
    assert l.ppnn_ks1d_rk4(1, None, 1, 0, 129, 1, 1, 3, 6, 4, 1e-4, .4, .16, .0256, None) == 0          # empty batch
    assert l.ppnn_ks2d_rk4(1, 1, 1, 129, 129, 1, 1, 3, 6, 4, .1, 1., 1e-3, .4, .4, .16, .16, .0256, .0256, None) == -5
    assert l.ppnn_ks2d_rk4(1, 1, 1, 65, 65, 1, 1, 4, 6, 4, .1, 1., 1e-3, .4, .4, .16, .16, .0256, .0256, None) == -3
    assert l.ppnn_ks1d_rhs(1, None, 1, 1, 129, 1, 3, 6, 4, .4, .16, .0256, None) == -6                   # f_out aliases u
    assert l.ppnn_ks2d_rhs(1, 2, 1, 129, 129, 1, 3, 6, 4, .1, 1., .4, .4, .16, .16, .0256, .0256, None) == -5
    assert l.ppnn_heat1d_pad(1, None, 1, 1, 2, 126, .4, None) == -1
    assert l.ppnn_heat1d_update(1, 2, 3, 4, 5, 1, 2, 126, .4, None) == -6                                # out aliases u
    assert l.ppnn_heat1d_bc_steps(1, 1, 1, None, None, 0, 2, 126, 5, .01, .4, .16, None) == -1          # no output at all
    assert l.ppnn_heat1d_bc_steps(1, 1, 1, 1, 1, 100, 2, 126, 5, .01, .4, .16, None) == -2              # snapshot stride < n+2
    assert l.ppnn_heat1d_coarse_increment(1, 1, 1, 1, 2, 128, 32, .5, 1.6, 2.56, None) == -6            # out aliases u
    assert l.ppnn_heat1d_coarse_increment(1, 1, 1, 2, 2, 128, 2, .5, 1.6, 2.56, None) == -2             # coarse mesh < 3


def test_reference_interface_names_defaults_and_asserts(lib):
    import ppnn_b200.operators as ops
    import ppnn_b200.rhs as rhs
    from ppnn_b200.utility.utils import mesh_convertor, model_count
    import inspect
    for cls in ("operator_base1D", "diffusion1D", "advection1d", "dudx_1D",
                "operator_base2D", "d2udx2_2D", "d2udy2_2D", "dudx_2D", "dudy_2D"):
        assert hasattr(ops, cls)
    assert inspect.signature(ops.d2udx2_2D).parameters["accuracy"].default == 2
    assert inspect.signature(ops.dudx_2D).parameters["scheme"].default == "Upwind1"
    assert inspect.signature(ops.dudx_2D).parameters["accuracy"].default == 1
    assert inspect.signature(ops.dudx_1D).parameters["accuracy"].default == 2
    with pytest.raises(AssertionError):
        ops.d2udx2_2D(accuracy=3)
    with pytest.raises(AssertionError):
        ops.d2udx2_2D(scheme="Central1")
    with pytest.raises(IndexError):
        ops.dudx_2D("Upwind1", 4)
    with pytest.raises(KeyError):
        ops.dudy_2D("Central", 2)          # the 2-D names are Central1/2/4
    names16 = ["burgers2DFu", "burgers2DFv", "burgers2Dpu", "burgers2Dpv", "burgers2Ddiffuseu",
               "burgers2Ddiffusev", "rdFu", "rdFv", "rdpu", "rdpv", "diffu", "diffv", "convecu",
               "convecv", "burgers2DWu", "burgers2DWv"]
    for n in names16:
        f = getattr(rhs, n)
        assert len(inspect.signature(f).parameters) == 13, n
    assert len(inspect.signature(rhs.ks1d).parameters) == 9
    m = mesh_convertor(257, 33, dim=2)
    assert callable(m.down) and callable(m.up) and m.align_corners is True
    assert model_count(torch.nn.Linear(3, 2)) == 8
    with pytest.raises(RuntimeError):
        ops.d2udx2_2D()(torch.zeros(1, 1, 8, 8))          # CPU tensor: no fallback


def test_rhs_composed_path_matches_oracle_on_cpu_operators():
    """The term-by-term composition (used when the operators are not ppnn_b200's)
    is plain tensor algebra: check it against the oracle with oracle operators."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ppnn_oracle as O
    import ppnn_b200.rhs as rhs
    g = torch.Generator().manual_seed(3)
    u1, v1 = torch.randn(2, 1, 9, 9, generator=g), torch.randn(2, 1, 9, 9, generator=g)
    mu = torch.tensor([0.03, 0.05]).reshape(-1, 1, 1, 1)
    o = O.Ops(1, 2)
    for name in O.RHS_NAMES_2D:
        fld = v1 if name.endswith("v") else u1
        ax, ay = O.wrap_x(fld), O.wrap_y(fld)
        ours = getattr(rhs, name)(u1, v1, mu, o.dudx, o.dudy, o.d2udx2, o.d2udy2, ax, ay, 0.1, 0.2, 0.01, 0.04)
        ref = O.rhs(name, u1, v1, mu, o, ax, ay, 0.1, 0.2, 0.01, 0.04)
        assert torch.allclose(ours, ref, rtol=1e-6, atol=1e-6), name


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ppnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "ppnn_oracle" not in text and "oracle/" not in text and "ref_closures" not in text, \
                    f"{f} refers to the oracle"


def test_dropin_directory_resolves_reference_import_spellings():
    """train2D.py:12-14 does `import rhs`, `from operators import ...`,
    `from utility.utils import mesh_convertor`; Bv/ uses `from src.operators import ...`."""
    code = ("import operators, rhs; from utility.utils import mesh_convertor, model_count; "
            "from src.operators import diffusion1D; import src.rhs; "
            "assert operators.dudx_2D.__module__ == 'ppnn_b200.operators'; "
            "assert rhs.burgers2Dpu.__module__ == 'ppnn_b200.rhs'; print('ok')")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "ppnn_b200", "dropin") + os.pathsep + ROOT)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr
