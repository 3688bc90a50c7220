"""Pin the CPU oracle against outputs of the reference itself
(tests/golden/*.npz, produced by tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

import ppnn_oracle as O
from conftest import assert_parity, load_golden

TOL = 2e-6     # same ATen ops in the same order; only thread-count / blocking may differ


def test_operator_tables_and_ops():
    g = load_golden("operators.npz")
    u, u1d = g["u2d"], g["u1d"]
    n = 0
    for key, ref in g.items():
        if "/" not in key:
            continue
        cls, scheme, acc = key.split("/")
        acc = int(acc)
        if cls.endswith("_2D"):
            axis = 2 if cls in ("dudx_2D", "d2udx2_2D") else 3
            ours = O.stencil2d(u, scheme, acc, axis)
        else:
            ours = O.stencil1d(u1d, scheme, acc)
        assert_parity(ours, ref, TOL, key)
        n += 1
    assert n == 54


def test_rhs_all_names():
    g = load_golden("rhs.npz")
    u1, v1, mu = g["u1"], g["v1"], g["mu"]
    names = [str(s) for s in g["names"]]
    assert sorted(names) == sorted(O.RHS_NAMES_2D)
    for aa, da in ((1, 2), (3, 6)):
        o = O.Ops(adv_accuracy=aa, diff_accuracy=da)
        for name in names:
            fld = v1 if name.endswith("v") else u1
            ours = O.rhs(name, u1, v1, mu, o, O.wrap_x(fld, aa), O.wrap_y(fld, aa), 0.1, 0.08, 0.1 ** 2, 0.08 ** 2)
            assert_parity(ours, g[f"{name}/{aa}{da}"], TOL, f"{name}/{aa}{da}")
    w = g["ks_w"]
    ours = O.ks1d(torch.cat((w[:, :, -2:], w, w[:, :, :2]), 2), None,
                  torch.cat((w[:, :, -1:], w, w[:, :, :1]), 2), 0.3, 0.09, 0.0081)
    assert_parity(ours, g["ks1d"], TOL, "ks1d")


def test_mesh_converter():
    g = load_golden("mesh.npz")
    for c in (33, 49):
        d = O.mesh_down(g["f257"], c)
        assert_parity(d, g[f"down/257/{c}"], TOL, f"down 257->{c}")
        assert_parity(O.mesh_up(d, 257), g[f"up/{c}/257"], TOL, f"up {c}->257")
    for c in (17, 13, 9):
        d = O.mesh_down(g["r65"], c)
        assert_parity(d, g[f"down/65/{c}"], TOL, f"down 65->{c}")
        assert_parity(O.mesh_up(d, 65), g[f"up/{c}/65"], TOL, f"up {c}->65")
    assert_parity(O.mesh_down(g["rect"], (9, 13)), g["rect_down"], TOL, "rect down")
    assert_parity(O.mesh_up(g["rect"], (40, 57)), g["rect_up"], TOL, "rect up")
    for ac in (1, 0):
        d = O.mesh_down(g[f"ac{ac}_in"], 7, align_corners=bool(ac))
        assert_parity(d, g[f"ac{ac}_down"], TOL, f"ac{ac} down")
        assert_parity(O.mesh_up(d, 31, align_corners=bool(ac)), g[f"ac{ac}_up"], TOL, f"ac{ac} up")
    d = O.mesh_down(g["line"], 11)
    assert_parity(d, g["line_down"], TOL, "1-D down")
    assert_parity(O.mesh_up(d, 101), g["line_up"], TOL, "1-D up")


def test_closed_form_interpolation_matrices():
    """bilinear / Keys-cubic (A=-0.75, clamped) closed forms."""
    for n_in, n_out in ((257, 33), (257, 49), (65, 13)):
        x = torch.randn(1, 1, n_in, n_in, dtype=torch.float64)
        m = torch.from_numpy(O.bilinear_matrix(n_in, n_out))
        assert (m @ x[0, 0] @ m.T - O.mesh_down(x, n_out)[0, 0]).abs().max() < 1e-13
    for n_in, n_out in ((33, 257), (49, 257), (13, 65)):
        x = torch.randn(1, 1, n_in, n_in, dtype=torch.float64)
        m = torch.from_numpy(O.bicubic_matrix(n_in, n_out))
        assert (m @ x[0, 0] @ m.T - O.mesh_up(x, n_out)[0, 0]).abs().max() < 1e-13


CASES = {
    "burgers_p": dict(cmesh=33, rhsu="burgers2Dpu", rhsv="burgers2Dpv", dt=0.02, length=3.2, mu=0.07),
    "rd_p": dict(cmesh=49, rhsu="rdpu", rhsv="rdpv", dt=0.002, length=6.4, mu=1.3),
    "burgers_f": dict(cmesh=49, rhsu="burgers2DFu", rhsv="burgers2DFv", dt=0.02, length=3.2, mu=0.07),
    "rd_f": dict(cmesh=49, rhsu="rdFu", rhsv="rdFv", dt=0.002, length=6.4, mu=1.3),
}
SMALL = {"burgers_p": 9, "rd_p": 13, "burgers_f": 13, "rd_f": 13}


@pytest.mark.parametrize("case", sorted(CASES))
def test_pde_du_full_size(case):
    g = load_golden("pde_du_257.npz")
    c = CASES[case]
    step = O.CoarseStep(257, c["cmesh"], c["rhsu"], c["rhsv"], c["dt"], c["length"])
    mu = torch.tensor([c["mu"]]).reshape(1, 1, 1, 1)
    assert_parity(step.pde_du(g["u257"], mu), g[f"{case}/257"], TOL, case)
    assert_parity(step.pde_du_iter(g["u257"], mu, 1), g[f"{case}/257_iter1"], TOL, case + " iter variant (test2D2.py:70-91)")


@pytest.mark.parametrize("case", sorted(CASES))
def test_pde_du_small_with_grads_update_rollout(case):
    g = load_golden("pde_du_65.npz")
    c = CASES[case]
    step = O.CoarseStep(65, SMALL[case], c["rhsu"], c["rhsv"], c["dt"], float(g[f"{case}/length"]))
    mu = g[f"{case}/mu"]
    u = g["u65"].clone().requires_grad_(True)
    m = mu.clone().requires_grad_(True)
    y = step.pde_du(u, m)
    assert_parity(y.detach(), g[f"{case}/pdeu"], TOL, case + " pdeu")
    for inner in (1, 2):
        assert_parity(step.pde_du_iter(g["u65"], mu, inner), g[f"{case}/pdeu_iter{inner}"], TOL, f"{case} iter{inner}")
    gu, gm = torch.autograd.grad(y, (u, m), g["cot65"], allow_unused=True)
    assert_parity(gu, g[f"{case}/grad_u"], TOL, case + " grad_u")
    if gm is not None and g[f"{case}/grad_mu"].abs().max() > 0:
        assert_parity(gm, g[f"{case}/grad_mu"], 2e-5, case + " grad_mu")
    upd = O.ppnn_update(g["u65"], y.detach(), g["nn65"], g["outmean"], g["outstd"])
    assert_parity(upd, g[f"{case}/update"], TOL, case + " update")
    roll = O.pde_only_rollout(step, g["u65"], mu, 5)[-1]
    assert_parity(roll, g[f"{case}/rollout5"], TOL, case + " rollout5")


@pytest.mark.parametrize("case", ["rd_p", "burgers_p"])
def test_headline_step_full_size(case):
    """train2D.py:235-238 (update with a network-output operand) and :235-237 (network inputs)
    at 257^2, B = 3 -- the shapes of the kernels bench.py times."""
    g = load_golden("step_257.npz")
    c = CASES[case]
    step = O.CoarseStep(257, c["cmesh"], c["rhsu"], c["rhsv"], c["dt"], c["length"])
    u, mu = g["u"], g[f"{case}/mu"]
    pde = step.pde_du(u, mu)
    assert_parity(O.ppnn_update(u, pde, g["nn"], g["outmean"], g["outstd"]), g[f"{case}/update"], TOL, case + " update")
    a, _, cc = O.ppnn_net_inputs(u, mu, pde, g["inmean"], g["instd"], torch.tensor(0.0), torch.tensor(1.0),
                                 g[f"{case}/pdemean"], g[f"{case}/pdestd"])
    assert_parity(a[:, :, ::8, ::8], g[f"{case}/u_norm"], TOL, case + " u_norm")
    assert_parity(cc * g[f"{case}/pdestd"], g[f"{case}/pdeu_norm"] * g[f"{case}/pdestd"], TOL, case + " pdeu_norm")


@pytest.mark.parametrize("mode", ["as_written", "intended"])
def test_fine_solver(mode):
    g = load_golden("fine.npz")
    u, v = g["u0"], g["v0"]
    for i in range(4):
        u, v = O.fine_burgers_step(u, v, g["mu"], mode=mode)
        if i in (0, 3):
            assert_parity(u, g[f"{mode}/u{i + 1}"], TOL, f"{mode} u{i + 1}")
            assert_parity(v, g[f"{mode}/v{i + 1}"], TOL, f"{mode} v{i + 1}")


def test_fine_solver_q3_quirk_is_real():
    """as_written and intended must differ, or the fixture is vacuous."""
    g = load_golden("fine.npz")
    assert (g["as_written/v4"] - g["intended/v4"]).abs().max() > 1e-6


def test_fourier_ic_range():
    coef = torch.randn(2, 81, 2)
    f = O.fourier_initial_condition(coef, 33)
    assert f.shape == (2, 1, 33, 33)
    assert abs(f.min().item() - 0.1) < 1e-6 and abs(f.max().item() - 1.1) < 1e-6


@pytest.mark.parametrize("case", ["burgers_p", "rd_p"])
def test_training_set_precompute(case):
    """oracle.training_set against train2D.py:88-151 run on the reference modules."""
    g = load_golden("dataset_65.npz")
    c = CASES[case]
    step = O.CoarseStep(65, SMALL[case], c["rhsu"], c["rhsv"], c["dt"], float(g[f"{case}/length"]))
    out = O.training_set(step, g[f"{case}/fdata"], g[f"{case}/mus"])
    for k in ("inmean", "instd", "pdemean", "pdestd", "outmean", "outstd", "mumean", "mustd"):
        assert torch.allclose(out[k], g[f"{case}/{k}"], rtol=1e-6, atol=1e-9), k
    for k in ("u0_normd", "pdeu_normd", "du_normd", "mu_normd"):
        assert_parity(out[k], g[f"{case}/{k}"], 1e-6, f"{case}/{k}")


def test_staged_reference_is_unmodified_and_agrees_with_golden_and_port():
    """oracle/_ref (what bench.py's CPU legs execute, kind "reference"): byte-identical to the manifest
    written at staging time, and its pde_du / update reproduce the golden fixture and the port."""
    import hashlib
    import json
    import os
    import ref_closures as R
    if not R.available():
        pytest.skip("oracle/_ref not staged")
    man = json.load(open(os.path.join(R.HERE, "_ref", "MANIFEST.json")))["sha256"]
    for rel, digest in man.items():
        assert hashlib.sha256(open(os.path.join(R.HERE, "_ref", rel), "rb").read()).hexdigest() == digest, rel
    if os.path.isdir("/root/reference"):          # build container: the staged copy IS the reference
        for rel, digest in man.items():
            assert hashlib.sha256(open(os.path.join("/root/reference", rel), "rb").read()).hexdigest() == digest, rel
    g = load_golden("step_257.npz")
    c = CASES["rd_p"]
    ref = R.RefCoarseStep(257, c["cmesh"], c["rhsu"], c["rhsv"], c["dt"], c["length"])
    with torch.no_grad():
        upd = ref.update(g["u"], g["rd_p/mu"], g["nn"], g["outmean"], g["outstd"])
    assert_parity(upd, g["rd_p/update"], TOL, "staged reference update")
    port = O.CoarseStep(257, c["cmesh"], c["rhsu"], c["rhsv"], c["dt"], c["length"])
    assert_parity(O.ppnn_update(g["u"], port.pde_du(g["u"], g["rd_p/mu"]), g["nn"], g["outmean"], g["outstd"]), upd, TOL, "port")


def test_small_solvers_ks_rk4_and_bv_heat():
    """Kuramoto-Sivashinsky RK4 generators (operators.py:750-918) and the Bv/ heat step with ghost cells
    (Bv/data.py): oracle restatement vs the reference's own lines / imported code (solvers.npz)."""
    g = load_golden("solvers.npz")
    u = g["ks2d/u0"]
    assert_parity(O.ks2d_rhs(u, np.pi / 8, np.pi / 8), g["ks2d/f0"], TOL, "ks2d f")
    for i in range(3):
        u = O.ks2d_rk4_step(u, 1e-3, np.pi / 8)
        assert_parity(u, g[f"ks2d/u{i + 1}"], TOL, f"ks2d step {i + 1}")
    u, src = g["ks1d/u0"], g["ks1d/source"].reshape(1, 1, -1)
    assert_parity(O.ks1d_rhs(u, np.pi / 8, src), g["ks1d/f0"], TOL, "ks1d f")
    assert_parity(O.ks1d_rhs(u, np.pi / 8, None), g["ks1d/rhs_ks1d"], TOL, "rhs.ks1d")
    for i in range(3):
        u = O.ks1d_rk4_step(u, 1e-4, np.pi / 8, src)
        assert_parity(u, g[f"ks1d/u{i + 1}"], TOL, f"ks1d step {i + 1}")
    cls, bvr, dx = g["bv/cls"], g["bv/BVr"], g["bv/dx"]
    assert torch.equal(bvr, (cls == 1) * g["bv/BV"] + (cls == 0) * 5 * g["bv/BV"])        # denormBV
    assert_parity(O.bv_pad(g["bv/u0"], cls, bvr, dx), g["bv/pad0"], TOL, "padder")
    uf, res = O.bv_heat_run(g["bv/u0"], cls, bvr, 151, int(g["bv/saveevery"]), float(g["bv/dt"]), dx, dx ** 2)
    assert_parity(res, g["bv/results"], TOL, "bv generator snapshots")
    assert_parity(uf, g["bv/u_final"], TOL, "bv final state")
    scale = float(g["bv/scale"])
    cdx = dx / scale
    for t in range(res.shape[1]):
        inc = O.bv_coarse_increment(g["bv/results"][:, t], cls, bvr, scale, float(g["bv/dt"]) * int(g["bv/saveevery"]), cdx, cdx ** 2)
        assert_parity(inc, g["bv/pdeu"][:, t], TOL, f"genforward t={t}")
