"""The reference's own training script, UNCHANGED, on top of the drop-in modules.

``oracle/_ref/src/train2D.py`` is a byte-identical copy of the reference's script (staged by
``oracle/stage_ref.py`` in the build container; SHA-256 in ``oracle/_ref/MANIFEST.json``).  It is
executed as ``__main__`` with ``ppnn_b200/dropin`` first on the module path, on the reference's
own ``cases/burgers_pPDE.yaml`` / ``RD_pPDE.yaml`` with only the machine-specific keys (device,
file paths) and the problem SIZE keys (num_para, repeat, timesteps, epochs, batchsize) replaced.
It builds the training set through ``pde_du`` (train2D.py:121), trains one epoch and runs the
autoregressive rollout test of train2D.py:230-245 -- every operator / rhs / mesh_convertor call
on that path must land in the kernels of ``libppnn_b200.so``."""
import hashlib
import json
import os
import subprocess
import sys

import pytest
import torch
import yaml

from conftest import ROOT

pytestmark = pytest.mark.gpu
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def _fields(P, T, F, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.linspace(0, 2 * torch.pi, F)
    k = torch.arange(-4, 5, dtype=torch.float32)
    sx, cx = torch.sin(k[:, None] * x[None]), torch.cos(k[:, None] * x[None])
    a, c = torch.randn(P * 2, 9, 9, generator=g), torch.randn(P * 2, 9, 9, generator=g)
    f = (torch.einsum("nkl,ki,lj->nij", a, sx, cx) + torch.einsum("nkl,ki,lj->nij", a, cx, sx)
         + torch.einsum("nkl,ki,lj->nij", c, cx, cx) - torch.einsum("nkl,ki,lj->nij", c, sx, sx))
    lo, hi = f.amin(dim=(1, 2), keepdim=True), f.amax(dim=(1, 2), keepdim=True)
    u0 = ((f - lo) / (hi - lo) + 0.1).reshape(P, 2, F, F)
    # a smooth, slowly decaying stand-in trajectory [P, T+1, 2, F, F] in the generator's wire format
    return torch.stack([u0 * (1.0 - 0.01 * t) for t in range(T + 1)], dim=1).contiguous()


@pytest.mark.parametrize("case", ["burgers_pPDE", "RD_pPDE"])
def test_train2d_runs_unchanged_on_the_dropin(case, tmp_path):
    script = os.path.join(REFDIR, "src", "train2D.py")
    if not os.path.exists(script):
        pytest.skip("oracle/_ref not staged (python oracle/stage_ref.py in the build container)")
    manifest = json.load(open(os.path.join(REFDIR, "MANIFEST.json")))["sha256"]
    assert hashlib.sha256(open(script, "rb").read()).hexdigest() == manifest["src/train2D.py"]

    cfg = yaml.load(open(os.path.join(REFDIR, "cases", case + ".yaml")), Loader=yaml.FullLoader)
    P, T = 2, 3
    data = tmp_path / "data.pt"
    torch.save({"u": _fields(P, T, cfg["finemeshsize"], 5)}, data)
    cfg.update(device="cuda:0", num_para=P, repeat=1, timesteps=T, epochs=0, batchsize=3, datatimestart=0,
               datafile=str(data), tensorboarddir=str(tmp_path / "tb"), modelsavepath=str(tmp_path / "model.pth"))
    os.makedirs(cfg["tensorboarddir"])
    ypath = tmp_path / (case + ".yaml")
    yaml.safe_dump(cfg, open(ypath, "w"))
    report = tmp_path / "report.json"
    env = dict(os.environ)
    env.pop("PPNN_RUNNER_NO_DROPIN", None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_reference_script.py"), str(report), script, str(ypath)],
                       capture_output=True, text=True, timeout=900, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    rep = json.load(open(report))

    # the hot-path modules resolved to the drop-in, the network module to the reference's own models.py
    for name in ("operators", "rhs", "utility.utils"):
        assert os.path.join("ppnn_b200", "dropin") in rep["modules"][name], rep["modules"]
    assert rep["modules"]["models"].startswith(REFDIR)
    # ... and their arithmetic ran in our kernels: no ATen interpolation, no single-channel ATen stencil convs
    kernels = " ".join(rep["kernels"])
    assert "resample" in kernels, kernels
    assert "rhs2d" in kernels, kernels
    for op in ("aten::upsample_bicubic2d", "aten::upsample_bilinear2d"):
        assert rep["aten_ops"].get(op, 0) == 0, (op, rep["aten_ops"].get(op))
    # the script's own outputs: normalisation statistics, checkpoint, finite rollout error
    for f in ("inmean.pt", "instd.pt", "outmean.pt", "outstd.pt", "pdemean.pt", "pdestd.pt", "parsmean.pt", "parsstd.pt"):
        assert os.path.exists(os.path.join(cfg["tensorboarddir"], f)), f
    assert os.path.exists(cfg["modelsavepath"])
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        json.dump(rep, open(os.path.join(out, f"train2d_unchanged_{case}.json"), "w"), indent=1)
