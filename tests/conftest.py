"""Shared pytest plumbing: the ``gpu`` marker, import paths, golden loader."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """npz fixture -> dict of torch tensors (and raw numpy for string arrays)."""
    out = {}
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as z:
        for k in z.files:
            a = z[k]
            out[k] = a if a.dtype.kind in "US" else torch.from_numpy(a.copy())
    return out


def rel_l2(a, b):
    a, b = a.double(), b.double()
    d = (a - b).norm()
    n = b.norm()
    return (d / n).item() if n > 0 else d.item()


def rel_max(a, b):
    a, b = a.double(), b.double()
    d = (a - b).abs().max()
    n = b.abs().max()
    return (d / n).item() if n > 0 else d.item()


def assert_parity(ours, ref, tol=1e-5, what=""):
    """The parity metric of SURVEY.md 8(c): rel-L2 and max-abs/||ref||_inf."""
    assert ours.shape == ref.shape, f"{what}: shape {tuple(ours.shape)} vs {tuple(ref.shape)}"
    r2, rm = rel_l2(ours, ref), rel_max(ours, ref)
    assert r2 <= tol and rm <= tol, f"{what}: rel-L2 {r2:.3e}, rel-max {rm:.3e} > {tol:g}"
