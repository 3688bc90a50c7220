#!/usr/bin/env python
"""Run a staged reference script UNCHANGED on top of the ppnn_b200 drop-in modules.

    python tests/run_reference_script.py <report.json> <script.py> [script args...]

sys.path is ordered the way INTEGRATION.md tells a user to set PYTHONPATH:
``ppnn_b200/dropin`` first (so ``import operators`` / ``import rhs`` /
``from utility.utils import mesh_convertor`` resolve to the CUDA drop-in), then the staged
reference ``src`` (for ``models``: the trainable networks stay in PyTorch), then the
matplotlib stand-in.  The script file itself is executed as ``__main__`` by ``runpy`` --
not one byte of it is edited (the test checks its SHA-256 against ``oracle/_ref/MANIFEST.json``).

Two environment shims, unrelated to the hot path -- the reference targets an older, unpinned
PyTorch (environment.yml:1): (1) it passes ``verbose=True`` to ``ReduceLROnPlateau``
(train2D.py:189), a keyword PyTorch >= 2.7 removed; it is dropped here.  (2) it indexes a CPU
tensor with a CUDA index tensor (train2D.py:88-89, ``fdata[testIDs, 0]``), which old PyTorch
accepted and current PyTorch rejects; on exactly that error the index is moved to the indexed
tensor's device and the indexing retried.

The run is traced with the torch profiler; the report lists which files the hot-path modules
were imported from, the CUDA kernels launched and the ATen ops called.
"""
import json
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    report_path, script, argv = sys.argv[1], sys.argv[2], sys.argv[3:]
    paths = [os.path.join(ROOT, "tests", "stubs"), os.path.join(ROOT, "oracle", "_ref", "src")]
    if not os.environ.get("PPNN_RUNNER_NO_DROPIN"):      # set => the pure reference (CPU dry runs, eager baseline)
        paths += [os.path.join(ROOT, "ppnn_b200", "dropin"), ROOT]
    for p in paths:
        sys.path.insert(0, p)                     # last inserted = first searched: ROOT, dropin, ref src, stubs
    import torch
    from torch.profiler import ProfilerActivity, profile

    sched = torch.optim.lr_scheduler.ReduceLROnPlateau
    orig_init = sched.__init__

    def init_without_verbose(self, *a, verbose=None, **kw):
        orig_init(self, *a, **kw)
    sched.__init__ = init_without_verbose

    orig_getitem = torch.Tensor.__getitem__

    def getitem_with_device_fix(self, idx):
        try:
            return orig_getitem(self, idx)
        except RuntimeError as exc:
            if "indices should be either on cpu" not in str(exc):
                raise
            fix = lambda i: i.to(self.device) if isinstance(i, torch.Tensor) else i   # noqa: E731
            return orig_getitem(self, tuple(fix(i) for i in idx) if isinstance(idx, tuple) else fix(idx))
    torch.Tensor.__getitem__ = getitem_with_device_fix

    sys.argv = [script] + argv
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        runpy.run_path(script, run_name="__main__")
        if torch.cuda.is_available():
            torch.cuda.synchronize()
    kernels, ops = {}, {}
    for ev in prof.key_averages():
        dt = str(getattr(ev, "device_type", ""))
        if "CUDA" in dt:
            kernels[ev.key] = kernels.get(ev.key, 0) + ev.count
        elif ev.key.startswith("aten::"):
            ops[ev.key] = ops.get(ev.key, 0) + ev.count
    mods = {n: getattr(sys.modules.get(n), "__file__", None) for n in ("operators", "rhs", "utility.utils", "models")}
    json.dump({"modules": mods, "kernels": kernels, "aten_ops": ops}, open(report_path, "w"), indent=1)


if __name__ == "__main__":
    main()
