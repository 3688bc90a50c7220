#!/usr/bin/env python
"""Stage the UNMODIFIED reference modules of the hot path into ``oracle/_ref/``.

TEST INFRASTRUCTURE.  ``oracle/_ref/`` is git-ignored (the reference's sources never
enter this repository's history) but it is NOT gpurun-ignored, so the staged files
travel to the GPU box, where ``/root/reference`` does not exist.  They are used only
as the checker / baseline:

* ``bench.py``'s ``cpu_baseline`` leg and ``--impl reference`` arm execute the staged
  ``operators.py`` / ``rhs.py`` / ``utility/utils.py`` (``kind: "reference"``);
* ``tests/test_gpu_reference_scripts.py`` runs the staged ``train2D.py`` UNCHANGED with
  ``ppnn_b200/dropin`` first on ``PYTHONPATH`` (the staged ``models.py`` supplies the
  trainable network, which north_star leaves in PyTorch).

Run by ``__graft_entry__.build()``; a no-op when ``/root/reference`` is absent (the GPU
box uses what was staged in the build container).  Files are byte-for-byte copies; the
manifest records their SHA-256 so a test can prove nothing was edited.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
REF = os.environ.get("PPNN_REFERENCE", "/root/reference")
FILES = [
    "src/__init__.py", "src/operators.py", "src/rhs.py", "src/models.py", "src/train2D.py",
    "src/utility/utils.py",
    "Bv/data.py",
    "cases/burgers_pPDE.yaml", "cases/RD_pPDE.yaml", "cases/burgers_fPDE.yaml",
]


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    if not os.path.isdir(REF):
        have = os.path.exists(os.path.join(DEST, "MANIFEST.json"))
        print(f"[stage_ref] {REF} not present; " + ("using the staged copy" if have else "nothing staged"))
        return 0
    manifest = {}
    for rel in FILES:
        src = os.path.join(REF, rel)
        if not os.path.exists(src):
            print(f"[stage_ref] missing in the reference: {rel}", file=sys.stderr)
            continue
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and sha(dst) == sha(src)):
            shutil.copyfile(src, dst)
        manifest[rel] = sha(dst)
    json.dump({"source": REF, "sha256": manifest}, open(os.path.join(DEST, "MANIFEST.json"), "w"), indent=1)
    print(f"[stage_ref] {len(manifest)} reference files staged in {DEST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
