#!/usr/bin/env python
"""Place the UNMODIFIED reference where it can travel to the GPU box.

    python tools/install_ref.py            # /root/reference/codes -> baseline/_ref/codes

`baseline/_ref/` is git-ignored (reference sources never enter the history) but NOT gpurun-ignored, so the
copy ships with the repo snapshot: the `-m gpu` parity tests (tests/test_reference_gpu.py) and bench.py's
`gpu_reference` / `--impl reference` legs then run the reference's own modules -- real
`flash_attn_varlen_func`, CUDA autocast -- on the B200 next to the engine.  The whole `codes/` tree is
copied (608 KB): `modeling/__init__.py` imports all of its sub-packages, partial copies break.  Nothing
under `baseline/_ref` is imported by the product (`unimedvl_b200/`); it is test / measurement infrastructure
only.  A manifest with the sha256 of every copied file is written next to the copy so a test can show the
tree is byte-identical to what was installed.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("UMV_REFERENCE_SRC", "/root/reference/codes")
DST = os.path.join(ROOT, "baseline", "_ref", "codes")


def tree_manifest(base: str) -> dict:
    out = {}
    for d, _, files in sorted(os.walk(base)):
        if "__pycache__" in d:
            continue
        for f in sorted(files):
            p = os.path.join(d, f)
            out[os.path.relpath(p, base)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    return out


def install(force: bool = False) -> str:
    if not os.path.isdir(SRC):
        if os.path.isdir(DST):
            return DST                       # GPU box: the prebuilt copy is all there is
        raise SystemExit(f"reference sources not found at {SRC}")
    if os.path.isdir(DST) and not force and tree_manifest(DST) == tree_manifest(SRC):
        return DST
    shutil.rmtree(DST, ignore_errors=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(os.path.dirname(DST), "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": tree_manifest(DST)}, f, indent=1)
    return DST


if __name__ == "__main__":
    p = install(force="--force" in sys.argv)
    print("reference at", p, f"({len(tree_manifest(p))} files)")
