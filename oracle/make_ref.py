#!/usr/bin/env python
"""Stage the UNMODIFIED reference next to the checker so that it travels to the GPU box.

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` exists only in the build container; the GPU box gets a
snapshot of this repository.  This script copies — byte for byte, nothing edited —

* ``/root/reference/libreco``                      (the Python package; ``*.py`` only: the Cython / Rust
  extensions are optional at import time, ``algorithms/als.py:137`` / ``bpr.py:311`` only warn), and
* ``/root/reference/examples/sample_data/sample_movielens_rating.dat`` and ``sample_movielens_merged.csv``
  (C1's data sets)

into ``oracle/_ref/`` (git-ignored: reference sources never enter this repository's history; NOT
gpurun-ignored: the directory ships with the snapshot like the built ``.so``).  ``oracle/ref_loader``
imports the package from there when ``/root/reference`` is absent, with the same two stub modules
(``tensorflow``, ``gensim``) — so ``tests/test_gpu_dropin.py`` can drive the reference's own classes
on the GPU box and ``bench.py --impl reference`` can time the reference's own
``recommend_from_embedding`` (``cpu_baseline.kind = "reference"``).

A manifest with the SHA-256 of every staged file is written beside them; ``verify()`` re-checks it, so
a staged tree that was edited is detected.

    python oracle/make_ref.py            # stage (idempotent)
    python oracle/make_ref.py --verify   # check the staged tree against its manifest
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC_ROOT = os.environ.get("B200RECO_REFERENCE", "/root/reference")
DATA_FILES = ["examples/sample_data/sample_movielens_rating.dat", "examples/sample_data/sample_movielens_merged.csv"]


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def stage() -> str:
    src_pkg = os.path.join(SRC_ROOT, "libreco")
    if not os.path.isdir(src_pkg):
        raise RuntimeError(f"reference tree not present at {SRC_ROOT}")
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    manifest = {}
    for root, _dirs, files in os.walk(src_pkg):
        for fn in files:
            if not fn.endswith(".py"):
                continue
            s = os.path.join(root, fn)
            rel = os.path.relpath(s, SRC_ROOT)
            d = os.path.join(DEST, rel)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
            manifest[rel] = _sha(d)
    for rel in DATA_FILES:
        d = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(os.path.join(SRC_ROOT, rel), d)
        manifest[rel] = _sha(d)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC_ROOT, "files": manifest}, f, indent=0, sort_keys=True)
    return DEST


def staged() -> bool:
    return os.path.isfile(os.path.join(DEST, "MANIFEST.json"))


def verify() -> bool:
    """True when every staged file still has the hash recorded at staging time (and, when the
    original tree is mounted, the hash of the original file)."""
    if not staged():
        return False
    man = json.load(open(os.path.join(DEST, "MANIFEST.json")))["files"]
    for rel, h in man.items():
        p = os.path.join(DEST, rel)
        if not os.path.isfile(p) or _sha(p) != h:
            return False
        orig = os.path.join(SRC_ROOT, rel)
        if os.path.isfile(orig) and _sha(orig) != h:
            return False
    return True


if __name__ == "__main__":
    if "--verify" in sys.argv:
        ok = verify()
        print("oracle/_ref verified" if ok else "oracle/_ref missing or modified")
        sys.exit(0 if ok else 1)
    print(stage())
