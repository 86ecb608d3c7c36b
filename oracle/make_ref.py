"""Recipe that stages the UNMODIFIED reference next to the oracle so that it can be TIMED on the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference is pure Python: "building" it means
copying the files of the hot path, byte for byte, from where they lie under ``/root/reference`` into
``oracle/_ref/reference/``.  ``oracle/_ref/`` is git-ignored (reference sources never enter the history) but not
gpurun-ignored, so the staged copy travels to the GPU box like a built ``.so`` does, where ``bench.py --impl
reference`` and the ``cpu_baseline`` leg time it on the box's host cores (``kind: "reference"``).  Nothing in the
product, in the ``-m gpu`` tests or in ``smoke()`` reads it.

    python -m oracle.make_ref          # also run by __graft_entry__.build() when /root/reference is present

A manifest with the SHA-256 of every staged file is written next to the copy; ``verify()`` re-checks it before a
timing run so that the number is known to come from unmodified files.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("D2D_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref", "reference")
MANIFEST = os.path.join(HERE, "_ref", "MANIFEST.json")
FILES = ["__init__.py", "envs/env.py", "envs/combinatorial_env.py", "envs/channel_selection_env.py",
         "algorithms/baselines.py", "algorithms/ippo.py", "algorithms/d2d_ppo.py", "algorithms/irdqn.py",
         "combinatorial_load/setup.p", "combinatorial_load/setup_8_channels.p", "combinatorial_load/channel_switch_8.p"]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def available():
    return os.path.isfile(MANIFEST) and os.path.isfile(os.path.join(DST, "envs", "combinatorial_env.py"))


def stage():
    if not os.path.isfile(os.path.join(SRC, "envs", "combinatorial_env.py")):
        return None
    manifest = {}
    for rel in FILES:
        src = os.path.join(SRC, rel)
        if not os.path.isfile(src):
            continue
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    for pkg in ("envs", "algorithms"):          # the reference has no package __init__ files in these (namespace pkgs)
        init = os.path.join(SRC, pkg, "__init__.py")
        if os.path.isfile(init):
            shutil.copyfile(init, os.path.join(DST, pkg, "__init__.py"))
            manifest[f"{pkg}/__init__.py"] = _sha(os.path.join(DST, pkg, "__init__.py"))
    with open(MANIFEST, "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1, sort_keys=True)
    return DST


def verify():
    """True iff every staged file still has the hash recorded when it was copied from the reference tree."""
    if not available():
        return False
    files = json.load(open(MANIFEST))["files"]
    return all(os.path.isfile(os.path.join(DST, rel)) and _sha(os.path.join(DST, rel)) == h for rel, h in files.items())


if __name__ == "__main__":
    print(stage() or f"reference tree not found at {SRC}")
