"""Stage the UNMODIFIED reference modules for the CPU baseline on the GPU box.  TEST/BENCH INFRASTRUCTURE ONLY.

``/root/reference`` does not exist on the GPU box, and the reference is pure Python (nothing to compile), so the two
modules of the hot path - ``packages/dp_tokenize.py`` and ``packages/tokenizer_utils.py`` (+ ``__init__.py``) - are
copied byte for byte into the git-ignored ``oracle/_ref/packages/`` (an artefact like a built ``.so``: it travels with
the gpurun snapshot, it is never committed).  ``oracle/cpu_baseline.py`` imports them from there through
``oracle/ref_harness.py`` (shims for ipdb / bidict / the 5-argument call) and reports ``kind: "reference"``; without
the staged copy it falls back to the oracle port (``kind: "port"``).

    python -m oracle.make_ref          # run by __graft_entry__.build() when /root/reference is present
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

SRC = os.environ.get("DPT_REFERENCE_SRC", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["packages/__init__.py", "packages/dp_tokenize.py", "packages/tokenizer_utils.py"]


def staged() -> bool:
    return all(os.path.isfile(os.path.join(DST, f)) for f in FILES)


def make() -> bool:
    if not all(os.path.isfile(os.path.join(SRC, f)) for f in FILES):
        return staged()
    manifest = {}
    for f in FILES:
        dst = os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, f), dst)
        with open(dst, "rb") as fh:
            manifest[f] = hashlib.sha1(fh.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "sha1": manifest}, fh, indent=1)
    return True


if __name__ == "__main__":
    print("staged" if make() else "reference not available", DST)
