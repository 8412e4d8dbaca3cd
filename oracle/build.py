"""Compile oracle/dp_oracle.c -> oracle/_build/liboracle.so (gcc).  Building the checker is not using it.

The reference itself is pure Python (no C/C++ sources), so there is no ``oracle/_ref`` binary to build:
``/root/reference`` is imported directly by ``oracle/ref_harness.py`` where it exists.
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "dp_oracle.c")
    if not force and os.path.isfile(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(src):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-shared", "-fPIC", "-o", OUT, src])
    return OUT


if __name__ == "__main__":
    print(build(force=True))
