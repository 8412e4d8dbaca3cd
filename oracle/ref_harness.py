"""Load the UNMODIFIED reference (``/root/reference``) for fixture generation.

TEST INFRASTRUCTURE ONLY - used by ``tests/golden/make_golden.py`` and by the
optional cross-check in ``tests/test_oracle.py`` (skipped when the reference
tree is absent, e.g. on the GPU box).  Nothing is copied: the reference
modules are imported from where they lie under an alias package name.

Shims:
  * ``ipdb`` is not installed  -> stub whose ``set_trace`` raises.
  * ``bidict`` is not installed -> dict subclass with ``.inverse``.
  * tokenizer_utils.py:71 passes a 5th positional argument that
    dp_tokenize.py:6-11 does not accept -> the alias module's
    ``compute_shortest_tokenizations`` is wrapped to drop it.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DPT_REFERENCE_ROOT", "/root/reference")
_ALIAS = "dpt_reference_packages"


class ReferenceUnavailable(RuntimeError):
    pass


class _BiDict(dict):
    @property
    def inverse(self):
        inv = self.__dict__.get("_inv")
        if inv is None or len(inv) != len(self):
            inv = {v: k for k, v in self.items()}
            self.__dict__["_inv"] = inv
        return inv


def _install_shims():
    if "ipdb" not in sys.modules:
        try:
            importlib.import_module("ipdb")
        except ImportError:
            stub = types.ModuleType("ipdb")

            def set_trace(*_a, **_k):
                raise RuntimeError("reference called ipdb.set_trace()")

            stub.set_trace = set_trace
            sys.modules["ipdb"] = stub
    if "bidict" not in sys.modules:
        try:
            importlib.import_module("bidict")
        except ImportError:
            stub = types.ModuleType("bidict")
            stub.bidict = _BiDict
            sys.modules["bidict"] = stub


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "packages", "dp_tokenize.py"))


def load():
    """Returns (dp_tokenize module, tokenizer_utils module) of the reference."""
    if not available():
        raise ReferenceUnavailable(f"no reference tree at {REFERENCE_ROOT}")
    if _ALIAS + ".tokenizer_utils" in sys.modules:
        return sys.modules[_ALIAS + ".dp_tokenize"], sys.modules[_ALIAS + ".tokenizer_utils"]
    _install_shims()
    pkg_dir = os.path.join(REFERENCE_ROOT, "packages")
    spec = importlib.util.spec_from_file_location(
        _ALIAS, os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules[_ALIAS] = pkg
    spec.loader.exec_module(pkg)
    dp = importlib.import_module(_ALIAS + ".dp_tokenize")
    tu = importlib.import_module(_ALIAS + ".tokenizer_utils")
    four_arg = dp.compute_shortest_tokenizations
    tu.compute_shortest_tokenizations = lambda a, b, c, d, *_extra: four_arg(a, b, c, d)
    return dp, tu


def load_min_tokens():
    """inspect_tokenizer.py:77-86 (the second, shadowing definition)."""
    if not available():
        raise ReferenceUnavailable(f"no reference tree at {REFERENCE_ROOT}")
    _install_shims()
    spec = importlib.util.spec_from_file_location(
        "dpt_reference_inspect", os.path.join(REFERENCE_ROOT, "inspect_tokenizer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.min_tokens_for_string
