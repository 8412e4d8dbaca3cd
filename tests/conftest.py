import gzip
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "dp-tokenization_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    path = os.path.join(GOLDEN, name)
    if name.endswith(".gz"):
        with gzip.open(path, "rt", encoding="utf-8") as f:
            return json.load(f)
    with open(path, "r", encoding="utf-8") as f:
        return json.load(f)


@pytest.fixture(scope="session")
def product_lib():
    """Build (if needed) and return the path of the product shared library."""
    lib = os.path.join(PKG, "lib", "libdptok.so")
    if not os.path.isfile(lib):
        subprocess.check_call(["make", "-C", os.path.join(PKG, "csrc")])
    return lib


def build_host_sim(extra=(), name="libsim.so"):
    """TEST-ONLY g++ build of the device headers (tests/host_sim/sim.cpp); `extra`: the kernels' tuning macros."""
    import ctypes as C
    out_dir = os.path.join(ROOT, "tests", "host_sim", "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, name)
    srcs = [os.path.join(ROOT, "tests", "host_sim", "sim.cpp"), os.path.join(PKG, "csrc", "vocab.cpp")]
    deps = srcs + [os.path.join(PKG, "csrc", h) for h in os.listdir(os.path.join(PKG, "csrc")) if h.endswith(".h")]
    extra = list(extra)
    if not os.path.isfile(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++20", "-pthread", "-shared", "-fPIC", "-I", os.path.join(PKG, "csrc"),
                               "-o", out] + extra + srcs)
    lib = C.CDLL(out)
    lib.sim_vocab_create.restype = C.c_void_p
    lib.sim_vocab_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
    lib.sim_vocab_destroy.argtypes = [C.c_void_p]
    lib.sim_vocab_set_merges.restype = C.c_int32
    lib.sim_vocab_set_merges.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    lib.sim_lookup.argtypes = [C.c_void_p, C.c_char_p, C.c_int32]
    lib.sim_word.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    lib.sim_info.argtypes = [C.c_void_p, C.c_void_p]
    lib.sim_spm_normalise.restype = C.c_int64
    lib.sim_spm_normalise.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.sim_encode_corpus_pipe.restype = C.c_int32
    lib.sim_encode_corpus_pipe.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                           C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64,
                                           C.c_int64, C.c_int64, C.c_int32]
    lib.sim_roundtrip.restype = None
    lib.sim_roundtrip.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                  C.c_int32]
    return lib


@pytest.fixture(scope="session")
def host_sim():
    extra = os.environ.get("DPT_SIM_EXTRA", "").split()  # development: -DDPT_PA_T=... etc.
    return build_host_sim(extra, "libsim_variant.so" if extra else "libsim.so")


@pytest.fixture(scope="session")
def host_sim_noskip():
    """The same kernels with the split scanner walking letter runs character by character (no letter mask)."""
    return build_host_sim(["-DDPT_NO_LETTER_SKIP"], "libsim_noskip.so")

