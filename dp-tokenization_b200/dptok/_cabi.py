"""ctypes binding of include/dptok.h (the stub INTEGRATION.md shows).

The shared library is built in-tree by ``__graft_entry__.build()`` /
``make -C dp-tokenization_b200/csrc`` into ``dp-tokenization_b200/lib``.
There is NO fallback: if the library is missing, importing this module fails.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libdptok.so")
if os.environ.get("DPT_LIB_PATH"):  # development: a tuning variant of the same library (tools/variants.py)
    LIB_PATH = os.path.abspath(os.environ["DPT_LIB_PATH"])

OK, EINVAL, ECUDA, ECAPACITY, ENOMEM, ESTATE = range(6)
UNIT_BYTES, UNIT_CODEPOINTS = 0, 1
RULE_PRESPLIT, RULE_SPM_LLAMA, RULE_GPT2, RULE_LLAMA3, RULE_BLOOM = range(5)
WF_UNTOKENIZABLE, WF_DOC_FIRST, WF_LONG = 1, 2, 4
DF_AMBIGUOUS = 1
CTR_BYTES, CTR_WORDS, CTR_TOKENS, CTR_UNTOKENIZABLE = range(4)
NOUT_IDS, NOUT_WORDS, NOUT_POOL_REQ, NOUT_POOL_CAP, NOUT_NORM_REQ, NOUT_NORM_CAP, NOUT_ODD_REQ, NOUT_ODD_CAP = range(8)


class DptError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"dptok error {code}: {msg}")
        self.code = code


class VocabInfo(C.Structure):
    _fields_ = [("n_tokens", C.c_int32), ("unit_mode", C.c_int32), ("n_nodes", C.c_int32),
                ("n_slots", C.c_int32), ("max_token_bytes", C.c_int32), ("ph_buckets", C.c_int32),
                ("ph_slots", C.c_int32), ("marker_leading_only", C.c_int32), ("byte_fallback", C.c_int32),
                ("device", C.c_int32), ("blob_bytes", C.c_int64)]


class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("words_lo", C.c_int32), ("words_hi", C.c_int32), ("sentence_mean", C.c_int32),
                ("flags", C.c_int32), ("frac_b", C.c_uint32), ("suffix_prob", C.c_uint32)]


if not os.path.isfile(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build the CUDA extension first "
        "(python -c 'import __graft_entry__ as g; g.build()' or make -C dp-tokenization_b200/csrc). "
        "dptok has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

_p, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64

SIGNATURES = {
    # name: (restype, argtypes)
    "dpt_vocab_create": (C.c_int, [_p, _p, _p, _i32, _i32, C.POINTER(_p)]),
    "dpt_vocab_destroy": (None, [_p]),
    "dpt_vocab_get_info": (C.c_int, [_p, C.POINTER(VocabInfo)]),
    "dpt_vocab_lookup": (C.c_int, [_p, C.c_char_p, _i32, C.POINTER(_i32)]),
    "dpt_vocab_serialize": (C.c_int, [_p, _p, _i64, C.POINTER(_i64)]),
    "dpt_vocab_deserialize": (C.c_int, [_p, _i64, C.POINTER(_p)]),
    "dpt_vocab_set_merges": (C.c_int, [_p, _p, _p, _p, _i32]),
    "dpt_vocab_upload": (C.c_int, [_p, C.c_int]),
    "dpt_pretokenize_workspace": (_i64, [_i64, _i64]),
    "dpt_encode_words_workspace": (_i64, [_i64, _i64, _i32]),
    "dpt_encode_corpus_workspace": (_i64, [_i32, _i64, _i64, _i64, _i32]),
    "dpt_encode_corpus_general_workspace": (_i64, [_i32, _i64, _i64, _i64, _i32]),
    "dpt_pretokenize": (C.c_int, [_p, _i32, _p, _i64, _p, _i64, _p, _i64, _p, _p, _i64, _p, _p, _p, _p, _i64, _p]),
    "dpt_encode_words": (C.c_int, [_p, _p, _p, _i64, _i64, _p, _i64, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "dpt_encode_corpus": (C.c_int, [_p, _i32, _p, _i64, _p, _i64, _p, _i64, _p, _p, _i64, _p, _p, _p, _p, _p, _i64,
                                    _i32, _p]),
    "dpt_corpus_table_workspace": (_i64, [_i64, _i64, _i32]),
    "dpt_encode_corpus_range_workspace": (_i64, [_i32, _i64, _i64, _i64, _i32]),
    "dpt_encode_corpus_range": (C.c_int, [_p, _i32, _p, _i64, _p, _i64, _i64, _i64, _i64, _i64, _i32, _i64, _p, _i64, _p, _p,
                                          _i64, _p, _p, _p, _p, _p, _i64, _p, _i64, _i32, _i32, _p]),
    "dpt_encode_corpus_general": (C.c_int, [_p, _i32, _p, _i64, _p, _i64, _p, _i64, _p, _p, _i64, _p, _p, _p, _p, _p,
                                            _i64, _i32, _p]),
    "dpt_lattice_word": (C.c_int, [_p, _p, _i32, _p, _p, _p, _p, _i32, _p, _p, _p]),
    "dpt_min_tokens_word": (C.c_int, [_p, _p, _i32, _p, _p, _p, _p]),
    "dpt_roundtrip_check": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i32, _p, _p]),
    "dpt_narrow_ids_u16": (C.c_int, [_p, _p, _i64, _p, _p, _p]),
    "dpt_pad_batch": (C.c_int, [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _i32, _p, _p, _p, _p]),
    "dpt_synth_corpus": (C.c_int, [_p, _p, _p, _i32, _p, _p, _p, _i32, _p, _i64, _i64, _p, _p, _p, _p]),
    "dpt_last_error": (C.c_char_p, []),
    "dpt_version": (C.c_char_p, []),
    "dpt_launch_count": (_i64, []),
    "dpt_profile_enable": (None, [_i32]),
    "dpt_profile_report": (C.c_int, [C.c_char_p, _i64, C.POINTER(_i64)]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)   # AttributeError here = header and library out of sync
    _fn.restype = _res
    _fn.argtypes = _args


def check(rc: int):
    if rc != OK:
        raise DptError(rc, lib.dpt_last_error().decode("utf-8", "replace"))
