"""ctypes wrapper of oracle/dp_oracle.c.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
from typing import Dict

import numpy as np

from . import build as _build

_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.orc_vocab_new.restype = C.c_void_p
        _lib.orc_vocab_new.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
        _lib.orc_vocab_free.argtypes = [C.c_void_p]
        _lib.orc_encode_words.restype = C.c_int64
        _lib.orc_encode_words.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                          C.c_void_p, C.c_void_p]
    return _lib


class COracle:
    def __init__(self, vocab: Dict[bytes, int], unit_mode: int):
        lib = _load()
        toks = [(t, i) for t, i in vocab.items() if len(t) > 0]
        toks.sort(key=lambda x: x[1])
        blob = np.frombuffer(b"".join(t for t, _ in toks) + b"\0", dtype=np.uint8)
        offs = np.zeros(len(toks) + 1, dtype=np.int64)
        np.cumsum([len(t) for t, _ in toks], out=offs[1:])
        ids = np.asarray([i for _, i in toks], dtype=np.int32)
        self._h = C.c_void_p(lib.orc_vocab_new(blob.ctypes.data, offs.ctypes.data, ids.ctypes.data, len(toks), unit_mode))

    def encode_words(self, text: np.ndarray, word_offs: np.ndarray):
        """-> (ids int32[n], word_lens int32[n_words], untokenizable uint8[n_words])"""
        lib = _load()
        text = np.ascontiguousarray(text, dtype=np.uint8)
        word_offs = np.ascontiguousarray(word_offs, dtype=np.int64)
        n_words = len(word_offs) - 1
        cap = int(word_offs[-1] - word_offs[0]) + 8
        ids = np.empty(cap, dtype=np.int32)
        lens = np.empty(max(n_words, 1), dtype=np.int32)
        flags = np.empty(max(n_words, 1), dtype=np.uint8)
        n = lib.orc_encode_words(self._h, text.ctypes.data, word_offs.ctypes.data, n_words, ids.ctypes.data, cap,
                                 lens.ctypes.data, flags.ctypes.data)
        return ids[:n].copy(), lens[:n_words], flags[:n_words]

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.orc_vocab_free(h)
