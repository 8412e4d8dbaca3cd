"""Vocabulary compiler front-end: tokenizer vocab -> GPU-resident double-array trie + perfect hash.

Replaces ``vocab = set(tok.get_vocab())`` (tokenizer_utils.py:57) and the
``vocab_to_index`` bidict (tokenizer_utils.py:105-113).  Token strings become
canonical byte strings:

* ``family='spm'``       - UTF-8 of the HF key; DP units are code points.
* ``family='bytelevel'`` - inverse GPT-2 ``bytes_to_unicode`` of the HF key -> raw bytes; DP units are
  bytes (one mapped char of tokenizer_utils.py:149 == one raw byte).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
from typing import Dict, Iterable, Mapping

import numpy as np

from . import _cabi
from ._cabi import lib, check


def bytes_to_unicode() -> Dict[int, str]:
    """The GPT-2 byte -> printable-char table used by every byte-level BPE (ByteLevel pre-tokenizer)."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    table, extra = {}, 0
    for b in range(256):
        if b in keep:
            table[b] = chr(b)
        else:
            table[b] = chr(256 + extra)
            extra += 1
    return table


_B2U = bytes_to_unicode()
_U2B = {c: b for b, c in _B2U.items()}


def bytelevel_to_bytes(token: str) -> bytes:
    """Mapped string -> raw bytes.  Raises KeyError for a char outside the byte alphabet."""
    return bytes(_U2B[c] for c in token)


def bytes_to_bytelevel(raw: bytes) -> str:
    return "".join(_B2U[b] for b in raw)


class CompiledVocab:
    """Owns a ``dpt_vocab`` handle.  Immutable after ``upload``; shareable across threads/streams."""

    def __init__(self, handle, family: str, token_to_id: Mapping[str, int] | None = None):
        self._h = C.c_void_p(handle)
        self.family = family
        self.token_to_id = token_to_id
        self._id_to_token = None
        info = _cabi.VocabInfo()
        check(lib.dpt_vocab_get_info(self._h, C.byref(info)))
        self.info = info

    # ---- construction ---------------------------------------------------------------
    @classmethod
    def from_bytes_map(cls, tokens: Iterable[bytes], ids: Iterable[int], unit_mode: int, family: str,
                       token_to_id=None) -> "CompiledVocab":
        toks = list(tokens)
        idv = np.asarray(list(ids), dtype=np.int32)
        if len(toks) != len(idv) or not toks:
            raise ValueError("vocabulary must be non-empty with one id per token")
        offs = np.zeros(len(toks) + 1, dtype=np.int64)
        np.cumsum([len(t) for t in toks], out=offs[1:])
        blob = np.frombuffer(b"".join(toks) + b"\0", dtype=np.uint8)
        out = C.c_void_p()
        check(lib.dpt_vocab_create(blob.ctypes.data, offs.ctypes.data, idv.ctypes.data, len(toks), unit_mode,
                                   C.byref(out)))
        return cls(out.value, family, token_to_id)

    @classmethod
    def from_token_map(cls, token_to_id: Mapping[str, int], family: str) -> "CompiledVocab":
        """``token_to_id``: HF-style mapping of token STRING -> id (``tok.get_vocab()``)."""
        if family == "spm":
            items = [(t.encode("utf-8"), i) for t, i in token_to_id.items() if t != ""]
            mode = _cabi.UNIT_CODEPOINTS
        elif family == "bytelevel":
            items = []
            for t, i in token_to_id.items():
                try:
                    items.append((bytelevel_to_bytes(t), i))
                except KeyError:
                    # added/special tokens written outside the byte alphabet can never match the
                    # byte-level-mapped text the reference's DP sees (tokenizer_utils.py:149)
                    continue
            mode = _cabi.UNIT_BYTES
        else:
            raise ValueError("family must be 'spm' or 'bytelevel'")
        return cls.from_bytes_map([b for b, _ in items], [i for _, i in items], mode, family, dict(token_to_id))

    @classmethod
    def from_strings(cls, vocabulary: Iterable[str]) -> "CompiledVocab":
        """Arbitrary container of token strings (the 4-argument API, dp_tokenize.py:6-11); ids = rank."""
        uniq = []
        seen = set()
        for t in vocabulary:
            if t not in seen and t != "":
                seen.add(t)
                uniq.append(t)
        if not uniq:
            uniq = ["\U0010FFFF\U0010FFFE"]  # placeholder that never matches: keeps an empty vocab legal
        return cls.from_token_map({t: k for k, t in enumerate(uniq)}, "spm")

    # ---- serialisation (compiled-vocab cache) -----------------------
    def serialize(self) -> bytes:
        need = C.c_int64()
        check(lib.dpt_vocab_serialize(self._h, None, 0, C.byref(need)))
        buf = (C.c_uint8 * need.value)()
        check(lib.dpt_vocab_serialize(self._h, buf, need.value, C.byref(need)))
        return bytes(buf)

    @classmethod
    def deserialize(cls, data: bytes, family: str, token_to_id=None) -> "CompiledVocab":
        arr = np.frombuffer(data, dtype=np.uint8)
        out = C.c_void_p()
        check(lib.dpt_vocab_deserialize(arr.ctypes.data, len(data), C.byref(out)))
        return cls(out.value, family, token_to_id)

    @classmethod
    def cached(cls, token_to_id: Mapping[str, int], family: str, cache_dir: str | None) -> "CompiledVocab":
        if not cache_dir:
            return cls.from_token_map(token_to_id, family)
        h = hashlib.sha256()
        for t, i in sorted(token_to_id.items()):
            h.update(t.encode("utf-8"))
            h.update(b"\0%d\n" % i)
        path = os.path.join(cache_dir, f"dptok-{family}-{h.hexdigest()[:24]}.bin")
        if os.path.isfile(path):
            with open(path, "rb") as f:
                return cls.deserialize(f.read(), family, dict(token_to_id))
        v = cls.from_token_map(token_to_id, family)
        os.makedirs(cache_dir, exist_ok=True)
        with open(path + ".tmp", "wb") as f:
            f.write(v.serialize())
        os.replace(path + ".tmp", path)
        return v

    # ---- use ---------------------------------------------------------------------------
    def upload(self, device: int) -> "CompiledVocab":
        check(lib.dpt_vocab_upload(self._h, int(device)))
        check(lib.dpt_vocab_get_info(self._h, C.byref(self.info)))
        return self

    def lookup(self, raw: bytes) -> int:
        out = C.c_int32()
        check(lib.dpt_vocab_lookup(self._h, raw, len(raw), C.byref(out)))
        return out.value

    @property
    def handle(self):
        return self._h

    @property
    def id_to_token(self):
        if self._id_to_token is None and self.token_to_id is not None:
            self._id_to_token = {i: t for t, i in self.token_to_id.items()}
        return self._id_to_token

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib.dpt_vocab_destroy(h)
            except Exception:
                pass
