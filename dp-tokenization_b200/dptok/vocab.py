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
import tempfile
from typing import Dict, Iterable, Mapping

import numpy as np

from . import _cabi
from ._cabi import DptError, lib, check

_CACHE_FORMAT = b"dptok-cache-2"  # bump when the serialised layout or the Unicode class tables change


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

    def set_merges(self, merges) -> "CompiledVocab":
        """``merges``: the tokenizer's BPE merges in rank order as (left id, right id, merged id) triples
        (``dpt_vocab_set_merges``; SentencePiece-style vocabularies, before the first upload).  With them the device
        rule cuts runs of U+2581 / spaces like the reference's tokenizer-driven split (tokenizer_utils.py:7-31)."""
        m = np.ascontiguousarray(np.asarray(list(merges), dtype=np.int32).reshape(-1, 3))
        left, right, merged = (np.ascontiguousarray(m[:, k]) for k in range(3))
        check(lib.dpt_vocab_set_merges(self._h, left.ctypes.data, right.ctypes.data, merged.ctypes.data, len(m)))
        self.n_merges = len(m)
        return self

    @staticmethod
    def merges_of(tokenizer, token_to_id: Mapping[str, int]):
        """(left id, right id, merged id) triples from a HF fast tokenizer's ``model.merges``; None when the tokenizer has
        no ``tokenizers`` backend or is no BPE model."""
        import json
        backend = getattr(tokenizer, "backend_tokenizer", None) or getattr(tokenizer, "_tokenizer", None)
        if backend is None or not hasattr(backend, "to_str"):
            return None
        try:
            model = json.loads(backend.to_str()).get("model", {})
        except Exception:
            return None
        if model.get("type") != "BPE" or model.get("dropout") or model.get("ignore_merges"):
            return None
        out = []
        for m in model.get("merges", []):
            a, b = m.split(" ", 1) if isinstance(m, str) else m
            if a in token_to_id and b in token_to_id and (a + b) in token_to_id:
                out.append((token_to_id[a], token_to_id[b], token_to_id[a + b]))
            else:
                return None  # a merge the vocabulary cannot express: the device would not reproduce the tokenizer
        return out

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
        # the key names the library build too: a cache written by another trie layout / Unicode table is never read
        h.update(lib.dpt_version() + b"\0" + _CACHE_FORMAT + b"\0")
        for t, i in sorted(token_to_id.items()):
            h.update(t.encode("utf-8"))
            h.update(b"\0%d\n" % i)
        path = os.path.join(cache_dir, f"dptok-{family}-{h.hexdigest()[:24]}.bin")
        if os.path.isfile(path):
            try:
                with open(path, "rb") as f:
                    blob = f.read()
                # file = sha256(payload) + payload: a partly written or damaged file is rebuilt, not trusted
                if len(blob) > 32 and hashlib.sha256(blob[32:]).digest() == blob[:32]:
                    return cls.deserialize(blob[32:], family, dict(token_to_id))
            except (OSError, DptError):
                pass
        v = cls.from_token_map(token_to_id, family)
        try:
            os.makedirs(cache_dir, exist_ok=True)
            payload = v.serialize()
            # one process per GPU compiles the same vocabulary at the same time: every writer gets its own temporary file
            fd, tmp = tempfile.mkstemp(prefix=os.path.basename(path) + ".", suffix=".tmp", dir=cache_dir)
            try:
                with os.fdopen(fd, "wb") as f:
                    f.write(hashlib.sha256(payload).digest())
                    f.write(payload)
                os.replace(tmp, path)
            except BaseException:
                try:
                    os.unlink(tmp)
                except OSError:
                    pass
                raise
        except OSError:
            pass  # an unwritable cache directory costs a recompilation next time, nothing else
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
