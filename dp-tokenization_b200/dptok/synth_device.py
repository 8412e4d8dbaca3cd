"""Synthetic corpora generated on the device (``dpt_synth_corpus``): measurement support for the configurations whose
text must not cross PCIe (BASELINE.json configs[3]: 1 GB of 8-64 KB documents; configs[4]: 10 GB of Arabic-script text over
8 GPUs).  A document is a pure function of (seed, global document index): every rank can generate
the same corpus (strong scaling through ``sharded.shard_bounds``) or only its own document range of it.

The lexicons are the host generators' (``synth.make_lexicon`` / ``synth.make_arabic_lexicon``): Zipf(1.0) word
frequencies, the tokenizer assets were trained on text of the same generators.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi
from ._cabi import lib, check


class DeviceLexicon:
    """Word strings + 32-bit cumulative Zipf table, resident on one GPU."""

    def __init__(self, words: Sequence[str], device: int, zipf_s: float = 1.0):
        enc = [w.encode("utf-8") for w in words]
        offs = np.zeros(len(enc) + 1, dtype=np.int64)
        np.cumsum([len(w) for w in enc], out=offs[1:])
        p = 1.0 / np.arange(1, len(enc) + 1, dtype=np.float64) ** zipf_s
        cdf = np.cumsum(p / p.sum())
        cdf32 = np.minimum(np.floor(cdf * 4294967296.0), 4294967295.0).astype(np.uint32)
        cdf32[-1] = 0xFFFFFFFF
        self.n = len(enc)
        # Latin-script lexicon (sentence starts are capitalised; only an ASCII lowercase first letter is changed)
        self.ascii = sum(1 for w in enc[:2000] if w[:1].isascii()) > 1800
        self.mean_len = float(np.sum(np.diff(offs) * p / p.sum()))  # frequency-weighted mean word length (bytes)
        self.bytes = torch.from_numpy(np.frombuffer(b"".join(enc), dtype=np.uint8).copy()).to(device)
        self.offs = torch.from_numpy(offs).to(device)
        self.cdf = torch.from_numpy(cdf32.view(np.int32).copy()).to(device)


def generate(lex_a: DeviceLexicon, n_docs: int, seed: int, words_per_doc: Tuple[int, int], device: int,
             lex_b: Optional[DeviceLexicon] = None, frac_b: float = 0.0, sentence_mean: int = 22, doc_base: int = 0,
             suffix_prob: float = 0.0):
    """-> (text uint8[N] device, doc_offs int64[n_docs + 1] device) of documents ``doc_base .. doc_base + n_docs``."""
    sp = _cabi.SynthParams(seed=seed, words_lo=words_per_doc[0], words_hi=words_per_doc[1], sentence_mean=sentence_mean,
                           flags=(1 if lex_a.ascii else 0) | (2 if (lex_b is not None and lex_b.ascii) else 0),
                           frac_b=min(int(frac_b * 4294967296.0), 0xFFFFFFFF) if lex_b is not None else 0,
                           suffix_prob=min(int(suffix_prob * 4294967296.0), 0xFFFFFFFF))
    st = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)

    def ptrs(L):
        if L is None:
            return None, None, None, 0
        return C.c_void_p(L.bytes.data_ptr()), C.c_void_p(L.offs.data_ptr()), C.c_void_p(L.cdf.data_ptr()), L.n

    a, b = ptrs(lex_a), ptrs(lex_b)
    with torch.cuda.device(device):
        lens = torch.empty(n_docs, dtype=torch.int64, device=device)
        check(lib.dpt_synth_corpus(*a, *b, C.byref(sp), doc_base, n_docs, C.c_void_p(lens.data_ptr()), None, None, st))
        offs = torch.zeros(n_docs + 1, dtype=torch.int64, device=device)
        torch.cumsum(lens, 0, out=offs[1:])
        total = int(offs[-1].item())
        text = torch.empty(total, dtype=torch.uint8, device=device)
        check(lib.dpt_synth_corpus(*a, *b, C.byref(sp), doc_base, n_docs, None, C.c_void_p(offs.data_ptr()),
                                   C.c_void_p(text.data_ptr()), st))
    return text, offs


def docs_for_bytes(n_bytes: int, words_per_doc: Tuple[int, int], mean_word_bytes: float) -> int:
    """Number of documents that gives about ``n_bytes`` of text."""
    per_doc = 0.5 * (words_per_doc[0] + words_per_doc[1]) * (mean_word_bytes + 1.25)
    return max(1, int(round(n_bytes / per_doc)))


# ---- host port of the generator (bit-identical; a few MB at Python speed) ------------------------------------------------
# The reference arm of bench.py (--impl reference) runs without a GPU: it needs the first documents of the SAME corpus the
# GPU arm generates on the device.  tests/test_gpu_synth.py compares the two.
_M64 = (1 << 64) - 1


def _mix(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


class HostLexicon:
    def __init__(self, words: Sequence[str], zipf_s: float = 1.0):
        self.enc = [w.encode("utf-8") for w in words]
        p = 1.0 / np.arange(1, len(self.enc) + 1, dtype=np.float64) ** zipf_s
        cdf = np.cumsum(p / p.sum())
        self.cdf = np.minimum(np.floor(cdf * 4294967296.0), 4294967295.0).astype(np.uint32)
        self.cdf[-1] = 0xFFFFFFFF
        self.n = len(self.enc)
        self.ascii = sum(1 for w in self.enc[:2000] if w[:1].isascii()) > 1800

    def draw(self, u: int) -> int:
        return min(int(np.searchsorted(self.cdf, np.uint32(u), side="left")), self.n - 1)


def generate_host(lex_a: HostLexicon, n_docs: int, seed: int, words_per_doc: Tuple[int, int],
                  lex_b: Optional[HostLexicon] = None, frac_b: float = 0.0, sentence_mean: int = 22, doc_base: int = 0,
                  suffix_prob: float = 0.0):
    """-> list of document byte strings, identical to ``generate`` on the device."""
    flags = (1 if lex_a.ascii else 0) | (2 if (lex_b is not None and lex_b.ascii) else 0)
    fb = min(int(frac_b * 4294967296.0), 0xFFFFFFFF) if lex_b is not None else 0
    sfx = min(int(suffix_prob * 4294967296.0), 0xFFFFFFFF)
    lo, hi = words_per_doc

    def sentence_end(doc, w, nw):
        if w < 0 or w == nw - 1:
            return True
        return ((_mix(seed ^ 0xA5A5A5A5 ^ ((doc << 24) & _M64) ^ (w & 0xFFFFFFFF)) >> 33) & 0xFFFFFFFF) % sentence_mean == 0

    def put_lex(L, k, cap):
        w = L.enc[k]
        if cap and w and 97 <= w[0] <= 122:
            w = bytes([w[0] - 32]) + w[1:]
        return w

    docs = []
    for d in range(n_docs):
        doc = doc_base + d
        use_b = lex_b is not None and ((_mix(seed ^ 0xB10B ^ ((doc * 3) & _M64)) >> 32) & 0xFFFFFFFF) < fb
        L = lex_b if use_b else lex_a
        fl = (flags >> 1) if use_b else flags
        nw = lo + ((_mix(seed ^ 0x5EED ^ ((doc << 1) & _M64)) >> 32) & 0xFFFFFFFF) % (hi - lo + 1)
        parts = []
        for w in range(nw):
            r = _mix(seed ^ ((doc * 0x100000001B3) & _M64) ^ ((w & 0xFFFFFFFF) << 1))
            r2 = _mix(r)
            cap = bool(fl & 1) and sentence_end(doc, w - 1, nw)
            kind = r2 & 0xFFFF
            k = L.draw((r >> 32) & 0xFFFFFFFF)
            if kind < 1966:
                v = (r2 >> 16) & 3
                if v == 0:
                    tok = b"%d" % (((r2 >> 20) & 0xFFFFFFFF) % 3000)
                elif v == 1:
                    tok = b"%d.%d%%" % (((r2 >> 20) & 0xFFFFFFFF) % 100, ((r2 >> 40) & 0xFFFFFFFF) % 10)
                elif v == 2:
                    tok = b"(" + put_lex(L, k, False) + b")"
                else:
                    tok = put_lex(L, k, cap) + b"-" + put_lex(L, L.draw((r2 >> 32) & 0xFFFFFFFF), False)
            else:
                tok = put_lex(L, k, cap)
                if sfx:
                    r3 = _mix(r2 ^ 0x5AFF)
                    if ((r3 >> 32) & 0xFFFFFFFF) < sfx or sfx == 0xFFFFFFFF:
                        tok += bytes(97 + ((r3 >> (5 * q)) & 31) % 26 for q in range(4))
            if sentence_end(doc, w, nw):
                tok += b"."
            elif ((r2 >> 48) & 15) == 1:
                tok += b"," if (r2 >> 52) & 7 else b";"
            parts.append(tok)
        docs.append(b" ".join(parts))
    return docs
