"""Shared test helpers (vocab conversion, word packing)."""
import hashlib
import json

import numpy as np


def bytelevel_table():
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    table, extra = {}, 0
    for b in range(256):
        if b in keep:
            table[b] = chr(b)
        else:
            table[b] = chr(256 + extra)
            extra += 1
    return table


_U2B = {c: b for b, c in bytelevel_table().items()}


def vocab_bytes(token_to_id, family):
    """HF token->id map -> {raw bytes: id} exactly as dptok.vocab does, but written independently."""
    out = {}
    for t, i in token_to_id.items():
        if t == "":
            continue
        if family == "spm":
            out[t.encode("utf-8")] = i
        else:
            try:
                out[bytes(_U2B[c] for c in t)] = i
            except KeyError:
                pass
    return out


def pack(words):
    offs = np.zeros(len(words) + 1, dtype=np.int64)
    np.cumsum([len(w) for w in words], out=offs[1:])
    return np.frombuffer(b"".join(words) + b"\0", dtype=np.uint8)[:-1].copy(), offs


def sha1_json(obj):
    return hashlib.sha1(json.dumps(obj, ensure_ascii=False, separators=(",", ":")).encode("utf-8")).hexdigest()


def make_sim_vocab(sim, vocab, unit_mode):
    toks = list(vocab.items())
    blob = np.frombuffer(b"".join(t for t, _ in toks) + b"\0", np.uint8)
    offs = np.zeros(len(toks) + 1, np.int64)
    offs[1:] = np.cumsum([len(t) for t, _ in toks])
    ids = np.array([i for _, i in toks], np.int32)
    h = sim.sim_vocab_create(blob.ctypes.data, offs.ctypes.data, ids.ctypes.data, len(toks), unit_mode)
    assert h, "vocab compile failed"
    return h


def sim_word(sim, h, data, bnd=None):
    import ctypes
    out = np.zeros(len(data) + 1, np.int32)
    wl = ctypes.c_int32()
    us = None
    if bnd is not None:
        a = np.zeros(len(data), np.uint8)
        a[[p for p in bnd if p < len(data)]] = 1
        us = a.ctypes.data
    r = sim.sim_word(h, data, len(data), us, out.ctypes.data, len(out), ctypes.byref(wl))
    return r, wl.value, out[:max(r, 0)].tolist()


def py_roundtrip_ok(id2tok, ids, doc, spm, skip_bos):
    """ids of one document -> bytes the way tokenizer.decode does (tokenizer_utils.py:82-84, :176-179; id2tok: id -> token
    bytes of the compiled vocabulary) == the document?  The test-side statement of dpt_roundtrip_check."""
    out = b""
    for k, i in enumerate(ids[1:] if skip_bos else ids):
        t = id2tok.get(int(i))
        if t is None:
            return False
        if spm and len(t) == 6 and t[:3] == b"<0x" and t[5:] == b">" and all(c in b"0123456789ABCDEF" for c in t[3:5]):
            out += bytes([int(t[3:5], 16)])
            continue
        piece = t.replace("\u2581".encode(), b" ") if spm else t
        if spm and k == 0 and piece[:1] == b" ":
            piece = piece[1:]
        out += piece
    return out == doc
