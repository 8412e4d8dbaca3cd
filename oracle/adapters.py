"""CPU restatement of the reference's tokenizer adapters.  TEST INFRASTRUCTURE ONLY.

* ``llama_words`` / ``llama_encode``       - tokenizer_utils.py:24-31, 52-80 ('llama' option; the stray 5th
  argument of :71 dropped)
* ``bytelevel_pieces`` / ``bytelevel_encode`` - tokenizer_utils.py:105-113, 147-174
* ``spm_normalise``                        - what the concatenated default-token strings of :26-29 spell for
  text whose word split is unambiguous (single markers): "<s>" + U+2581 + text with ' '->U+2581 and every
  out-of-vocabulary character spelled "<0xHH>" per byte.  Checked against ``llama_words`` in tests.

Parity status: PINNED by tests/golden/llama_adapter.json.gz and bytelevel_adapter.json.gz (outputs of the
unmodified reference adapters on the committed stand-in tokenizers).
"""
from __future__ import annotations

from typing import Dict, List

from . import dp_oracle

MARK = "▁"


def llama_words(tok, text: str) -> List[str]:
    inv = {i: t for t, i in tok.get_vocab().items()}
    toks = [inv[i] for i in tok.encode(text)]
    words: List[str] = []
    for k, t in enumerate(toks):
        if k == 0 or t.startswith(MARK):
            words.append(t)
        else:
            words[-1] += t
    return words


def llama_encode(tok, text: str) -> List[int]:
    t2i = tok.get_vocab()
    vocab = set(t2i)
    ids: List[int] = []
    for w in llama_words(tok, text):
        sel, _, _ = dp_oracle.select_shortest(w, vocab, max_token_units=64)
        if sel is None:
            raise ValueError("max() arg is an empty sequence")
        ids.extend(t2i[t] for t in sel)
    return ids


def spm_normalise(text: str, vocab) -> List[str]:
    """Words of the SPM_LLAMA device rule for unambiguous text (see module docstring)."""
    out = ["<s>"]
    if text == "":
        return out
    cur = MARK
    for ch in text:
        if ch == " " or ch == MARK:
            out.append(cur)
            cur = MARK
        elif ch in vocab:
            cur += ch
        else:
            cur += "".join("<0x%02X>" % b for b in ch.encode("utf-8"))
    out.append(cur)
    return out


def bytelevel_pieces(tok, text: str) -> List[str]:
    return [p[0] for p in tok._tokenizer.pre_tokenizer.pre_tokenize_str(text)]


def bytelevel_encode(tok, vocab_to_index: Dict[str, int], text: str) -> List[int]:
    ids: List[int] = []
    for piece in bytelevel_pieces(tok, text):
        units = [c for c in piece]
        for c in units:
            vocab_to_index[c]  # KeyError like tokenizer_utils.py:149
        sel, _, _ = dp_oracle.select_shortest(units, vocab_to_index, max_token_units=256)
        if sel is None:
            raise ValueError("max() arg is an empty sequence")
        ids.extend(vocab_to_index[t] for t in sel)
    return ids
