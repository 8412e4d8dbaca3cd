"""Seeded synthetic corpora shaped like the reference's inputs.

There is no network, so the S2ORC abstracts (main_analyze_s2orc.py:253-255), the
WMT biomedical sentence pairs (main_biomed_translation.py:71-73) and the MADAR
lexicon (dialect_arabic.py:24) are replaced by generators with the same surface
statistics: Zipfian word frequencies over a large lexicon, mean word length
~5.5 letters with a tail to 30, punctuation, digits, ~1 % accented letters,
German compounds with umlauts, Arabic-script 2-byte letters with diacritics.

Everything is numpy-vectorised: 100 MB of text takes a few seconds.
"""
from __future__ import annotations

import numpy as np

_ONSETS = ["", "b", "c", "d", "f", "g", "h", "j", "k", "l", "m", "n", "p", "r", "s", "t", "v", "w",
           "st", "tr", "pr", "ch", "sh", "th", "pl", "gr", "br", "cl", "sp", "qu", "fl", "cr", "ph"]
_VOWELS = ["a", "e", "i", "o", "u", "a", "e", "i", "o", "ea", "io", "ou", "ai", "ie", "y"]
_CODAS = ["", "", "", "n", "r", "s", "t", "l", "m", "d", "c", "nt", "st", "ng", "ns", "ct", "ss", "x"]
_SUFFIXES = ["", "", "", "", "s", "ed", "ing", "tion", "al", "ic", "ly", "ity", "ment", "ous", "ive",
             "ation", "ized", "ological", "ability", "ically"]
_ACCENTED = ["é", "è", "à", "ö", "ü", "ç", "ñ", "í", "ó", "ä"]
_DE_JOIN = ["", "", "s", "en", "er"]
_DE_EXTRA = ["ä", "ö", "ü", "ß", "sch", "ei", "ch", "ung", "keit", "lich"]
_AR_LETTERS = [chr(c) for c in range(0x0621, 0x064B)]
_AR_MARKS = [chr(c) for c in range(0x064B, 0x0653)]


def _syllable(rng, n):
    o = rng.integers(0, len(_ONSETS), n)
    v = rng.integers(0, len(_VOWELS), n)
    c = rng.integers(0, len(_CODAS), n)
    return o, v, c


def make_lexicon(n_types: int = 200_000, seed: int = 0, flavour: str = "en") -> list:
    """``n_types`` distinct lowercase word forms, short ones first (rank order)."""
    rng = np.random.default_rng(seed + 17)
    words = []
    seen = set()
    # frequent ranks get fewer syllables (law of abbreviation)
    target = 0
    while len(words) < n_types:
        target += 1
        batch = 4096
        rank_frac = len(words) / n_types
        rank = len(words)
        if rank < 64:
            max_syl = 1
        elif rank < 4096:
            max_syl = 2
        else:
            max_syl = 2 + int(4 * rank_frac ** 0.5)
        nsyl = rng.integers(1, max_syl + 1, batch)
        if rank < 4096:
            batch = 64
        for k in range(batch):
            parts = []
            for _ in range(int(nsyl[k])):
                parts.append(_ONSETS[rng.integers(0, len(_ONSETS))])
                parts.append(_VOWELS[rng.integers(0, len(_VOWELS))])
                parts.append(_CODAS[rng.integers(0, len(_CODAS))])
            if rank >= 4096:
                parts.append(_SUFFIXES[rng.integers(0, len(_SUFFIXES))])
            if flavour == "en" and rng.random() < 0.01:
                parts.insert(rng.integers(0, len(parts)), _ACCENTED[rng.integers(0, len(_ACCENTED))])
            if flavour == "de":
                if rng.random() < 0.25:
                    parts.insert(rng.integers(0, len(parts)), _DE_EXTRA[rng.integers(0, len(_DE_EXTRA))])
            w = "".join(parts)
            if flavour == "de" and rng.random() < 0.15 and len(words) > 100:
                # compound: glue two or three earlier words
                k2 = rng.integers(2, 4)
                pieces = [words[rng.integers(0, len(words))] for _ in range(k2)]
                w = _DE_JOIN[rng.integers(0, len(_DE_JOIN))].join(pieces)[:40]
            if not w or len(w) > 30 and flavour != "de" or w in seen:
                continue
            seen.add(w)
            words.append(w)
            if len(words) >= n_types:
                break
    return words


def make_arabic_lexicon(n_types: int = 100_000, seed: int = 0) -> list:
    rng = np.random.default_rng(seed + 29)
    # length distribution of explore.ipynb cell 8: mostly 3-6 letters, ~6 % >= 8
    lens = rng.choice([2, 3, 4, 5, 6, 7, 8, 9, 10, 12], n_types * 2,
                      p=[0.05, 0.2, 0.27, 0.22, 0.13, 0.07, 0.03, 0.015, 0.01, 0.005])
    out, seen = [], set()
    for L in lens:
        letters = [_AR_LETTERS[i] for i in rng.integers(0, len(_AR_LETTERS), int(L))]
        if rng.random() < 0.10:
            for _ in range(int(rng.integers(1, 3))):
                pos = int(rng.integers(1, len(letters) + 1))
                letters.insert(pos, _AR_MARKS[rng.integers(0, len(_AR_MARKS))])
        w = "".join(letters)
        if w in seen:
            continue
        seen.add(w)
        out.append(w)
        if len(out) >= n_types:
            break
    return out


def _zipf_cdf(n, s=1.0):
    p = 1.0 / np.arange(1, n + 1) ** s
    return np.cumsum(p / p.sum())


def _decorate(words_enc, rng, idx, sentence_len=(15, 30), capitalise=True):
    """Turn a stream of lexicon indices into word tokens with punctuation/digits."""
    n = len(idx)
    toks = [words_enc[i] for i in idx]
    # numbers / percentages / parentheses / hyphens
    r = rng.random(n)
    for k in np.nonzero(r < 0.03)[0]:
        v = rng.integers(0, 4)
        if v == 0:
            toks[k] = b"%d" % rng.integers(0, 3000)
        elif v == 1:
            toks[k] = b"%d.%d%%" % (rng.integers(0, 100), rng.integers(0, 10))
        elif v == 2:
            toks[k] = b"(" + toks[k] + b")"
        else:
            toks[k] = toks[k] + b"-" + toks[(k + 1) % n]
    # sentence structure
    pos = 0
    while pos < n:
        L = int(rng.integers(sentence_len[0], sentence_len[1] + 1))
        end = min(n, pos + L) - 1
        if capitalise:
            t = toks[pos]
            if t[:1].isalpha() and t[0] < 128:
                toks[pos] = t[:1].upper() + t[1:]
        for c in range(pos + 3, end, 7):
            if rng.random() < 0.5:
                toks[c] = toks[c] + (b"," if rng.random() < 0.8 else b";")
        toks[end] = toks[end] + b"."
        pos = end + 1
    return toks


def gen_documents(n_bytes: int, seed: int = 0, flavour: str = "en", lexicon=None,
                  words_per_doc=(150, 250), newline_headers: bool = False):
    """(text: np.uint8[N], doc_offs: np.int64[n_docs+1]) of about ``n_bytes`` bytes.

    Single spaces between words; documents are concatenated WITHOUT a separator
    (``doc_offs`` delimits them) so each document is exactly what a caller would
    pass to ``dp_tokenize`` (main_analyze_s2orc.py:78).
    """
    rng = np.random.default_rng(seed)
    if lexicon is None:
        lexicon = make_arabic_lexicon(seed=seed) if flavour == "ar" else make_lexicon(seed=seed, flavour=flavour)
    words_enc = [w.encode("utf-8") for w in lexicon]
    mean_len = float(np.mean([len(w) for w in words_enc[:2000]])) + 1.6
    cdf = _zipf_cdf(len(words_enc))
    chunks, offs, total = [], [0], 0
    while total < n_bytes:
        n_words = int(min(4_000_000, max(2000, (n_bytes - total) / mean_len * 1.02)))
        idx = np.searchsorted(cdf, rng.random(n_words))
        idx = np.minimum(idx, len(words_enc) - 1)
        toks = _decorate(words_enc, rng, idx, capitalise=(flavour != "ar"))
        pos = 0
        while pos < n_words and total < n_bytes:
            L = int(rng.integers(words_per_doc[0], words_per_doc[1] + 1))
            doc_toks = toks[pos:pos + L]
            pos += L
            doc = b" ".join(doc_toks)
            if newline_headers and rng.random() < 0.2:
                doc = b"PURPOSE\n" + doc + b"\n\n\nRESULTS\n" + b" ".join(doc_toks[:20])
            chunks.append(doc)
            total += len(doc)
            offs.append(total)
    text = np.frombuffer(b"".join(chunks), dtype=np.uint8)
    return text, np.asarray(offs, dtype=np.int64)


def gen_sentence_pairs(n_bytes: int, seed: int = 0):
    """en/de sentence pairs, one pair per document, 'en text\\nde text' (config C2)."""
    rng = np.random.default_rng(seed + 5)
    en = [w.encode() for w in make_lexicon(120_000, seed, "en")]
    de = [w.encode() for w in make_lexicon(120_000, seed + 1, "de")]
    cdf_en, cdf_de = _zipf_cdf(len(en)), _zipf_cdf(len(de))
    chunks, offs, total = [], [0], 0
    while total < n_bytes:
        nw = 400_000
        ie = np.minimum(np.searchsorted(cdf_en, rng.random(nw)), len(en) - 1)
        idd = np.minimum(np.searchsorted(cdf_de, rng.random(nw)), len(de) - 1)
        te = _decorate(en, rng, ie, sentence_len=(12, 40))
        td = _decorate(de, rng, idd, sentence_len=(10, 35))
        pe = pd = 0
        while pe < nw - 50 and pd < nw - 50 and total < n_bytes:
            Le, Ld = int(rng.integers(12, 40)), int(rng.integers(10, 35))
            s_en = b" ".join(te[pe:pe + Le])
            s_de = b" ".join(td[pd:pd + Ld])
            if rng.random() < 0.1:
                s_de = "„".encode() + s_de + "“".encode()
            pe += Le
            pd += Ld
            for doc in (s_en, s_de):
                chunks.append(doc)
                total += len(doc)
                offs.append(total)
    text = np.frombuffer(b"".join(chunks), dtype=np.uint8)
    return text, np.asarray(offs, dtype=np.int64)


def sample_text(n_bytes: int, seed: int = 0, flavour: str = "en") -> list:
    """List of document strings (for tokenizer training and small tests)."""
    text, offs = gen_documents(n_bytes, seed, flavour)
    raw = text.tobytes()
    return [raw[offs[k]:offs[k + 1]].decode("utf-8") for k in range(len(offs) - 1)]
