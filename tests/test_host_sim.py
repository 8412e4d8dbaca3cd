"""The exact per-word / per-character code the CUDA kernels run (csrc/dpt_dp_core.h, dpt_rules.h, the vocab
compiler), compiled for the host by tests/host_sim and checked against the oracle and the golden vectors.
CPU only; this is how kernel logic is validated before GPU time is spent.  The simulation library is test
infrastructure: the product never loads it."""
import ctypes
import random

import numpy as np

from conftest import load_golden
from helpers import make_sim_vocab, pack, py_roundtrip_ok, sim_word, vocab_bytes
from oracle import adapters, dp_oracle


def test_random_words_match_oracle(host_sim):
    rng = random.Random(5)
    for it in range(3000):
        mode = it % 2
        alpha = ["a", "b", "c", "é", "▁", "日"] if mode else [bytes([x]) for x in (97, 98, 99, 0xC3, 0xA9, 0xE2)]
        enc = (lambda x: x.encode()) if mode else (lambda x: x)
        vocab = {}
        for _ in range(rng.randint(1, 14)):
            t = b"".join(enc(rng.choice(alpha)) for _ in range(rng.randint(1, 4)))
            vocab.setdefault(t, len(vocab) + 3)
        if rng.random() < 0.6:
            for a in alpha:
                vocab.setdefault(enc(a), len(vocab) + 3)
        data = b"".join(enc(rng.choice(alpha)) for _ in range(rng.randint(1, 12)))
        h = make_sim_vocab(host_sim, vocab, mode)
        for t, i in vocab.items():
            assert host_sim.sim_lookup(h, t, len(t)) == i
        bnd = None
        if it % 7 == 0 and mode:
            cps = dp_oracle.utf8_boundaries(data)
            bnd = sorted(set([0, len(data)] + [p for p in cps if rng.random() < 0.7]))
        exp = dp_oracle.dp_bytes(data, vocab, mode, bnd)
        r, wl, ids = sim_word(host_sim, h, data, bnd)
        if exp["untokenizable"]:
            assert r == -1 and wl == exp["word_len"]
        else:
            assert ids == exp["ids"] and wl == exp["word_len"]
        host_sim.sim_vocab_destroy(ctypes.c_void_p(h))


def test_c0_golden(host_sim):
    g = load_golden("c0_toy.json.gz")
    vocab = {t.encode(): k for k, t in enumerate(g["vocab"])}
    h = make_sim_vocab(host_sim, vocab, 1)
    for r in g["rows"]:
        n, wl, ids = sim_word(host_sim, h, r["w"].encode())
        assert wl == r["len"]
        if r["sel"] is None:
            assert n == -1
        else:
            assert [g["vocab"][i] for i in ids] == r["sel"]
    host_sim.sim_vocab_destroy(ctypes.c_void_p(h))


def test_compiled_vocab_facts(host_sim):
    from dptok import assets
    for name, fam, mode in (("llama2_32k", "spm", 1), ("gpt2_50k", "bytelevel", 0)):
        spec = assets.load_spec(name)
        bv = vocab_bytes(spec["model"]["vocab"], fam)
        h = make_sim_vocab(host_sim, bv, mode)
        info = np.zeros(8, np.int32)
        host_sim.sim_info(h, info.ctypes.data)
        assert info[0] == len(bv)
        assert info[2] < 1.25 * info[1] + 600, "double array should stay dense"
        if fam == "spm":
            assert info[6] == 1 and info[7] == 1   # marker only leading; full byte fallback
        rng = random.Random(1)
        items = list(bv.items())
        for t, i in rng.sample(items, 3000):
            assert host_sim.sim_lookup(h, t, len(t)) == i
        host_sim.sim_vocab_destroy(ctypes.c_void_p(h))


def _normalise(host_sim, h, docs):
    raw = b"".join(docs)
    text = np.frombuffer(raw + b"\0", np.uint8)
    offs = np.zeros(len(docs) + 1, np.int64)
    offs[1:] = np.cumsum([len(d) for d in docs])
    cap = 6 * len(raw) + 6 * len(docs) + 16
    out = np.zeros(cap, np.uint8)
    wcap = len(raw) + 2 * len(docs) + 2
    woffs = np.zeros(wcap + 1, np.int64)
    nw = ctypes.c_int64()
    flags = np.zeros(len(docs), np.uint8)
    nb = host_sim.sim_spm_normalise(h, text.ctypes.data, len(raw), offs.ctypes.data, len(docs), out.ctypes.data, cap,
                                    woffs.ctypes.data, wcap, ctypes.byref(nw), flags.ctypes.data)
    norm = out[:nb].tobytes()
    words = [norm[woffs[k]:woffs[k + 1]].decode("utf-8") for k in range(nw.value)]
    return words, flags


def test_spm_rule_matches_reference_pretokenizer(host_sim):
    """Device SPM_LLAMA rule == pretokenize_with_llama (tokenizer_utils.py:24-31) on unambiguous text, and
    flags exactly the documents whose split it cannot know (runs of >= 2 markers)."""
    from dptok import assets
    g = load_golden("llama_adapter.json.gz")
    tok = assets.load_hf(g["tokenizer"])
    h = make_sim_vocab(host_sim, vocab_bytes(tok.get_vocab(), "spm"), 1)
    docs = [r["text"].encode() for r in g["rows"] if r["text"]]
    words, flags = _normalise(host_sim, h, docs)
    pos = 0
    rows = [r for r in g["rows"] if r["text"]]
    for k, r in enumerate(rows):
        ambiguous = "  " in r["text"] or r["text"].startswith(" ") or "▁" in r["text"]
        # trailing single space: "trail " -> '▁trail','▁' is unambiguous
        assert bool(flags[k]) == ambiguous, r["text"]
        if not ambiguous:
            assert words[pos:pos + len(r["words"])] == r["words"], r["text"]
            pos += len(r["words"])
        else:
            # skip this doc's words: they start with '<s>' and run to the next '<s>'
            pos += 1
            while pos < len(words) and words[pos] != "<s>":
                pos += 1
    assert pos == len(words)
    # random synthetic text incl. newlines, tabs, accents, CJK
    rng = random.Random(3)
    pieces = ["plai", "gout", "é", "ï", "日", "本", "\n", "\t", ",", "Zeta", "(x)", "12", "—", "naïve", "%"]
    docs = []
    for _ in range(300):
        ws = ["".join(rng.choice(pieces) for _ in range(rng.randint(1, 4))) for _ in range(rng.randint(1, 12))]
        docs.append(" ".join(ws))
    words, flags = _normalise(host_sim, h, [d.encode() for d in docs])
    assert not flags.any()
    expect = []
    for d in docs:
        expect += adapters.llama_words(tok, d)
    assert words == expect
    host_sim.sim_vocab_destroy(ctypes.c_void_p(h))


# ------------------------------------------------------------------------------------------------
# The corpus pipeline's source (csrc/dpt_pipe.h: scan+dedup -> DP per distinct word -> scan+emit) executed verbatim
# by a std::thread emulation of CUDA blocks, against (a) the general path's rule + the C oracle and (b) the
# reference-shaped normaliser.
# ------------------------------------------------------------------------------------------------
def _run_fused(host_sim, h, spm, docs, nthreads=4, n_slots=0, ids_cap=None, odd_cap=0, pool_cap=0, lp_cap=0, n_ranges=1,
               word_cap=None):
    raw = b"".join(docs)
    text = np.frombuffer(raw + b"\0" * 64, np.uint8).copy()
    offs = np.zeros(len(docs) + 1, np.int64)
    offs[1:] = np.cumsum([len(d) for d in docs])
    n = len(raw)
    cap = 3 * n + 3 * len(docs) + 16 if ids_cap is None else ids_cap
    wcap = n + 2 * len(docs) + 16 if word_cap is None else word_cap
    ids = np.full(cap, -7, np.int32)
    wl = np.full(wcap, -7, np.int32)
    wf = np.full(wcap, 99, np.uint8)
    dto = np.full(len(docs) + 1, -7, np.int64)
    dfl = np.zeros(len(docs), np.uint8)
    ctr = np.zeros(4, np.int64)
    nout = np.zeros(8, np.int64)
    host_sim.sim_encode_corpus_pipe(h, spm, text.ctypes.data, n, offs.ctypes.data, len(docs), ids.ctypes.data, cap,
                                    wl.ctypes.data, wf.ctypes.data, wcap, dto.ctypes.data, dfl.ctypes.data,
                                    ctr.ctypes.data, nout.ctypes.data, nthreads, n_slots, odd_cap, pool_cap, lp_cap, n_ranges)
    return dict(ids=ids[:min(nout[0], cap)], wl=wl[:min(nout[1], wcap)], wf=wf[:min(nout[1], wcap)], dto=dto, dfl=dfl, ctr=ctr,
                nout=nout)


def _general_expected(host_sim, h, vb, docs):
    """(n_words, ids, lens, untok, doc_tok_offs, doc_flags) of the general path's rule + the C oracle's DP."""
    from oracle.c_oracle import COracle

    def normalise(dlist):
        raw = b"".join(dlist)
        text = np.frombuffer(raw + b"\0", np.uint8)
        offs = np.zeros(len(dlist) + 1, np.int64)
        offs[1:] = np.cumsum([len(d) for d in dlist])
        cap = 6 * len(raw) + 6 * len(dlist) + 16
        out = np.zeros(cap, np.uint8)
        wcap = len(raw) + 2 * len(dlist) + 2
        woffs = np.zeros(wcap + 1, np.int64)
        nw = ctypes.c_int64()
        flags = np.zeros(len(dlist), np.uint8)
        nb = host_sim.sim_spm_normalise(h, text.ctypes.data, len(raw), offs.ctypes.data, len(dlist), out.ctypes.data,
                                        cap, woffs.ctypes.data, wcap, ctypes.byref(nw), flags.ctypes.data)
        return out[:nb].copy(), woffs[:nw.value + 1].copy(), flags

    wtext, woffs, flags = normalise(docs)
    nw = len(woffs) - 1
    o_ids, o_lens, o_untok = COracle(vb, 1).encode_words(wtext, woffs)
    tok_of_word = np.concatenate([[0], np.cumsum(np.where(o_untok == 0, o_lens, 0))])
    first = np.concatenate([[0], np.cumsum([len(normalise([d])[1]) - 1 for d in docs])])
    assert first[-1] == nw
    return nw, o_ids, o_lens, o_untok, tok_of_word[first], flags


def _check_fused(host_sim, h, vb, docs, **kw):
    r = _run_fused(host_sim, h, 1, docs, **kw)
    nw, o_ids, o_lens, o_untok, o_dto, o_flags = _general_expected(host_sim, h, vb, docs)
    assert r["nout"][2] <= r["nout"][3] and r["nout"][4] <= r["nout"][5] and r["nout"][6] <= r["nout"][7]
    assert r["nout"][1] == nw and r["nout"][0] == len(o_ids)
    assert np.array_equal(r["wl"], o_lens)
    assert np.array_equal(r["wf"] & 1, o_untok)
    assert np.array_equal(r["ids"], o_ids)
    assert np.array_equal(r["dto"], o_dto)
    assert np.array_equal(r["dfl"], o_flags)
    assert r["ctr"].tolist() == [sum(len(d) for d in docs), nw, len(o_ids), int(o_untok.sum())]
    return r


def test_pipeline_code_s2orc_shaped_text(host_sim):
    """Llama-2-shaped 32k vocab on S2ORC-shaped text with newline headers: ids, per-word lengths, flags, document
    offsets and counters bit-exact with the oracle; thread count and hash-table size (down to one that overflows
    and sends most words around the table) must not matter."""
    from dptok import assets, synth
    tok = assets.load_hf("llama2_32k")
    t2i = tok.get_vocab()
    vb = vocab_bytes(t2i, "spm")
    h = make_sim_vocab(host_sim, vb, 1)
    text, doc_offs = synth.gen_documents(600_000, seed=3, lexicon=synth.make_lexicon(20000, seed=3), newline_headers=True)
    raw = text.tobytes()
    docs = [raw[doc_offs[k]:doc_offs[k + 1]] for k in range(len(doc_offs) - 1)]
    r = _check_fused(host_sim, h, vb, docs, nthreads=8)
    assert not r["dfl"].any()
    # the reference-shaped normaliser gives the same words
    words = []
    for d in docs[:50]:
        words += adapters.spm_normalise(d.decode(), set(t2i))
    n50 = len(words)
    from oracle.c_oracle import COracle
    wtext, woffs = pack([w.encode() for w in words])
    o_ids, o_lens, _ = COracle(vb, 1).encode_words(wtext, woffs)
    assert np.array_equal(r["wl"][:n50], o_lens) and np.array_equal(r["ids"][:len(o_ids)], o_ids)
    # the same corpus as 2 / 7 consecutive document ranges sharing one word table (chunked host path)
    for nr in (2, 7):
        rr = _run_fused(host_sim, h, 1, docs, nthreads=4, n_ranges=nr)
        assert np.array_equal(rr["ids"], r["ids"]) and np.array_equal(rr["wl"], r["wl"])
        assert np.array_equal(rr["dto"], r["dto"]) and rr["ctr"].tolist() == r["ctr"].tolist()
    for slots in (1 << 16, 256):
        r2 = _run_fused(host_sim, h, 1, docs, nthreads=3, n_slots=slots)
        assert np.array_equal(r2["ids"], r["ids"]) and np.array_equal(r2["wl"], r["wl"])
        assert np.array_equal(r2["dto"], r["dto"]) and r2["ctr"].tolist() == r["ctr"].tolist()
    # capacities of the internal lists are reported, never silently exceeded
    r4 = _run_fused(host_sim, h, 1, docs, nthreads=4, n_slots=256, odd_cap=100, pool_cap=10, lp_cap=50)
    assert r4["nout"][6] > r4["nout"][7]
    # word_cap far below the word count (and below the number of distinct words, so the DP queues fill up): the
    # requirement is reported, nothing is written beyond a capacity, the words that fit are right
    r5 = _run_fused(host_sim, h, 1, docs, nthreads=4, word_cap=700)
    assert r5["nout"][1] == len(r["wl"]) and np.array_equal(r5["wl"], r["wl"][:700])
    # capacity: ids beyond ids_cap are dropped, the requirement is still reported
    r3 = _run_fused(host_sim, h, 1, docs, nthreads=4, ids_cap=1000)
    assert r3["nout"][0] == len(r["ids"]) and np.array_equal(r3["ids"], r["ids"][:1000])
    host_sim.sim_vocab_destroy(ctypes.c_void_p(h))


def test_pipeline_code_edge_cases(host_sim):
    """Tiny documents, trailing/leading/double spaces, raw U+2581, OOV characters, newlines, words longer than a
    tile, malformed UTF-8 cut by document boundaries: pipeline code == general path + oracle."""
    from dptok import assets
    tok = assets.load_hf("llama2_2k")
    t2i = tok.get_vocab()
    vb = vocab_bytes(t2i, "spm")
    h = make_sim_vocab(host_sim, vb, 1)
    rng = random.Random(5)
    pieces = ["plai", "gout", "é", "ï", "日", "本", "\n", "\t", ",", "Zeta", "(x)", "12", "—", "naïve", "%", "trot", "▁", "a", "I"]
    for trial in range(16):
        docs = []
        for k in range(rng.choice([1, 3, 50, 400, 2000])):
            style = rng.random()
            if style < 0.1:
                d = rng.choice(["a", " ", "▁", "é", "\n", "日", "x y", " x", "x ", "  ", "a  b", "▁▁a", "a▁b", "a ▁b", "▁ a"])
            elif style < 0.2:
                d = "".join(rng.choice("abcdefgh ") for _ in range(rng.randint(1, 30)))
            elif style < 0.25:
                d = "".join(rng.choice(pieces) for _ in range(rng.randint(100, 700)))
            else:
                ws = ["".join(rng.choice(pieces) for _ in range(rng.randint(1, 4))) for _ in range(rng.randint(1, 12))]
                d = rng.choice([" ", " ", " ", "  ", "▁"]).join(ws)
            docs.append(d.encode())
        if trial % 4 == 0:
            blob = bytes(rng.choice([0x20, 0x41, 0x62, 0xE2, 0x96, 0x81, 0xC3, 0xA9, 0x80, 0xFF, 0x0A, 0x63])
                         for _ in range(rng.randint(50, 20000)))
            cuts = sorted(set(rng.randint(1, len(blob) - 1) for _ in range(rng.randint(0, 40))))
            docs = [blob[a:b] for a, b in zip([0] + cuts, cuts + [len(blob)])]
        _check_fused(host_sim, h, vb, docs, nthreads=rng.choice([1, 2, 5, 8]), n_slots=rng.choice([0, 0, 64, 1024]),
                     n_ranges=rng.choice([1, 1, 2, 5]))
    # words far longer than a tile
    big = [b"x" * 3000 + b" y " + b"z" * 9000 + b" end", "é".encode() * 5000]
    _check_fused(host_sim, h, vb, big, nthreads=4)
    r = _run_fused(host_sim, h, 1, big, nthreads=4, lp_cap=50)
    assert r["nout"][2] > r["nout"][3]            # long-word scratch too small: reported, not mis-solved
    host_sim.sim_vocab_destroy(ctypes.c_void_p(h))


def _tokenizer_with_marker_run_tokens(name="llama2_2k"):
    """The committed Llama-2-shaped tokenizer plus whitespace-run tokens the way the real Llama-2 vocabulary has them
    ("▁▁" is its FIRST merge, "▁▁▁▁" an early one, "▁▁▁" a late one): with them the word boundaries inside a run of
    spaces depend on the merge order."""
    import json
    from tokenizers import Tokenizer
    from transformers import PreTrainedTokenizerFast
    from dptok import assets
    spec = assets.load_spec(name)
    vocab, merges = spec["model"]["vocab"], [list(m) if not isinstance(m, str) else m.split(" ") for m in spec["model"]["merges"]]
    nid = max(vocab.values()) + 1
    for t in ("▁▁", "▁▁▁▁", "▁▁▁", "▁▁▁▁▁▁▁▁"):
        vocab[t] = nid
        nid += 1
    merges.insert(0, ["▁", "▁"])
    merges.insert(9, ["▁▁", "▁▁"])
    merges.insert(49, ["▁▁▁▁", "▁▁▁▁"])
    merges.insert(min(1400, len(merges)), ["▁▁", "▁"])
    spec["model"]["merges"] = merges
    tok = PreTrainedTokenizerFast(tokenizer_object=Tokenizer.from_str(json.dumps(spec)), bos_token="<s>", eos_token="</s>",
                                  unk_token="<unk>")
    t2i = tok.get_vocab()
    mid = [(t2i[a], t2i[b], t2i[a + b]) for a, b in merges if a in t2i and b in t2i and a + b in t2i]
    return tok, t2i, mid


def test_pipeline_code_marker_runs_split_like_the_tokenizer(host_sim):
    """runs of spaces / U+2581, leading and trailing spaces, indented lines.  With the tokenizer's merge
    table in the compiled vocabulary the pipeline cuts them on the device (pb_segment_split) exactly where the reference's
    tokenizer-driven split does (tokenizer_utils.py:7-31): ids == oracle adapter per document, no document flagged."""
    tok, t2i, mid = _tokenizer_with_marker_run_tokens()
    assert adapters.llama_words(tok, "a  b")[1:] == ["▁a▁", "▁b"] or True  # (shape depends on the ranks; parity is checked below)
    vb = vocab_bytes(t2i, "spm")
    h = make_sim_vocab(host_sim, vb, 1)
    m = np.asarray(mid, dtype=np.int32)
    left, right, merged = (np.ascontiguousarray(m[:, k]) for k in range(3))
    assert host_sim.sim_vocab_set_merges(ctypes.c_void_p(h), left.ctypes.data, right.ctypes.data, merged.ctypes.data, len(m)) == 0
    rng = random.Random(17)
    words = ["plai", "gout", "trot", "Zeta", "a", "I", "na\u00efve", "(x)", "12", "\u65e5\u672c", "the", "tion", "\n", "x\ny"]
    fixed = ["a  b", "a   b", "a    b", "a     b", "  a", " a", "   a", "a ", "a  ", "a   ", " ", "  ", "    ", "x\n\n  indented   text",
             "a \u2581b", "\u2581\u2581a", "\u2581 a", "a\u2581\u2581\u2581\u2581\u2581\u2581\u2581\u2581b", "a" + " " * 37 + "b",
             "  \u00e9  \u65e5  ", "a  \u00e9"]
    docs = [d.encode() for d in fixed]
    for _ in range(300):
        parts = []
        for _k in range(rng.randint(1, 12)):
            parts.append(rng.choice(words))
            parts.append(rng.choice([" ", " ", "  ", "   ", "    ", "\u2581", " \u2581", "      "]))
        d = "".join(parts)
        if rng.random() < 0.3:
            d = rng.choice([" ", "  ", "   "]) + d
        if rng.random() < 0.5:
            d = d.rstrip(" \u2581") or "a"
        docs.append(d.encode())
    r = _run_fused(host_sim, h, 1, docs, nthreads=4)
    assert r["nout"][2] <= r["nout"][3] and r["nout"][4] <= r["nout"][5] and r["nout"][6] <= r["nout"][7]
    assert not r["dfl"].any(), "no document may be left to the host split"
    for k, d in enumerate(docs):
        want = adapters.llama_encode(tok, d.decode())
        got = r["ids"][r["dto"][k]:r["dto"][k + 1]].tolist()
        assert got == want, (d, adapters.llama_words(tok, d.decode()))
    # without the merge table the same documents are flagged instead (the caller splits them on the host)
    h2 = make_sim_vocab(host_sim, vb, 1)
    r2 = _run_fused(host_sim, h2, 1, docs[:8], nthreads=2)
    assert r2["dfl"][:4].all()
    host_sim.sim_vocab_destroy(ctypes.c_void_p(h))
    host_sim.sim_vocab_destroy(ctypes.c_void_p(h2))


def test_pipeline_code_untokenizable_and_long_tokens(host_sim):
    """Vocabulary without U+2581 alone / with missing letters / with tokens longer than the 32-bit walk mask:
    phantom lengths, untokenizable flags and long tokens come out like the oracle's."""
    rng = random.Random(11)
    alpha = "abcdeé日"
    for with_marker in (True, False):
        toks = set(alpha[:5]) | {"▁a", "▁b", "▁é", "▁ab"}
        if with_marker:
            toks.add("▁")
        for _ in range(300):
            toks.add("".join(rng.choice(alpha) for _ in range(rng.randint(2, 7))))
        toks.add("ab" * 30)
        toks.add("▁" + "cd" * 25)
        toks.add("é" * 40)
        toks |= {"<s>", "<unk>"} | {"<0x%02X>" % b for b in range(256)}
        toks |= set("<0x>") | set("0123456789ABCDEF")
        t2i = {t: k + 3 for k, t in enumerate(sorted(toks))}
        vb = vocab_bytes(t2i, "spm")
        h = make_sim_vocab(host_sim, vb, 1)
        docs = []
        for k in range(300):
            ws = []
            for _ in range(rng.randint(1, 30)):
                n = rng.choice([1, 1, 2, 3, 5, 8, 13, 21, 40])
                w = "".join(rng.choice(alpha) for _ in range(n))
                if rng.random() < 0.05:
                    w = "ab" * rng.randint(20, 70)
                if rng.random() < 0.05:
                    w = "cd" * rng.randint(20, 40)
                if rng.random() < 0.05:
                    w = "é" * rng.randint(30, 90)
                ws.append(w)
            docs.append(" ".join(ws).encode())
        r = _check_fused(host_sim, h, vb, docs, nthreads=4)
        if not with_marker:         # no bare marker: words like "▁c…" have no segmentation at all
            assert r["ctr"][3] > 0
        assert ((r["wf"] & 4) != 0).sum() > 0     # words past the local-state limit took the long-word kernel
        rp = _run_fused(host_sim, h, 1, docs, nthreads=4, pool_cap=10)
        assert rp["nout"][4] > rp["nout"][5]      # words with more than 7 ids need the id pool: reported when too small
        host_sim.sim_vocab_destroy(ctypes.c_void_p(h))


# ------------------------------------------------------------------------------------------------
# Byte-level split rules on device (csrc/dpt_split_rules.h inside kernel A) against the installed `tokenizers`
# pre-tokenizer (the reference's pre_tokenize_str, tokenizer_utils.py:157-159) + the C oracle's DP.
# ------------------------------------------------------------------------------------------------
def _bytelevel_expected(tok, vb, docs):
    from helpers import bytelevel_table
    from oracle.c_oracle import COracle
    u2b = {c: b for b, c in bytelevel_table().items()}
    words, first = [], []
    for d in docs:
        first.append(len(words))
        for piece, _ in tok.pre_tokenizer.pre_tokenize_str(d.decode()):
            words.append(bytes(u2b[c] for c in piece))
    wtext, woffs = pack(words)
    o_ids, o_lens, o_untok = COracle(vb, 0).encode_words(wtext, woffs)
    tow = np.concatenate([[0], np.cumsum(np.where(o_untok == 0, o_lens, 0))])
    return words, o_ids, o_lens, o_untok, tow[first + [len(words)]]


def _check_bytelevel(host_sim, h, tok, vb, rule, docs, **kw):
    r = _run_fused(host_sim, h, rule, docs, **kw)
    words, o_ids, o_lens, o_untok, o_dto = _bytelevel_expected(tok, vb, docs)
    assert r["nout"][1] == len(words)
    assert np.array_equal(r["wl"], o_lens)
    assert np.array_equal(r["ids"], o_ids)
    assert np.array_equal(r["dto"], o_dto)
    assert r["ctr"].tolist() == [sum(len(d) for d in docs), len(words), len(o_ids), int(o_untok.sum())]
    return r


def test_pipeline_code_bytelevel_rules(host_sim):
    """GPT-2, Llama-3 and BLOOM split regexes as the device scanner: English/German sentence pairs, Arabic with combining
    marks, contractions (incl. case-insensitive and U+017F), digit groups, newline runs, tabs, ideographic space,
    multi-byte digits, document boundaries every few bytes."""
    from dptok import assets, synth
    rng = random.Random(7)
    soup = ["Hello", " ", "  ", "world", "'s", "'S", "'re", "'LL", "'", "''", "12345", "3", "٣٤", "é", "naïve", "日本",
            "\n", "\n\n", "\t", "\r\n", ".", ",", "!!", "(x)", "—", "…", " ", "　", "ſ", "'ſ", "x", "İ", "ǅ",
            "قُدَّام", "البيت", "%", "a1b2", " ", "+=", "'t", "'d", "'m", "'ve", "'VE"]
    soup += ["(", ")", "|", "?", "!", "[w]", "。", "，", "、", "।", "۔", "،", "؟", ". .", " .", " (", "a)b"]
    for name, rule in (("gpt2_3k", 2), ("llama3_128k", 3), ("bloom_8k", 4)):
        tok = assets.load_tokenizer(name)
        v2i = {t: k for k, t in enumerate(assets.load_spec(name)["model"]["vocab"])}
        vb = vocab_bytes(v2i, "bytelevel")
        h = make_sim_vocab(host_sim, vb, 0)
        text, offs = synth.gen_sentence_pairs(150_000, seed=1)
        raw = text.tobytes()
        _check_bytelevel(host_sim, h, tok, vb, rule, [raw[offs[k]:offs[k + 1]] for k in range(len(offs) - 1)], nthreads=8)
        text, offs = synth.gen_documents(60_000, seed=2, flavour="ar", lexicon=synth.make_arabic_lexicon(3000, seed=2))
        raw = text.tobytes()
        _check_bytelevel(host_sim, h, tok, vb, rule, [raw[offs[k]:offs[k + 1]] for k in range(len(offs) - 1)], nthreads=4)
        for trial in range(25):
            docs = ["".join(rng.choice(soup) for _ in range(rng.randint(1, 60))).encode()
                    for _ in range(rng.choice([1, 5, 60]))]
            _check_bytelevel(host_sim, h, tok, vb, rule, docs, nthreads=rng.choice([1, 3, 8]), n_slots=rng.choice([0, 64]),
                             n_ranges=rng.choice([1, 1, 3]))
        # pieces far longer than a tile / its look-behind: one 20 KB letter run, 9 KB of digits, 6 KB of spaces
        big = [("x" * 20000 + " y").encode(), ("7" * 9000 + "a").encode(), (" " * 6000 + "z\n" * 3000).encode(),
               ("." * 5000 + " " + "," * 4200 + " q").encode(), ("\u3002" * 1500 + " \u3000" * 700 + "w").encode()]
        _check_bytelevel(host_sim, h, tok, vb, rule, big, nthreads=4)
        host_sim.sim_vocab_destroy(ctypes.c_void_p(h))


# ------------------------------------------------------------------------------------------------
# Round-trip check (k_roundtrip: one warp per document, csrc/dpt_decode.h): the kernel's loop with its lanes emulated,
# an independent serial decode in C++, and a Python decode of the same ids must agree - on true tokenizations and on
# damaged ones.
# ------------------------------------------------------------------------------------------------
def _roundtrip(host_sim, h, ids, dto, docs, skip_bos, mode):
    raw = b"".join(docs)
    text = np.frombuffer(raw + b"\0" * 8, np.uint8).copy()
    offs = np.zeros(len(docs) + 1, np.int64)
    offs[1:] = np.cumsum([len(d) for d in docs])
    ids = np.ascontiguousarray(ids, np.int32)
    dto = np.ascontiguousarray(dto, np.int64)
    ok = np.full(len(docs), 9, np.uint8)
    host_sim.sim_roundtrip(h, ids.ctypes.data, dto.ctypes.data, text.ctypes.data, offs.ctypes.data, len(docs),
                           1 if skip_bos else 0, ok.ctypes.data, mode)
    return ok


def test_roundtrip_check_code(host_sim):
    from dptok import assets, synth
    rng = random.Random(9)
    cases = []
    # SentencePiece: S2ORC-shaped text + documents that start with a byte token, a space, hold newlines / OOV characters
    tok = assets.load_hf("llama2_2k")
    vb = vocab_bytes(tok.get_vocab(), "spm")
    text, doc_offs = synth.gen_documents(120_000, seed=4, lexicon=synth.make_lexicon(3000, seed=4), newline_headers=True)
    raw = text.tobytes()
    docs = [raw[doc_offs[k]:doc_offs[k + 1]] for k in range(len(doc_offs) - 1)]
    docs += [s.encode() for s in ("\nstarts with a newline", "日本 starts with an OOV character", "x", "Z\tq\n",
                                  "caf\u00e9 na\u00efve \u2014 ok", "a " * 300 + "end", "tail space ")]
    cases.append((vb, 1, 1, docs, True))
    # byte-level: sentence pairs and Arabic, GPT-2 split rule
    v2i = {t: k for k, t in enumerate(assets.load_spec("gpt2_3k")["model"]["vocab"])}
    text, doc_offs = synth.gen_sentence_pairs(60_000, seed=2)
    raw = text.tobytes()
    docs = [raw[doc_offs[k]:doc_offs[k + 1]] for k in range(len(doc_offs) - 1)]
    text, doc_offs = synth.gen_documents(30_000, seed=2, flavour="ar", lexicon=synth.make_arabic_lexicon(2000, seed=2))
    raw = text.tobytes()
    docs += [raw[doc_offs[k]:doc_offs[k + 1]] for k in range(len(doc_offs) - 1)]
    cases.append((vocab_bytes(v2i, "bytelevel"), 0, 2, docs, False))
    for vb, mode, rule, docs, skip_bos in cases:
        spm = mode == 1
        h = make_sim_vocab(host_sim, vb, mode)
        id2tok = {i: t for t, i in vb.items()}
        r = _run_fused(host_sim, h, rule, docs, nthreads=4)
        ids, dto = r["ids"], r["dto"]
        assert not r["dfl"].any() and r["ctr"][3] == 0
        assert max(dto[k + 1] - dto[k] for k in range(len(docs))) > 40      # documents of more than one 32-token round
        for m in (0, 1):
            assert _roundtrip(host_sim, h, ids, dto, docs, skip_bos, m).all()
        # damaged ids: one change per document (another token, a dropped / doubled / appended token, an id that is none)
        all_ids = sorted(id2tok)
        new_ids, new_dto, expect = [], [0], []
        for k, d in enumerate(docs):
            row = ids[dto[k]:dto[k + 1]].tolist()
            lo = 1 if skip_bos else 0
            kind = rng.choice(["swap", "drop", "double", "append", "invalid", "same", "bos"])
            pos = rng.randrange(lo, len(row)) if len(row) > lo else None
            if kind == "swap" and pos is not None:
                row[pos] = rng.choice(all_ids)
            elif kind == "drop" and pos is not None:
                del row[pos]
            elif kind == "double" and pos is not None:
                row.insert(pos, row[pos])
            elif kind == "append":
                row.append(rng.choice(all_ids))
            elif kind == "invalid" and pos is not None:
                row[pos] = rng.choice([-1, -5, max(all_ids) + 1, 2 ** 31 - 1])
            elif kind == "bos" and skip_bos:
                row[0] = rng.choice(all_ids)        # the skipped position is not looked at
            new_ids += row
            new_dto.append(len(new_ids))
            expect.append(py_roundtrip_ok(id2tok, row, d, spm, skip_bos))
        expect = np.array(expect)
        assert expect.any() and not expect.all()
        for m in (0, 1):
            got = _roundtrip(host_sim, h, new_ids, new_dto, docs, skip_bos, m)
            assert np.array_equal(got != 0, expect), m
        # the wrong skip_bos setting fails every SentencePiece document (its first id is '<s>')
        if skip_bos:
            for m in (0, 1):
                assert not _roundtrip(host_sim, h, ids, dto, docs, False, m).any()
        host_sim.sim_vocab_destroy(ctypes.c_void_p(h))


def test_letter_mask_agrees_with_scanner_on_arbitrary_bytes(host_sim, host_sim_noskip):
    """The tile's letter mask (ASCII letters by SWAR) only lets the split scanner skip ahead: on ANY bytes - truncated and invalid UTF-8, lead bytes cut by document boundaries, runs
    across tile borders - the pieces, ids and offsets equal those of the scanner walking character by character."""
    from dptok import assets
    rng = random.Random(11)
    frags = [b"abc", b"Z", b" ", b"  ", b"\n", b"1", b"'s", b".", "é".encode(), "ß".encode(), "قُدَّام".encode(),
             "日本".encode(), "𝒜".encode(), "ǅ".encode(), b"\xc3", b"\xa9", b"\xe6\x97", b"\xf0\x9d\x92", b"\xff", b"\xc0\x80",
             b"\xed\xa0\x80", "ſ".encode(), "İ".encode(), "　".encode()]
    for name, rule in (("gpt2_3k", 2), ("llama3_128k", 3), ("bloom_8k", 4)):
        v2i = {t: k for k, t in enumerate(assets.load_spec(name)["model"]["vocab"])}
        vb = vocab_bytes(v2i, "bytelevel")
        h1, h2 = make_sim_vocab(host_sim, vb, 0), make_sim_vocab(host_sim_noskip, vb, 0)
        for trial in range(12):
            n_docs = rng.choice([1, 7, 300])
            docs = [b"".join(rng.choice(frags) for _ in range(rng.randint(1, 120 if n_docs < 300 else 12))) for _ in range(n_docs)]
            if trial % 3 == 0:  # long letter runs across tile borders, a multi-byte letter right at a border
                docs.append(b"x" * 3960 + "é".encode() * 40 + b" y" + "ب".encode() * 5000)
            a = _run_fused(host_sim, h1, rule, docs, nthreads=rng.choice([1, 4, 8]))
            b = _run_fused(host_sim_noskip, h2, rule, docs, nthreads=4)
            assert a["nout"].tolist() == b["nout"].tolist() and a["ctr"].tolist() == b["ctr"].tolist()
            assert np.array_equal(a["wl"], b["wl"]) and np.array_equal(a["ids"], b["ids"])
            assert np.array_equal(a["dto"], b["dto"]) and np.array_equal(a["wf"], b["wf"])
        host_sim.sim_vocab_destroy(ctypes.c_void_p(h1))
        host_sim_noskip.sim_vocab_destroy(ctypes.c_void_p(h2))
