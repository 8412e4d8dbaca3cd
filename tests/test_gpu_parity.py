"""GPU parity tests: the CUDA path, called through the reference-shaped Python surface and the C ABI, against
the committed golden vectors (outputs of the unmodified reference) and against the CPU oracle on the same
seeded inputs.  Bit-exact: token ids, per-word lengths, tie choices, counters.  Run with ``-m gpu`` on a B200.
"""
import random

import numpy as np
import pytest
import torch

from conftest import load_golden
from helpers import bytelevel_table, pack, py_roundtrip_ok, sha1_json, vocab_bytes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.cuda.current_device()


def _to_dev(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


# ------------------------------------------------------------------------------------------------
# reference-shaped surface: packages.dp_tokenize
# ------------------------------------------------------------------------------------------------
def test_known_answers_through_reference_surface(dev):
    """tests/test_tokenization_algorithms.py:32-48 verbatim, plus the full returns the reference gives."""
    from packages.dp_tokenize import compute_shortest_tokenizations, obtain_longest_token
    V = ["un", "desirable"] + ["und", "esirable"] + list("undesirable")
    shortest_tokenizations, shortest_length = compute_shortest_tokenizations("undesirable", V, False, "")
    assert ["un", "desirable"] in shortest_tokenizations
    assert ["und", "esirable"] in shortest_tokenizations
    V = list("desireableish") + ["desire", "able", "ish"] + ["des", "ireable", "ish"]
    shortest_tokenizations, shortest_length = compute_shortest_tokenizations("desireableish", V, False, "")
    assert ["desire", "able", "ish"] in shortest_tokenizations
    assert ["des", "ireable", "ish"] in shortest_tokenizations

    k = load_golden("known_answers.json")
    for r in k["shortest"]:
        got = compute_shortest_tokenizations(r["s"], r["vocab"], False, "")
        assert got == (r["all"], r["len"])
        assert obtain_longest_token(got[0]) == r["sel"]
    for r in k["phantom"]:
        assert compute_shortest_tokenizations(r["s"], set(r["vocab"]), False, "") == ([], r["len"])
    sm = k["strip_marker"]
    assert compute_shortest_tokenizations(sm["s"], sm["vocab"], True, sm["marker"]) == (sm["all"], sm["len"])
    # a 5th positional argument is accepted (tokenizer_utils.py:71)
    assert compute_shortest_tokenizations("abcd", k["shortest"][2]["vocab"], False, None, 1)[1] == 2
    with pytest.raises(IndexError):
        compute_shortest_tokenizations("", ["a"], False, "")
    with pytest.raises(ValueError):
        obtain_longest_token([])
    # list-of-units input with multi-character units (test_llama_tokenizer builds such lists, :58-68)
    units = ["▁t", "h", "e", "▁", "w"]
    vocab = {"▁t", "h", "e", "▁", "w", "▁the", "he", "▁w"}
    from oracle import dp_oracle
    assert compute_shortest_tokenizations(units, vocab, False, None) == dp_oracle.enumerate_shortest(units, vocab)


def test_c0_full_return_every_word(dev):
    """Config C0: the FULL return of compute_shortest_tokenizations for 10,000 words (first 2,500 through the
    enumerate-all API; all 10,000 through the batched select-on-device path)."""
    from packages.dp_tokenize import compute_shortest_tokenizations, obtain_longest_token
    from dptok.engine import Engine
    from dptok.vocab import CompiledVocab
    g = load_golden("c0_toy.json.gz")
    vocab = set(g["vocab"])
    for r in g["rows"][:2500]:
        got, length = compute_shortest_tokenizations(r["w"], vocab, False, "")
        assert length == r["len"] and len(got) == r["n"] and sha1_json(got) == r["sha1"], r["w"]
        if got:
            assert obtain_longest_token(got) == r["sel"]
    eng = Engine(CompiledVocab.from_token_map({t: k for k, t in enumerate(g["vocab"])}, "spm"))
    text, offs = pack([r["w"].encode() for r in g["rows"]])
    res = eng.encode_words(_to_dev(text, dev), _to_dev(offs, dev), want_tok_offs=True)
    ids = res.ids.cpu().numpy()
    lens = res.word_lens.cpu().numpy()
    flags = res.word_flags.cpu().numpy()
    to = res.word_tok_offs.cpu().numpy()
    n_untok = 0
    for k, r in enumerate(g["rows"]):
        assert lens[k] == r["len"]
        if r["sel"] is None:
            assert flags[k] & 1 and to[k + 1] == to[k]
            n_untok += 1
        else:
            assert not (flags[k] & 1)
            assert [g["vocab"][i] for i in ids[to[k]:to[k + 1]]] == r["sel"]
    c = res.counters.cpu().tolist()
    assert c == [len(text), len(g["rows"]), len(ids), n_untok]


# ------------------------------------------------------------------------------------------------
# reference-shaped surface: packages.tokenizer_utils
# ------------------------------------------------------------------------------------------------
def test_llama_adapter_golden(dev):
    from dptok import assets
    from packages.tokenizer_utils import dp_tokenize_llama, pretokenize_with_llama, _BiMap
    g = load_golden("llama_adapter.json.gz")
    tok = assets.load_hf(g["tokenizer"])
    dp_encode_llama, invert_dp_tokenize = dp_tokenize_llama(tok)
    split = pretokenize_with_llama(tok, _BiMap(tok.get_vocab()))
    for r in g["rows"]:
        encoding = dp_encode_llama(r["text"])
        assert encoding == r["ids"], r["text"][:60]
        assert invert_dp_tokenize(encoding) == r["text"]
        assert len(encoding) <= r["default_len"]
        assert split(r["text"]) == r["words"]
    # batched call = same ids
    texts = [r["text"] for r in g["rows"]]
    assert dp_encode_llama.batch(texts) == [r["ids"] for r in g["rows"]]
    # 'raw' option (tokenizer_utils.py:33-50)
    enc_raw, _ = dp_tokenize_llama(tok, "raw")
    for r in g["raw"]:
        if "ids" in r:
            assert enc_raw(r["text"]) == r["ids"]


def test_bytelevel_adapter_golden(dev):
    from dptok import assets
    from packages.tokenizer_utils import dp_tokenize_bloom
    g = load_golden("bytelevel_adapter.json.gz")
    for name, rows in g.items():
        tok = assets.load_hf(name)
        dp_encode_bloom, invert_dp_tokenize = dp_tokenize_bloom(tok, None)
        for r in rows:
            encoding_dp = dp_encode_bloom(r["text"])
            assert encoding_dp == r["ids"], (name, r["text"][:60])
            assert invert_dp_tokenize(encoding_dp) == r["text"]
            assert len(encoding_dp) <= r["default_len"]
        assert dp_encode_bloom.batch([r["text"] for r in rows]) == [r["ids"] for r in rows]


# ------------------------------------------------------------------------------------------------
# C ABI at scale against the oracle
def test_min_tokens_for_string_on_gpu(dev):
    """a6: inspect_tokenizer.py:77-86 through packages.dp_tokenize.min_tokens_for_string (dpt_min_tokens_word): the five
    vectors of tests/test_tokenization_algorithms.py:14-30, infinity for untokenizable input (where len_dp of the 4-arg
    API carries the phantom value instead), list-of-units input, and 2,000 random C0 words against the golden values the
    unmodified reference produced."""
    from packages.dp_tokenize import compute_shortest_tokenizations, min_tokens_for_string
    k = load_golden("known_answers.json")
    assert [min_tokens_for_string(r["s"], set(r["vocab"])) for r in k["min_tokens"]] == [2, 6, 3, 2, 3]
    for r in k["phantom"]:  # a real segmentation may exist although the 4-arg API returns ([], phantom)
        from oracle import dp_oracle
        assert min_tokens_for_string(r["s"], set(r["vocab"])) == dp_oracle.min_tokens(r["s"], set(r["vocab"]))
    assert min_tokens_for_string("qrsTUV", {"qr", "s", "T", "U", "V", "rsTUV"}) == 5      # phantom len_dp says 2
    assert compute_shortest_tokenizations("qrsTUV", {"qr", "s", "T", "U", "V", "rsTUV"}, False, "")[1] == 2
    assert min_tokens_for_string("xyz", {"x", "y"}) == float("inf")
    assert min_tokens_for_string("", {"x"}) == 0
    units = ["▁t", "h", "e", "▁", "w"]
    assert min_tokens_for_string(units, {"▁t", "h", "e", "▁", "w", "▁the", "he", "▁w"}) == 2
    assert min_tokens_for_string(["ab", "c"], {"a", "bc", "ab", "c"}) == 2
    assert min_tokens_for_string(["ab", "c"], {"a", "bc"}) == float("inf")   # "a" + "bc" would split inside the unit "ab"
    assert min_tokens_for_string("abc", {"a", "bc"}) == 2
    c0 = load_golden("c0_toy.json.gz")
    vocab = frozenset(c0["vocab"])
    for r in c0["rows"][:2000]:
        got = min_tokens_for_string(r["w"], vocab)
        assert (None if got == float("inf") else got) == r["min_tokens"], r["w"]


# ------------------------------------------------------------------------------------------------
def _llama_engine(name, dev):
    from dptok import assets
    from dptok.engine import Engine
    from dptok.vocab import CompiledVocab
    tok = assets.load_hf(name)
    t2i = tok.get_vocab()
    return tok, t2i, Engine(CompiledVocab.from_token_map(t2i, "spm"), dev)


def test_corpus_path_vs_oracle_llama32k(dev):
    """8 MB of S2ORC-shaped text, Llama-2-shaped 32k vocab: ids, per-word lengths, flags, counters and document
    offsets bit-exact against the C oracle run on independently normalised words; the device word split is
    additionally checked against the tokenizer-driven split of the reference on a document sample."""
    from dptok import _cabi, synth
    from oracle import adapters
    from oracle.c_oracle import COracle
    tok, t2i, eng = _llama_engine("llama2_32k", dev)
    text, doc_offs = synth.gen_documents(8_000_000, seed=0, newline_headers=True)
    raw = text.tobytes()
    n_docs = len(doc_offs) - 1
    res = eng.encode_corpus(_to_dev(text, dev), _to_dev(doc_offs, dev), _cabi.RULE_SPM_LLAMA)
    vocab = set(t2i)
    words = []
    first_word = []
    for d in range(n_docs):
        first_word.append(len(words))
        words += [w.encode() for w in adapters.spm_normalise(raw[doc_offs[d]:doc_offs[d + 1]].decode(), vocab)]
    wtext, woffs = pack(words)
    o_ids, o_lens, o_untok = COracle(vocab_bytes(t2i, "spm"), 1).encode_words(wtext, woffs)
    assert res.n_words == len(words)
    assert np.array_equal(res.word_lens.cpu().numpy(), o_lens)
    assert np.array_equal(res.word_flags.cpu().numpy() & 1, o_untok)
    assert np.array_equal(res.ids.cpu().numpy(), o_ids)
    assert res.counters.cpu().tolist() == [len(raw), len(words), len(o_ids), int(o_untok.sum())]
    assert not res.doc_flags.cpu().numpy().any()
    tok_of_word = np.concatenate([[0], np.cumsum(np.where(o_untok == 0, o_lens, 0))])
    assert np.array_equal(res.doc_tok_offs.cpu().numpy(), tok_of_word[first_word + [len(words)]])
    # the rule itself vs. the reference's tokenizer-driven split
    for d in random.Random(0).sample(range(n_docs), 60):
        doc = raw[doc_offs[d]:doc_offs[d + 1]].decode()
        assert [w.decode() for w in words[first_word[d]:(first_word + [len(words)])[d + 1]]] == adapters.llama_words(tok, doc)
    # decode + round trip on device (main_analyze_s2orc.py:85)
    ok = eng.roundtrip_ok(res, _to_dev(text, dev), _to_dev(doc_offs, dev), skip_bos=True)
    assert bool(ok.all())
    # a corrupted id must be caught
    bad = res.ids.clone()
    bad[5] = bad[5] + 1 if int(bad[5]) + 1 < len(t2i) else 3
    res_bad = type(res)(bad, res.word_lens, res.word_flags, None, res.counters, res.n_ids, res.n_words, res.doc_tok_offs,
                        res.doc_flags)
    assert not bool(eng.roundtrip_ok(res_bad, _to_dev(text, dev), _to_dev(doc_offs, dev), skip_bos=True)[0])


def test_roundtrip_check_on_damaged_ids(dev):
    """dpt_roundtrip_check (one warp per document, 32 tokens per round): every document gets one id replaced at a random
    position (any round, any lane) or stays intact; the per-document verdicts equal a Python decode of the same ids.
    SentencePiece (skip_bos, byte tokens, the dropped leading space) and byte-level vocabularies."""
    from dptok import _cabi, assets, synth
    from dptok.engine import Engine
    from dptok.vocab import CompiledVocab
    rng = random.Random(3)
    tok, t2i, eng = _llama_engine("llama2_32k", dev)
    text, doc_offs = synth.gen_documents(1_500_000, seed=6, newline_headers=True)
    v2i = {t: k for k, t in enumerate(assets.load_spec("gpt2_50k")["model"]["vocab"])}
    eng_b = Engine(CompiledVocab.from_token_map(v2i, "bytelevel"), dev)
    text_b, offs_b = synth.gen_sentence_pairs(600_000, seed=6)
    for engine, vb, txt, offs, rule, spm in ((eng, vocab_bytes(t2i, "spm"), text, doc_offs, _cabi.RULE_SPM_LLAMA, True),
                                             (eng_b, vocab_bytes(v2i, "bytelevel"), text_b, offs_b, _cabi.RULE_GPT2, False)):
        d_text, d_offs = _to_dev(txt, dev), _to_dev(offs, dev)
        res = engine.encode_corpus(d_text, d_offs, rule)
        assert bool(engine.roundtrip_ok(res, d_text, d_offs, skip_bos=spm).all())
        ids = res.ids.cpu().numpy().copy()
        dto = res.doc_tok_offs.cpu().numpy()
        raw = txt.tobytes()
        id2tok = {i: t for t, i in vb.items()}
        all_ids = sorted(id2tok)
        n_docs = len(offs) - 1
        for d in range(n_docs):
            if rng.random() < 0.7 and dto[d + 1] - dto[d] > 1:
                ids[rng.randrange(dto[d] + (1 if spm else 0), dto[d + 1])] = rng.choice(all_ids + [-1, 2 ** 31 - 1])
        expect = np.array([py_roundtrip_ok(id2tok, ids[dto[d]:dto[d + 1]], raw[offs[d]:offs[d + 1]], spm, spm)
                           for d in range(n_docs)])
        assert expect.any() and not expect.all()
        bad = type(res)(_to_dev(ids, dev), res.word_lens, res.word_flags, None, res.counters, res.n_ids, res.n_words,
                        res.doc_tok_offs, res.doc_flags)
        got = engine.roundtrip_ok(bad, d_text, d_offs, skip_bos=spm).cpu().numpy() != 0
        assert np.array_equal(got, expect)


def test_presplit_bytelevel_vs_oracle_gpt2_50k(dev):
    """Byte-level path (GPT-2-shaped 50k vocab): words from the installed `tokenizers` pre-tokenizer
    (tokenizer_utils.py:157-159), DP on the GPU, against the C oracle."""
    from dptok import assets, synth
    from dptok.engine import Engine
    from dptok.vocab import CompiledVocab
    from oracle.c_oracle import COracle
    tok = assets.load_tokenizer("gpt2_50k")
    v2i = {t: k for k, t in enumerate(assets.load_spec("gpt2_50k")["model"]["vocab"])}
    eng = Engine(CompiledVocab.from_token_map(v2i, "bytelevel"), dev)
    text, doc_offs = synth.gen_sentence_pairs(3_000_000, seed=0)
    raw = text.tobytes()
    u2b = {c: b for b, c in bytelevel_table().items()}
    words = []
    for d in range(len(doc_offs) - 1):
        for piece, _span in tok.pre_tokenizer.pre_tokenize_str(raw[doc_offs[d]:doc_offs[d + 1]].decode()):
            words.append(bytes(u2b[c] for c in piece))
    assert b"".join(words) == raw          # byte-level pieces tile the text
    wtext, woffs = pack(words)
    res = eng.encode_words(_to_dev(wtext, dev), _to_dev(woffs, dev))
    o_ids, o_lens, o_untok = COracle(vocab_bytes(v2i, "bytelevel"), 0).encode_words(wtext, woffs)
    assert np.array_equal(res.ids.cpu().numpy(), o_ids)
    assert np.array_equal(res.word_lens.cpu().numpy(), o_lens)
    assert not o_untok.any() and not (res.word_flags.cpu().numpy() & 1).any()


def test_edge_cases_long_oov_untokenizable(dev):
    """Ragged inputs: 1-byte words, words longer than the local-state limit (long path, thousands of bytes),
    out-of-vocabulary characters (literal <0xHH> expansion), untokenizable words (phantom lengths), a vocabulary
    with long tokens; capacity retry paths."""
    from dptok.engine import Engine
    from dptok.vocab import CompiledVocab
    from oracle.c_oracle import COracle
    rng = random.Random(11)
    alpha = "abcdeé▁日"
    toks = set(alpha[:6])
    for _ in range(400):
        toks.add("".join(rng.choice(alpha) for _ in range(rng.randint(2, 9))))
    toks.add("ab" * 40)            # an 80-byte token
    toks.add("é" * 50)             # 100 bytes, 50 code points
    t2i = {t: k + 5 for k, t in enumerate(sorted(toks))}
    eng = Engine(CompiledVocab.from_token_map(t2i, "spm"), dev)
    words = []
    for k in range(6000):
        n = rng.choice([1, 1, 2, 3, 5, 8, 13, 21, 34, 63, 64, 65, 66, 100, 200])
        if k % 500 == 0:
            n = rng.randint(1500, 6000)
        w = "".join(rng.choice(alpha) for _ in range(n))
        if k % 97 == 0:
            w = "ab" * rng.randint(30, 90)
        if k % 89 == 0:
            w = "é" * rng.randint(40, 120)
        words.append(w.encode())
    wtext, woffs = pack(words)
    o_ids, o_lens, o_untok = COracle(vocab_bytes(t2i, "spm"), 1).encode_words(wtext, woffs)
    assert 0 < o_untok.sum() < len(words)    # '▁' and '日' alone are not tokens: some words are untokenizable
    for ids_cap in (None, 16):               # 16 forces the ids-capacity retry
        res = eng.encode_words(_to_dev(wtext, dev), _to_dev(woffs, dev), ids_cap=ids_cap)
        assert np.array_equal(res.word_lens.cpu().numpy(), o_lens)
        assert np.array_equal(res.word_flags.cpu().numpy() & 1, o_untok)
        assert np.array_equal(res.ids.cpu().numpy(), o_ids)
    long_flag = (res.word_flags.cpu().numpy() & 4) != 0
    assert np.array_equal(long_flag, (woffs[1:] - woffs[:-1]) > 64)


def test_encode_words_id_stash_and_its_fallback(dev):
    """dpt_encode_words solves every word ONCE: the ids wait in a stash of one int32 per text byte (sized by
    n_text_bytes) until the scan has placed them.  A caller whose n_text_bytes understates the text still gets the
    same ids (words beyond the stash are solved again at placement time).  C ABI called directly with the full extent,
    half of it (a word straddles the end of the stash) and 0; all three bit-exact against the C oracle."""
    from dptok import _cabi
    from dptok.engine import Engine, _ptr
    from dptok.vocab import CompiledVocab
    from oracle.c_oracle import COracle
    rng = random.Random(5)
    alpha = "abcdeé▁"
    toks = set(alpha[:6])
    for _ in range(300):
        toks.add("".join(rng.choice(alpha) for _ in range(rng.randint(2, 7))))
    t2i = {t: k for k, t in enumerate(sorted(toks))}
    eng = Engine(CompiledVocab.from_token_map(t2i, "spm"), dev)
    words = ["".join(rng.choice(alpha) for _ in range(rng.choice([1, 2, 3, 5, 8, 13, 21, 40, 64, 65, 90]))).encode()
             for _ in range(20000)]
    wtext, woffs = pack(words)
    o_ids, o_lens, o_untok = COracle(vocab_bytes(t2i, "spm"), 1).encode_words(wtext, woffs)
    d_text, d_offs = _to_dev(wtext, dev), _to_dev(woffs, dev)
    n_words, n_bytes = len(words), len(wtext)
    lib = _cabi.lib
    for extent in (n_bytes, n_bytes // 2 + 3, 0):
        ids = torch.full((n_bytes + 64,), -7, dtype=torch.int32, device=dev)
        lens = torch.empty(n_words, dtype=torch.int32, device=dev)
        flags = torch.empty(n_words, dtype=torch.uint8, device=dev)
        counters = torch.empty(4, dtype=torch.int64, device=dev)
        n_out = torch.empty(8, dtype=torch.int64, device=dev)
        ws = torch.empty(lib.dpt_encode_words_workspace(extent, n_words, 1) + 12 * n_bytes, dtype=torch.uint8, device=dev)
        _cabi.check(lib.dpt_encode_words(eng.vocab.handle, _ptr(d_text), _ptr(d_offs), n_words, extent, _ptr(ids),
                                         ids.numel(), _ptr(lens), _ptr(flags), None, _ptr(counters), _ptr(n_out),
                                         _ptr(ws), ws.numel(), None))
        torch.cuda.synchronize()
        h = n_out.cpu().tolist()
        assert h[_cabi.NOUT_IDS] == len(o_ids) and h[_cabi.NOUT_POOL_REQ] <= h[_cabi.NOUT_POOL_CAP]
        assert np.array_equal(ids[:len(o_ids)].cpu().numpy(), o_ids), extent
        assert (ids[len(o_ids):] == -7).all()
        assert np.array_equal(lens.cpu().numpy(), o_lens)
        assert np.array_equal(flags.cpu().numpy() & 1, o_untok)
        assert counters.cpu().tolist() == [extent, n_words, len(o_ids), int(o_untok.sum())]


def test_undersized_word_table_odd_words_and_worst_case_retry(dev):
    """The word table holds n_bytes / 48 slots.  (1) a corpus whose DISTINCT words outnumber the slots of their
    neighbourhoods: the occurrences that find 16 probed slots taken become odd words (solved per occurrence from the raw
    text) - on the GPU, not only in the host emulation; (2) a corpus of distinct words only: the odd-word list overflows
    too, the pass reports it and the engine runs again with worst-case sizes.  Both bit-exact against the C oracle."""
    from dptok import _cabi
    from oracle import adapters
    from oracle.c_oracle import COracle
    tok, t2i, eng = _llama_engine("llama2_32k", dev)
    vocab = set(t2i)
    orc = COracle(vocab_bytes(t2i, "spm"), 1)
    rng = random.Random(11)
    letters = "abcdefghijklmnopqrstuvwxyz"

    def fresh(k):
        return "".join(rng.choice(letters) for _ in range(rng.randint(3, 11))) + format(k, "x")

    def corpus(n_words, n_distinct, doc_words=180):
        pool = [fresh(k) for k in range(n_distinct)]
        seq = pool + [rng.choice(pool) for _ in range(n_words - n_distinct)]
        rng.shuffle(seq)
        docs = [" ".join(seq[k:k + doc_words]).encode() for k in range(0, len(seq), doc_words)]
        offs = np.zeros(len(docs) + 1, dtype=np.int64)
        np.cumsum([len(d) for d in docs], out=offs[1:])
        return np.frombuffer(b"".join(docs), dtype=np.uint8).copy(), offs, docs

    for n_words, n_distinct, expect_worst in ((130_000, 36_000, 0), (120_000, 120_000, 1)):
        text, offs, docs = corpus(n_words, n_distinct)
        _tok, _t2i, eng = _llama_engine("llama2_32k", dev)   # a fresh engine: no table-size hint from the corpus before
        res = eng.encode_corpus(_to_dev(text, dev), _to_dev(offs, dev), _cabi.RULE_SPM_LLAMA)
        h = eng.last_n_out
        assert eng.last_worst == expect_worst, (h, len(text))
        if not expect_worst:
            # > 2 % of the words were odd: the engine asks for the roomy table next time, and with it nothing overflows
            again = eng.encode_corpus(_to_dev(text, dev), _to_dev(offs, dev), _cabi.RULE_SPM_LLAMA)
            assert eng.last_worst == 2 and eng.last_n_out[6] < 50, eng.last_n_out
            assert torch.equal(again.ids, res.ids) and torch.equal(again.word_lens, res.word_lens)
        if not expect_worst:
            n_slots = 4096
            while n_slots < len(text) // 48:
                n_slots *= 2
            assert n_distinct > n_slots and h[6] > 1000, ("the table was meant to overflow into odd words", n_slots, h)
        words = []
        for d in docs:
            words += [w.encode() for w in adapters.spm_normalise(d.decode(), vocab)]
        wtext, woffs = pack(words)
        o_ids, o_lens, o_untok = orc.encode_words(wtext, woffs)
        assert res.n_words == len(words)
        assert np.array_equal(res.word_lens.cpu().numpy(), o_lens)
        assert np.array_equal(res.ids.cpu().numpy(), o_ids)
        assert res.counters.cpu().tolist() == [len(text), len(words), len(o_ids), int(o_untok.sum())]


def test_spm_rule_oov_and_ambiguous_docs(dev):
    """Device rule on text with OOV characters (CJK, dashes), newlines/tabs, and documents whose split is
    ambiguous (double spaces): flagged, and the public adapter still returns the reference's ids for them."""
    from dptok import _cabi, assets
    from dptok.engine import pack_documents
    from oracle import adapters
    tok, t2i, eng = _llama_engine("llama2_2k", dev)
    from packages.tokenizer_utils import dp_tokenize_llama
    enc, dec = dp_tokenize_llama(tok)
    rng = random.Random(3)
    pieces = ["plai", "gout", "é", "ï", "日", "本", "\n", "\t", ",", "Zeta", "(x)", "12", "—", "naïve", "%", "trot"]
    docs = []
    for k in range(400):
        ws = ["".join(rng.choice(pieces) for _ in range(rng.randint(1, 4))) for _ in range(rng.randint(1, 12))]
        sep = "  " if k % 10 == 0 else " "
        docs.append(sep.join(ws))
    text, offs = pack_documents([d.encode() for d in docs])
    res = eng.encode_corpus(_to_dev(text, dev), _to_dev(offs, dev), _cabi.RULE_SPM_LLAMA)
    flags = res.doc_flags.cpu().numpy()
    ids = res.ids.cpu().numpy()
    dto = res.doc_tok_offs.cpu().numpy()
    for k, d in enumerate(docs):
        expect = adapters.llama_encode(tok, d)
        assert bool(flags[k]) == ("  " in d)
        if not flags[k]:
            assert ids[dto[k]:dto[k + 1]].tolist() == expect
        assert enc(d) == expect and dec(expect) == d
    assert enc.batch(docs[:50]) == [adapters.llama_encode(tok, d) for d in docs[:50]]


def test_marker_runs_split_on_the_device(dev):
    """SURVEY 8 row f1 on the GPU: a tokenizer WITH whitespace-run tokens ("▁▁" is its first merge, like the real Llama-2
    vocabulary).  The adapter hands the merge table to the compiled vocabulary; the corpus pipeline then cuts runs of spaces
    / U+2581 where the default tokenizer does (tokenizer_utils.py:7-31): NO document is flagged DPT_DF_AMBIGUOUS, ids equal
    the oracle adapter's (default-tokenizer-driven split + DP per word), through the corpus call, the single-string call and
    .batch().  Without the merge table the same documents are flagged and still come out right through the host split."""
    from dptok import _cabi
    from dptok.engine import Engine, pack_documents
    from dptok.vocab import CompiledVocab
    from oracle import adapters
    from packages.tokenizer_utils import dp_tokenize_llama
    from test_host_sim import _tokenizer_with_marker_run_tokens
    tok, t2i, mid = _tokenizer_with_marker_run_tokens()
    assert adapters.llama_words(tok, "a   b")[1:] == ["▁a", "▁▁", "▁b"] and adapters.llama_words(tok, "a  b")[1:] == ["▁a", "▁▁b"]
    rng = random.Random(23)
    words = ["plai", "gout", "trot", "Zeta", "a", "I", "naïve", "(x)", "12", "日本", "the", "tion", "\n", "x\ny", "é"]
    docs = ["a  b", "a   b", "a    b", "  a", " a", "   a", "a ", "a  ", " ", "  ", "    ", "x\n\n  indented   text", "a ▁b", "▁▁a",
            "a" + " " * 37 + "b", "  é  日  ", "a" + " " * 700 + "b"]
    for _ in range(1500):
        parts = []
        for _k in range(rng.randint(1, 14)):
            parts.append(rng.choice(words))
            parts.append(rng.choice([" ", " ", " ", "  ", "   ", "    ", "▁", " ▁", "      "]))
        d = "".join(parts)
        if rng.random() < 0.3:
            d = rng.choice([" ", "  ", "   "]) + d
        if rng.random() < 0.5:
            d = d.rstrip(" ▁") or "a"
        docs.append(d)
    expect = [adapters.llama_encode(tok, d) for d in docs]
    eng = Engine(CompiledVocab.from_token_map(t2i, "spm").set_merges(mid), dev)
    text, offs = pack_documents([d.encode() for d in docs])
    res = eng.encode_corpus(_to_dev(text, dev), _to_dev(offs, dev), _cabi.RULE_SPM_LLAMA)
    flags = res.doc_flags.cpu().numpy()
    ids = res.ids.cpu().numpy()
    dto = res.doc_tok_offs.cpu().numpy()
    long_run = docs.index("a" + " " * 700 + "b")   # longer than the device splits (PB_SEG_MAX symbols): flagged for the host
    for k, d in enumerate(docs):
        if k == long_run:
            assert flags[k]
            continue
        assert not flags[k], d
        assert ids[dto[k]:dto[k + 1]].tolist() == expect[k], (d, adapters.llama_words(tok, d))
    enc, _dec = dp_tokenize_llama(tok)
    assert getattr(enc.engine.vocab, "n_merges", 0) == len(mid)
    for k in list(range(20)) + [long_run]:
        assert enc(docs[k]) == expect[k]
    assert enc.batch(docs[:200]) == expect[:200]
    # no merge table: flagged, and the adapter falls back to the tokenizer's split on the host
    eng0 = Engine(CompiledVocab.from_token_map(t2i, "spm"), dev)
    res0 = eng0.encode_corpus(_to_dev(text, dev), _to_dev(offs, dev), _cabi.RULE_SPM_LLAMA)
    f0 = res0.doc_flags.cpu().numpy()
    assert f0[0] and f0[1] and f0[3] and not f0[6]


def test_unaligned_corpus_buffer(dev):
    """A corpus that does not start on a 16-byte boundary (a shard cut out of a larger device buffer, as the multi-GPU
    driver does): same result as the aligned copy, for the SentencePiece and a byte-level rule."""
    from dptok import _cabi, assets, synth
    from dptok.engine import Engine
    from dptok.vocab import CompiledVocab
    tok, t2i, eng = _llama_engine("llama2_32k", dev)
    text, doc_offs = synth.gen_documents(1_500_000, seed=5, newline_headers=True)
    spec = assets.load_spec("gpt2_3k")
    beng = Engine(CompiledVocab.from_token_map({t: k for k, t in enumerate(spec["model"]["vocab"])}, "bytelevel"), dev)
    for engine, rule in ((eng, _cabi.RULE_SPM_LLAMA), (beng, _cabi.RULE_GPT2)):
        ref = engine.encode_corpus(_to_dev(text, dev), _to_dev(doc_offs, dev), rule)
        for mis in (1, 3, 4, 9):
            buf = torch.zeros(len(text) + 64, dtype=torch.uint8, device=dev)
            buf[mis:mis + len(text)] = _to_dev(text, dev)
            view = buf[mis:mis + len(text)]
            assert view.data_ptr() % 16 == mis
            got = engine.encode_corpus(view, _to_dev(doc_offs, dev), rule)
            assert got.n_ids == ref.n_ids and got.n_words == ref.n_words
            assert torch.equal(got.ids, ref.ids) and torch.equal(got.word_lens, ref.word_lens)
            assert torch.equal(got.doc_tok_offs, ref.doc_tok_offs)


def test_sharded_driver_with_the_real_engine(dev):
    """SURVEY 8 row e with the real kernels: `ShardedTokenizer.run_global` cuts ONE corpus into contiguous document ranges
    of equal bytes (`shard_bounds`), every "rank" tokenizes its shard with `Engine.encode_corpus`; the shards' ids,
    concatenated in rank order, must equal the single-pass ids, and the counters must add up to the single-pass counters
    (what the NCCL sum gives at N > 1; here the ranks run one after the other in one process).  Both rules."""
    from dptok import _cabi, assets, synth
    from dptok.engine import Engine
    from dptok.sharded import ShardedTokenizer, shard_bounds
    from dptok.vocab import CompiledVocab
    tok, t2i, eng = _llama_engine("llama2_32k", dev)
    spec = assets.load_spec("gpt2_3k")
    beng = Engine(CompiledVocab.from_token_map({t: k for k, t in enumerate(spec["model"]["vocab"])}, "bytelevel"), dev)
    text, doc_offs = synth.gen_documents(3_000_000, seed=8, newline_headers=True)
    for engine, rule in ((eng, _cabi.RULE_SPM_LLAMA), (beng, _cabi.RULE_GPT2)):
        whole = engine.encode_corpus(_to_dev(text, dev), _to_dev(doc_offs, dev), rule)
        want_ids = whole.ids.cpu().numpy()
        want_ctr = whole.counters.cpu().numpy()
        want_dto = whole.doc_tok_offs.cpu().numpy()

        def encode_fn(t, o):
            return engine.encode_corpus(_to_dev(np.ascontiguousarray(t), dev), _to_dev(np.ascontiguousarray(o), dev), rule)

        for world in (2, 3, 7):
            b = shard_bounds(doc_offs, world)
            ids, ctr, dto = [], np.zeros(4, dtype=np.int64), [np.zeros(1, dtype=np.int64)]
            for rank in range(world):
                res, stats = ShardedTokenizer(encode_fn, world=world, rank=rank).run_global(text, doc_offs)
                assert res is not None and (stats.bytes, stats.words, stats.tokens) == tuple(res.counters.cpu().tolist()[:3])
                ids.append(res.ids.cpu().numpy())
                ctr += res.counters.cpu().numpy()
                dto.append(res.doc_tok_offs.cpu().numpy()[1:] + dto[-1][-1])
                assert int(res.counters[0]) == int(doc_offs[b[rank + 1]] - doc_offs[b[rank]])
            assert np.array_equal(np.concatenate(ids), want_ids)
            assert np.array_equal(ctr, want_ctr)
            assert np.array_equal(np.concatenate(dto), want_dto)


def test_full_size_properties_100mb(dev):
    """BASELINE.json configs[1] at full size (100 MB, Llama-2-shaped 32k vocab): size-independent properties -
    device decode round-trips every document, counters are consistent, the DP never uses more tokens than the
    default BPE, two runs are identical - plus exact oracle parity on a 1,500-document sample."""
    from dptok import _cabi, synth
    from oracle import adapters
    from oracle.c_oracle import COracle
    tok, t2i, eng = _llama_engine("llama2_32k", dev)
    text, doc_offs = synth.gen_documents(100_000_000, seed=0)
    n_docs = len(doc_offs) - 1
    d_text, d_offs = _to_dev(text, dev), _to_dev(doc_offs, dev)
    res = eng.encode_corpus(d_text, d_offs, _cabi.RULE_SPM_LLAMA)
    c = res.counters.cpu().tolist()
    assert c[0] == len(text) and c[1] == res.n_words and c[2] == res.n_ids and c[3] == 0
    assert int(res.word_lens.sum()) == res.n_ids
    assert bool(eng.roundtrip_ok(res, d_text, d_offs, skip_bos=True).all())
    dto = res.doc_tok_offs.cpu().numpy()
    assert dto[0] == 0 and dto[-1] == res.n_ids and (np.diff(dto) > 0).all()
    res2 = eng.encode_corpus(d_text, d_offs, _cabi.RULE_SPM_LLAMA)
    assert torch.equal(res.ids, res2.ids) and torch.equal(res.word_lens, res2.word_lens)
    ids = res.ids.cpu().numpy()
    raw = text.tobytes()
    vocab = set(t2i)
    co = COracle(vocab_bytes(t2i, "spm"), 1)
    sample = random.Random(1).sample(range(n_docs), 1500)
    for d in sample:
        doc = raw[doc_offs[d]:doc_offs[d + 1]].decode()
        wtext, woffs = pack([w.encode() for w in adapters.spm_normalise(doc, vocab)])
        o_ids, _, _ = co.encode_words(wtext, woffs)
        assert np.array_equal(ids[dto[d]:dto[d + 1]], o_ids)
    for d in sample[:40]:
        doc = raw[doc_offs[d]:doc_offs[d + 1]].decode()
        assert dto[d + 1] - dto[d] <= len(tok.encode(doc))


def test_pipeline_equals_general_path(dev):
    """The deduplicating corpus pipeline and the general per-occurrence path give identical outputs on edge-case
    input (tiny documents, double spaces, raw U+2581, OOV characters, malformed UTF-8 cut by document boundaries,
    words far longer than a scan tile)."""
    from dptok import _cabi
    from dptok.engine import pack_documents
    tok, t2i, eng = _llama_engine("llama2_2k", dev)
    rng = random.Random(5)
    pieces = ["plai", "gout", "é", "ï", "日", "本", "\n", "\t", ",", "Zeta", "(x)", "12", "—", "naïve", "%", "trot", "▁", "a", "I"]
    for trial in range(6):
        docs = []
        for k in range(rng.choice([1, 50, 400, 5000])):
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
        if trial % 3 == 0:
            blob = bytes(rng.choice([0x20, 0x41, 0x62, 0xE2, 0x96, 0x81, 0xC3, 0xA9, 0x80, 0xFF, 0x0A, 0x63])
                         for _ in range(rng.randint(5000, 200000)))
            cuts = sorted(set(rng.randint(1, len(blob) - 1) for _ in range(rng.randint(0, 400))))
            docs = [blob[a:b] for a, b in zip([0] + cuts, cuts + [len(blob)])]
        text, offs = pack_documents(docs)
        d_text, d_offs = _to_dev(text, dev), _to_dev(offs, dev)
        a = eng.encode_corpus(d_text, d_offs, _cabi.RULE_SPM_LLAMA)
        b = eng.encode_corpus(d_text, d_offs, _cabi.RULE_SPM_LLAMA, force_general=True)
        assert a.n_ids == b.n_ids and a.n_words == b.n_words
        assert torch.equal(a.ids, b.ids) and torch.equal(a.word_lens, b.word_lens)
        assert torch.equal(a.word_flags & 1, b.word_flags & 1)
        assert torch.equal(a.doc_tok_offs, b.doc_tok_offs) and torch.equal(a.doc_flags, b.doc_flags)
        assert a.counters.tolist() == b.counters.tolist()
    docs = [b"x" * 3000 + b" y " + b"z" * 9000 + b" end", b"short doc"]
    text, offs = pack_documents(docs)
    d_text, d_offs = _to_dev(text, dev), _to_dev(offs, dev)
    a = eng.encode_corpus(d_text, d_offs, _cabi.RULE_SPM_LLAMA)
    b = eng.encode_corpus(d_text, d_offs, _cabi.RULE_SPM_LLAMA, force_general=True)
    assert torch.equal(a.ids, b.ids) and torch.equal(a.word_lens, b.word_lens) and a.n_words == 8
    assert bool(eng.roundtrip_ok(a, d_text, d_offs, skip_bos=True).all())


def test_host_chunked_path_equals_resident_path(dev):
    """Engine.encode_corpus_host (pinned host text, chunks on 3 streams, ids back on the host) == the device-resident
    single-batch path: ids, document offsets, flags, counters."""
    from dptok import _cabi, synth
    tok, t2i, eng = _llama_engine("llama2_32k", dev)
    text, doc_offs = synth.gen_documents(9_000_000, seed=4, newline_headers=True)
    text = text.copy()
    text[doc_offs[7]:doc_offs[7] + 2] = 0x20          # a document that starts with two spaces: ambiguous flag
    res = eng.encode_corpus(_to_dev(text, dev), _to_dev(doc_offs, dev), _cabi.RULE_SPM_LLAMA)
    h_text = torch.from_numpy(text).pin_memory()
    for chunk in (1 << 20, 3_500_000, 64 << 20):
        hr = eng.encode_corpus_host(h_text, doc_offs, _cabi.RULE_SPM_LLAMA, chunk_bytes=chunk)
        assert hr.n_ids == res.n_ids and hr.n_chunks == max(1, -(-len(text) // chunk)) or hr.n_chunks >= 1
        assert np.array_equal(hr.ids.numpy(), res.ids.cpu().numpy())
        assert np.array_equal(hr.doc_tok_offs, res.doc_tok_offs.cpu().numpy())
        assert np.array_equal(hr.doc_flags, res.doc_flags.cpu().numpy()) and hr.doc_flags[7] == 1
        assert hr.counters.tolist() == res.counters.cpu().tolist()
    # the three-stream variant (scan / DP / emit of neighbouring ranges overlapped) and the compact uint16 ids
    ref_ids = res.ids.cpu().numpy()
    for rep in range(2):
        hr = eng.encode_corpus_host(h_text, doc_offs, _cabi.RULE_SPM_LLAMA, chunk_bytes=1 << 20, overlap=True)
        assert np.array_equal(hr.ids.numpy(), ref_ids) and hr.counters.tolist() == res.counters.cpu().tolist()
        assert np.array_equal(hr.doc_tok_offs, res.doc_tok_offs.cpu().numpy())
    h16 = eng.encode_corpus_host(h_text, doc_offs, _cabi.RULE_SPM_LLAMA, chunk_bytes=3_500_000, ids_dtype=torch.uint16)
    assert h16.ids.dtype == torch.uint16 and h16.n_ids == res.n_ids
    assert np.array_equal(h16.ids.view(torch.int16).numpy().view(np.uint16).astype(np.int32), ref_ids)


def test_narrow_ids_u16_overflow_is_reported(dev):
    """dpt_narrow_ids_u16: ids beyond 16 bits become 0xFFFF and are counted; the count comes from device memory."""
    import ctypes as C
    from dptok import _cabi
    ids = torch.tensor([1, 65535, 65536, 7, 200000, 3, 4, 5, 6, 9, 10], dtype=torch.int32, device=dev)
    n = torch.tensor([9], dtype=torch.int64, device=dev)
    out = torch.zeros(16, dtype=torch.uint16, device=dev)
    ovf = torch.zeros(1, dtype=torch.int64, device=dev)
    _cabi.check(_cabi.lib.dpt_narrow_ids_u16(C.c_void_p(ids.data_ptr()), C.c_void_p(n.data_ptr()), 11,
                                             C.c_void_p(out.data_ptr()), C.c_void_p(ovf.data_ptr()),
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    got = out.cpu().view(torch.int16).numpy().view(np.uint16).astype(np.int64).tolist()
    assert got[:9] == [1, 65535, 65535, 7, 65535, 3, 4, 5, 6] and got[9:] == [0] * 7
    assert int(ovf.item()) == 2


def test_pad_batch_on_device(dev):
    """Row f3: padded input_ids / attention_mask built on the device == what main_analyze_s2orc.py:87-89 +
    DataCollatorWithPadding, and the input_ids + labels collator of main_biomed_translation.py:104-124, build on the host."""
    from dptok import _cabi, synth
    from dptok.engine import pack_documents
    tok, t2i, eng = _llama_engine("llama2_2k", dev)
    docs = [d.encode() for d in synth.sample_text(40_000, seed=9)]
    tgt = [d[::-1].replace(b"  ", b" ").strip() or b"x" for d in docs]
    text, offs = pack_documents(docs)
    a = eng.encode_corpus(_to_dev(text, dev), _to_dev(offs, dev), _cabi.RULE_SPM_LLAMA)
    text2, offs2 = pack_documents(tgt)
    b = eng.encode_corpus(_to_dev(text2, dev), _to_dev(offs2, dev), _cabi.RULE_SPM_LLAMA)
    ids_a, oa = a.ids.cpu().numpy(), a.doc_tok_offs.cpu().numpy()
    ids_b, ob = b.ids.cpu().numpy(), b.doc_tok_offs.cpu().numpy()
    rows_a = [ids_a[oa[d]:oa[d + 1]].tolist() for d in range(len(docs))]
    rows_ab = [rows_a[d] + ids_b[ob[d]:ob[d + 1]].tolist() for d in range(len(docs))]
    pad = 0
    for rows, labels, left in ((rows_a, None, False), (rows_a, None, True), (rows_ab, b, False)):
        for doc_begin, n_rows in ((0, len(docs)), (3, 7)):
            inp, mask, lens = eng.pad_batch(a, pad, doc_begin=doc_begin, n_rows=n_rows, pad_left=left, labels=labels)
            sel = rows[doc_begin:doc_begin + n_rows]
            L = max(len(r) for r in sel)
            exp = np.full((n_rows, L), pad, np.int64)
            em = np.zeros((n_rows, L), np.int64)
            for r, row in enumerate(sel):
                if left:
                    exp[r, L - len(row):] = row
                    em[r, L - len(row):] = 1
                else:
                    exp[r, :len(row)] = row
                    em[r, :len(row)] = 1
            assert inp.dtype == torch.int64 and np.array_equal(inp.cpu().numpy(), exp)
            assert np.array_equal(mask.cpu().numpy(), em) and lens.cpu().tolist() == [len(r) for r in sel]
    # truncation to a fixed row length
    inp, mask, lens = eng.pad_batch(a, pad, row_len=16)
    assert inp.shape == (len(docs), 16) and all(inp[r, :16].tolist() == (rows_a[r] + [pad] * 16)[:16] for r in range(len(docs)))


def test_corpus_statistics_driver(dev):
    """Row f2: dptok.stats.probe_dp_vs_default == a literal restatement of main_analyze_s2orc.py:251-307 on the oracle."""
    from dptok import assets, stats, synth
    from oracle import adapters
    tok = assets.load_hf("llama2_2k")
    docs = synth.sample_text(60_000, seed=11)[:60] + ["the weather", "naïve 日本 12 %", "a  b"]
    out = stats.probe_dp_vs_default(tok, docs, domains=["d%d" % (k % 3) for k in range(len(docs))])
    inv = {i: t for t, i in tok.get_vocab().items()}
    n_improved = 0
    for k, a in enumerate(docs):
        dp = adapters.llama_encode(tok, a)
        default = tok.encode(a)
        assert out["dp_length"][k] == len(dp) and out["default_length"][k] == len(default)
        imp, worse = [], []
        if len(dp) < len(default):
            n_improved += 1
            for token in adapters.llama_words(tok, a):
                d2, f2 = adapters.llama_encode(tok, token), tok.encode(token)
                if len(d2) < len(f2):
                    imp.append([inv[i] for i in d2])
                    worse.append([inv[i] for i in f2])
        assert out["improved_tokens"][k] == imp and out["worse_tokens"][k] == worse
    assert out["total_improved"] == n_improved and out["total"] == len(docs)
    assert all(dl <= fl for dl, fl in zip(out["dp_length"], out["default_length"]))
