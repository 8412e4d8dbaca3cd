"""GPU parity of the byte-level split rules (GPT-2, Llama-3, BLOOM) running inside the corpus pipeline, through the C ABI:
pieces must equal the installed `tokenizers` pre-tokenizer's (tokenizer_utils.py:157-159), ids/lengths the oracle's."""
import random

import numpy as np
import pytest
import torch

from helpers import bytelevel_table, pack, vocab_bytes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    return 0


def _to_dev(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a).copy()).to(dev)


def _expected(tok, vb, docs):
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


def _engine(name, dev):
    from dptok import assets
    from dptok.engine import Engine
    from dptok.vocab import CompiledVocab
    tok = assets.load_tokenizer(name)
    v2i = {t: k for k, t in enumerate(assets.load_spec(name)["model"]["vocab"])}
    return tok, v2i, vocab_bytes(v2i, "bytelevel"), Engine(CompiledVocab.from_token_map(v2i, "bytelevel"), dev)


def _check(eng, tok, vb, rule, docs, dev):
    from dptok.engine import pack_documents
    text, offs = pack_documents(docs)
    res = eng.encode_corpus(_to_dev(text, dev), _to_dev(offs, dev), rule)
    words, o_ids, o_lens, o_untok, o_dto = _expected(tok, vb, docs)
    assert res.n_words == len(words)
    assert np.array_equal(res.word_lens.cpu().numpy(), o_lens)
    assert np.array_equal(res.ids.cpu().numpy(), o_ids)
    assert np.array_equal(res.doc_tok_offs.cpu().numpy(), o_dto)
    assert res.counters.cpu().tolist() == [sum(len(d) for d in docs), len(words), len(o_ids), int(o_untok.sum())]
    return res


@pytest.mark.parametrize("name,rule_name", [("gpt2_50k", "RULE_GPT2"), ("llama3_128k", "RULE_LLAMA3"),
                                            ("bloom_8k", "RULE_BLOOM")])
def test_device_split_rules_vs_tokenizers(dev, name, rule_name):
    """configs[2]/[3]-shaped input: en/de sentence pairs (GPT-2 50k) and long mixed documents incl. Arabic with
    diacritics (Llama-3 128k), plus a soup of contractions, digit runs, newline runs, tabs, multi-byte whitespace and
    digits, and pieces far longer than a scan tile."""
    from dptok import _cabi, synth
    rule = getattr(_cabi, rule_name)
    tok, v2i, vb, eng = _engine(name, dev)
    text, offs = synth.gen_sentence_pairs(3_000_000, seed=0)
    raw = text.tobytes()
    res = _check(eng, tok, vb, rule, [raw[offs[k]:offs[k + 1]] for k in range(len(offs) - 1)], dev)
    assert int(res.counters[3]) == 0
    text, offs = synth.gen_documents(2_000_000, seed=2, flavour="ar", lexicon=synth.make_arabic_lexicon(20000, seed=2))
    raw = text.tobytes()
    _check(eng, tok, vb, rule, [raw[offs[k]:offs[k + 1]] for k in range(len(offs) - 1)], dev)
    rng = random.Random(7)
    soup = ["Hello", " ", "  ", "world", "'s", "'S", "'re", "'LL", "'", "''", "12345", "3", "٣٤", "é", "naïve", "日本",
            "\n", "\n\n", "\t", "\r\n", ".", ",", "!!", "(x)", "—", "…", " ", "　", "ſ", "'ſ", "x", "İ", "ǅ",
            "قُدَّام", "البيت", "%", "a1b2", " ", "+=", "'t", "'d", "'m", "'ve", "'VE"]
    soup += ["(", ")", "|", "?", "!", "[w]", "。", "，", "、", "।", "۔", "،", "؟", ". .", " .", " (", "a)b"]
    for trial in range(6):
        docs = ["".join(rng.choice(soup) for _ in range(rng.randint(1, 80))).encode()
                for _ in range(rng.choice([1, 50, 3000]))]
        _check(eng, tok, vb, rule, docs, dev)
    big = [("x" * 20000 + " y").encode(), ("7" * 9000 + "a").encode(), (" " * 6000 + "z\n" * 3000).encode(),
           ("." * 5000 + " " + "," * 4200 + " q").encode(), ("\u3002" * 1500 + " \u3000" * 700 + "w").encode()]
    _check(eng, tok, vb, rule, big, dev)


def test_bytelevel_adapter_batch_uses_device_rule(dev):
    """`dp_tokenize_bloom(...)[0].batch` recognises the tokenizer's split regex and runs it on the GPU; the result
    equals the per-string path that pre-tokenizes on the host like the reference (tokenizer_utils.py:161-174)."""
    from dptok import _cabi, assets, synth
    from packages.tokenizer_utils import dp_tokenize_bloom
    for name, rule in (("gpt2_3k", _cabi.RULE_GPT2), ("llama3_128k", _cabi.RULE_LLAMA3), ("bloom_8k", _cabi.RULE_BLOOM)):
        tok = assets.load_hf(name)
        enc, dec = dp_tokenize_bloom(tok, None)
        assert enc.device_rule == rule
        docs = synth.sample_text(60_000, seed=5) + ["it's 12345 o'clock\n\nNEW  line\there", "قُدَّام البيت ١٢٣"]
        out = enc.batch(docs)
        for d, ids in zip(docs[:40] + docs[-2:], out[:40] + out[-2:]):
            assert ids == enc(d) and dec(ids) == d


def test_full_size_properties_llama3_100mb(dev):
    """BASELINE.json configs[3]/[4] shape at 100 MB (Llama-3-shaped 128k byte-level vocabulary, 8-64 KB English documents
    followed by Arabic-script documents with diacritics, split regex on the device): size-independent properties - the
    device decode round-trips every document, counters are consistent, two runs are identical, the DP never uses more
    tokens than the default BPE - plus exact oracle parity (tokenizers' pieces + C oracle) on a document sample."""
    from dptok import _cabi, synth
    tok, v2i, vb, eng = _engine("llama3_128k", dev)
    t_en, o_en = synth.gen_documents(60_000_000, seed=0, words_per_doc=(1200, 9000))
    t_ar, o_ar = synth.gen_documents(40_000_000, seed=1, flavour="ar", words_per_doc=(800, 6000))
    text = np.concatenate([t_en, t_ar])
    doc_offs = np.concatenate([o_en, o_ar[1:] + o_en[-1]])
    n_docs = len(doc_offs) - 1
    d_text, d_offs = _to_dev(text, dev), _to_dev(doc_offs, dev)
    res = eng.encode_corpus(d_text, d_offs, _cabi.RULE_LLAMA3)
    c = res.counters.cpu().tolist()
    assert c[0] == len(text) and c[1] == res.n_words and c[2] == res.n_ids and c[3] == 0
    assert int(res.word_lens.sum()) == res.n_ids
    assert bool(eng.roundtrip_ok(res, d_text, d_offs, skip_bos=False).all())
    dto = res.doc_tok_offs.cpu().numpy()
    assert dto[0] == 0 and dto[-1] == res.n_ids and (np.diff(dto) > 0).all()
    res2 = eng.encode_corpus(d_text, d_offs, _cabi.RULE_LLAMA3)
    assert torch.equal(res.ids, res2.ids) and torch.equal(res.word_lens, res2.word_lens)
    ids = res.ids.cpu().numpy()
    raw = text.tobytes()
    sample = random.Random(1).sample(range(n_docs), 120)
    docs = [raw[doc_offs[d]:doc_offs[d + 1]] for d in sample]
    words, o_ids, o_lens, o_untok, o_dto = _expected(tok, vb, docs)
    for k, d in enumerate(sample):
        assert np.array_equal(ids[dto[d]:dto[d + 1]], o_ids[o_dto[k]:o_dto[k + 1]])
    for k, d in enumerate(sample[:15]):
        assert dto[d + 1] - dto[d] <= len(tok.encode(docs[k].decode()).ids)


def test_host_chunked_path_bytelevel(dev):
    """Engine.encode_corpus_host with the Llama-3 rule: ranges overlap on several streams and share one word table; the
    stitched result must equal the resident single-batch result for every chunking (repeated: the overlap is timing
    dependent)."""
    from dptok import _cabi, synth
    tok, v2i, vb, eng = _engine("llama3_128k", dev)
    t_en, o_en = synth.gen_documents(18_000_000, seed=3, words_per_doc=(1200, 9000))
    t_ar, o_ar = synth.gen_documents(12_000_000, seed=4, flavour="ar", words_per_doc=(800, 6000))
    text = np.concatenate([t_en, t_ar])
    doc_offs = np.concatenate([o_en, o_ar[1:] + o_en[-1]])
    res = eng.encode_corpus(_to_dev(text, dev), _to_dev(doc_offs, dev), _cabi.RULE_LLAMA3)
    ids, dto, ctr = res.ids.cpu().numpy(), res.doc_tok_offs.cpu().numpy(), res.counters.cpu().tolist()
    h_text = torch.from_numpy(text.copy()).pin_memory()
    for rep in range(3):
        for chunk, ns in ((1 << 20, 3), (3_000_000, 2), (8 << 20, 4)):
            hr = eng.encode_corpus_host(h_text, doc_offs, _cabi.RULE_LLAMA3, chunk_bytes=chunk, n_streams=ns,
                                        overlap=bool(rep & 1))
            assert hr.n_ids == res.n_ids and hr.counters.tolist() == ctr
            assert np.array_equal(hr.ids.numpy(), ids) and np.array_equal(hr.doc_tok_offs, dto)
