"""On-device corpus generator (dpt_synth_corpus; measurement support for BASELINE.json configs[3] / configs[4]):
determinism, range generation == slice of the global corpus, valid UTF-8, and the tokenization of generated text
bit-exact against the oracle on a document sample.  Run with ``-m gpu`` on a B200."""
import random

import numpy as np
import pytest
import torch

from helpers import pack, vocab_bytes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.cuda.current_device()


def test_device_generator_properties_and_parity(dev):
    from dptok import _cabi, assets, synth, synth_device
    from dptok.engine import Engine
    from dptok.vocab import CompiledVocab
    from oracle.c_oracle import COracle
    en = synth_device.DeviceLexicon(synth.make_lexicon(20_000, seed=0), dev)
    ar = synth_device.DeviceLexicon(synth.make_arabic_lexicon(20_000, seed=0), dev)
    assert en.ascii and not ar.ascii
    n_docs = 600
    text, offs = synth_device.generate(en, n_docs, seed=7, words_per_doc=(150, 900), device=dev, lex_b=ar, frac_b=0.4)
    text2, offs2 = synth_device.generate(en, n_docs, seed=7, words_per_doc=(150, 900), device=dev, lex_b=ar, frac_b=0.4)
    assert torch.equal(text, text2) and torch.equal(offs, offs2)          # deterministic
    lo, hi = 123, 456                                                     # a document range of the same corpus
    part, poffs = synth_device.generate(en, hi - lo, seed=7, words_per_doc=(150, 900), device=dev, lex_b=ar, frac_b=0.4,
                                        doc_base=lo)
    assert torch.equal(part, text[int(offs[lo]):int(offs[hi])]) and torch.equal(poffs, offs[lo:hi + 1] - offs[lo])
    other, _ = synth_device.generate(en, n_docs, seed=8, words_per_doc=(150, 900), device=dev, lex_b=ar, frac_b=0.4)
    assert other.numel() != text.numel() or not torch.equal(other, text)  # the seed matters
    raw = text.cpu().numpy().tobytes()
    # the host port of the generator (bench.py --impl reference runs without a GPU) gives the same documents
    h_en = synth_device.HostLexicon(synth.make_lexicon(20_000, seed=0))
    h_ar = synth_device.HostLexicon(synth.make_arabic_lexicon(20_000, seed=0))
    h_docs = synth_device.generate_host(h_en, 12, seed=7, words_per_doc=(150, 900), lex_b=h_ar, frac_b=0.4, doc_base=3)
    ho = offs.cpu().numpy()
    assert h_docs == [raw[ho[d]:ho[d + 1]] for d in range(3, 15)]
    sfx, so = synth_device.generate(en, 4, seed=9, words_per_doc=(50, 60), device=dev, suffix_prob=0.5)
    sraw, so = sfx.cpu().numpy().tobytes(), so.cpu().numpy()
    assert synth_device.generate_host(h_en, 4, seed=9, words_per_doc=(50, 60), suffix_prob=0.5) == \
        [sraw[so[d]:so[d + 1]] for d in range(4)]
    h_offs = offs.cpu().numpy()
    docs = [raw[h_offs[d]:h_offs[d + 1]].decode("utf-8") for d in range(n_docs)]   # valid UTF-8, document by document
    n_ar = sum(1 for d in docs if any("؀" <= ch <= "ۿ" for ch in d[:200]))
    assert 0.25 * n_docs < n_ar < 0.55 * n_docs
    assert all(d and not d.startswith(" ") and not d.endswith(" ") and "  " not in d for d in docs)
    assert any(d[0].isupper() for d in docs) and all(d.endswith(".") for d in docs)

    # tokenization of the generated text, Llama-3-shaped byte-level vocabulary and split rule, against tokenizers + C oracle
    tok = assets.load_tokenizer("llama3_128k")
    t2i = {t: k for k, t in enumerate(assets.load_spec("llama3_128k")["model"]["vocab"])}
    eng = Engine(CompiledVocab.from_token_map(t2i, "bytelevel"), dev)
    res = eng.encode_corpus(text, offs, _cabi.RULE_LLAMA3)
    assert int(res.counters[3]) == 0 and int(res.counters[0]) == len(raw)
    from dptok.vocab import bytelevel_to_bytes
    sample = sorted(random.Random(1).sample(range(n_docs), 40))
    words = []
    for d in sample:
        words += [bytelevel_to_bytes(p) for p, _ in tok.pre_tokenizer.pre_tokenize_str(docs[d])]
    wtext, woffs = pack(words)
    o_ids, o_lens, o_untok = COracle(vocab_bytes(t2i, "bytelevel"), 0).encode_words(wtext, woffs)
    ids = res.ids.cpu().numpy()
    dto = res.doc_tok_offs.cpu().numpy()
    got = np.concatenate([ids[dto[d]:dto[d + 1]] for d in sample])
    assert np.array_equal(got, o_ids)
    ok = eng.roundtrip_ok(res, text, offs, skip_bos=False)
    assert bool(ok.all())
