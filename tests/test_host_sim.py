"""The exact per-word / per-character code the CUDA kernels run (csrc/dpt_dp_core.h, dpt_rules.h, the vocab
compiler), compiled for the host by tests/host_sim and checked against the oracle and the golden vectors.
CPU only; this is how kernel logic is validated before GPU time is spent.  The simulation library is test
infrastructure: the product never loads it."""
import ctypes
import random

import numpy as np

from conftest import load_golden
from helpers import make_sim_vocab, sim_word, vocab_bytes
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
