"""Generate golden vectors by running the UNMODIFIED reference (``/root/reference``) in this container.

    python tests/golden/make_golden.py

Writes (gzip JSON, committed):
  c0_toy.json.gz        config C0: union of the reference tests' toy vocabularies
                        (tests/test_tokenization_algorithms.py:15,24,28,33,44) + 10,000 seeded words;
                        per word the FULL return of compute_shortest_tokenizations (list kept verbatim when
                        it has <= 24 entries, sha1 of its JSON otherwise), obtain_longest_token, and
                        min_tokens_for_string (inspect_tokenizer.py:77-86).
  known_answers.json    the reference's own assertions (:14-48) replayed, with the survey's probes.
  llama_adapter.json.gz dp_tokenize_llama (tokenizer_utils.py:52-96, 5th call argument dropped)
                        on the committed llama2_2k stand-in tokenizer: words + ids per text.
  bytelevel_adapter.json.gz  dp_tokenize_bloom (tokenizer_utils.py:98-181) on gpt2_3k / bloom_8k.
The reference cannot travel to the GPU box, these files can.
"""
from __future__ import annotations

import gzip
import hashlib
import json
import os
import random
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dp-tokenization_b200"))

from oracle import ref_harness  # noqa: E402

TOY_VOCABS = [
    ["a", "ab", "abc", "d", "cd", "bcd", "b", "c"],
    ["un", "desirable", "und", "es", "ira", "ble", "u", "n"],
    ["un", "desirable", "und", "es", "ira", "ble", "ish"],
    ["un", "desirable"] + ["und", "esirable"] + list("undesirable"),
    list("desireableish") + ["desire", "able", "ish"] + ["des", "ireable", "ish"],
]

REFERENCE_TEST_STRINGS = [
    "the weather",
    "OptimalLengthTokenization",
    "midafternoon",
    'General thermodynamic relations for the work of polydisperse micelle formation in the model of ideal solution '
    'of molecular aggregates in nonionic surfactant solution and the model of "dressed micelles" in ionic solution '
    'have been considered.',
    "PURPOSE\nPatients with cancer frequently use herbal supplements and concomitant medications along with "
    "antineoplastic agents. These patients are at high risk of herb-drug interactions (HDIs) and drug-drug "
    "interactions (DDIs).\n\n\nMETHODS\nPatients starting a new anticancer therapy were asked to complete a "
    "questionnaire (17.4%; 95% CI, 11.3% to 23.5%), vis à vis doctors.",
    "Im klinischen Alltag ist das Pro-re-nata-Regime (PRN) weniger gut planbar, bedarf häufigerer Termine. "
    "„Best-Practice“-Empfehlungen für die Praxis: a) Vorbereitungsphase mit Ändern der Terminorganisation.",
    "Ġ",
]


def dump(name, obj, gz=True):
    data = json.dumps(obj, ensure_ascii=False, separators=(",", ":")).encode("utf-8")
    path = os.path.join(HERE, name)
    if gz:
        with gzip.GzipFile(path, "wb", mtime=0) as f:
            f.write(data)
    else:
        with open(path, "wb") as f:
            f.write(json.dumps(obj, ensure_ascii=False, indent=1).encode("utf-8"))
    print(name, os.path.getsize(path))


def c0_words(seed=0, n=10_000):
    """10,000 words: random vocab pieces and random alphabet chars, length 1-24."""
    rng = random.Random(seed)
    vocab = sorted({t for v in TOY_VOCABS for t in v})
    alphabet = sorted({c for t in vocab for c in t} | set("xyz"))
    words = []
    while len(words) < n:
        target = rng.randint(1, 24)
        w = ""
        while len(w) < target:
            w += rng.choice(vocab) if rng.random() < 0.7 else rng.choice(alphabet)
        words.append(w[:24])
    return vocab, words


def make_c0(dp):
    min_tokens = ref_harness.load_min_tokens()
    vocab, words = c0_words()
    vset = set(vocab)
    rows = []
    for w in words:
        all_opt, length = dp.compute_shortest_tokenizations(w, vset, False, "")
        sel = dp.obtain_longest_token(all_opt) if all_opt else None
        blob = json.dumps(all_opt, ensure_ascii=False, separators=(",", ":"))
        mt = min_tokens(w, vset)
        rows.append({
            "w": w, "len": length, "n": len(all_opt), "sel": sel,
            "all": all_opt if len(all_opt) <= 24 else None,
            "sha1": hashlib.sha1(blob.encode("utf-8")).hexdigest(),
            "min_tokens": None if mt == float("inf") else mt,
        })
    dump("c0_toy.json.gz", {"vocab": vocab, "rows": rows})


def make_known(dp):
    min_tokens = ref_harness.load_min_tokens()
    out = {"min_tokens": [], "shortest": [], "phantom": []}
    v1 = {"a", "ab", "abc", "d", "cd", "bcd", "b", "c"}
    for s, v in [("abcd", v1), ("adcbdab", v1), ("abdcd", v1),
                 ("undesirable", {"un", "desirable", "und", "es", "ira", "ble", "u", "n"}),
                 ("undesirableish", {"un", "desirable", "und", "es", "ira", "ble", "ish"})]:
        out["min_tokens"].append({"s": s, "vocab": sorted(v), "expect": min_tokens(s, v)})
    for s, v in [("undesirable", TOY_VOCABS[3]), ("desireableish", TOY_VOCABS[4]), ("abcd", TOY_VOCABS[0])]:
        all_opt, length = dp.compute_shortest_tokenizations(s, v, False, "")
        out["shortest"].append({"s": s, "vocab": v, "all": all_opt, "len": length,
                                "sel": dp.obtain_longest_token(all_opt)})
    for s, v in [("qrsTUV", ["qr", "s", "T", "U", "V", "rsTUV"]), ("xyz", ["x", "y"])]:
        all_opt, length = dp.compute_shortest_tokenizations(s, set(v), False, "")
        out["phantom"].append({"s": s, "vocab": v, "all": all_opt, "len": length})
    # the dead flag (dp_tokenize.py:24-25)
    all_opt, length = dp.compute_shortest_tokenizations("the", ["##the", "t", "h", "e"], True, "#")
    out["strip_marker"] = {"s": "the", "vocab": ["##the", "t", "h", "e"], "marker": "#", "all": all_opt, "len": length}
    dump("known_answers.json", out, gz=False)


def sample_texts(flavour, n_docs, seed):
    from dptok import synth
    docs = synth.sample_text(60_000, seed=seed, flavour=flavour)[:n_docs]
    return docs


def make_llama(tu):
    from dptok import assets
    tok = assets.load_hf("llama2_2k")
    enc, dec = tu.dp_tokenize_llama(tok)
    vocab = tu.bidict(tok.get_vocab())
    split = tu.pretokenize_with_llama(tok, vocab)
    texts = list(REFERENCE_TEST_STRINGS[:5]) + sample_texts("en", 24, 7)
    texts += ["a  b", " lead", "trail ", "tab\tsep", "日本語 text", "x" * 70 + " " + "ab" * 40, "é", "a",
              "multiple   spaces   here", "naïve café — déjà vu", "semi;colon,comma.(paren)"]
    rows = []
    for t in texts:
        ids = enc(t)
        rows.append({"text": t, "words": split(t), "ids": ids, "decoded": dec(ids),
                     "default_len": len(tok.encode(t))})
    # the 'raw' option (tokenizer_utils.py:33-50,62-63) on strings it can tokenize
    enc_raw, _ = tu.dp_tokenize_llama(tok, "raw")
    raw_rows = []
    for t in ["the weather", "midafternoon", "plai gout trot"]:
        try:
            raw_rows.append({"text": t, "ids": enc_raw(t)})
        except Exception as e:  # noqa: BLE001
            raw_rows.append({"text": t, "error": type(e).__name__})
    dump("llama_adapter.json.gz", {"tokenizer": "llama2_2k", "rows": rows, "raw": raw_rows})


def make_bytelevel(tu):
    from dptok import assets
    out = {}
    for name in ("gpt2_3k", "bloom_8k"):
        tok = assets.load_hf(name)
        spec = assets.load_spec(name)
        # the reference needs legacy "a b" merge strings and the hard-coded snapshot path
        spec["model"]["merges"] = [m if isinstance(m, str) else " ".join(m) for m in spec["model"]["merges"]]
        with tempfile.TemporaryDirectory() as cache:
            d = os.path.join(cache, "models--bigscience--bloom-3b", "snapshots", "52bc5b43010b4844513826b8be3f78c7344c37d7")
            os.makedirs(d)
            with open(os.path.join(d, "tokenizer.json"), "w") as f:
                json.dump(spec, f)
            enc, dec = tu.dp_tokenize_bloom(tok, cache)
            texts = list(REFERENCE_TEST_STRINGS) + sample_texts("en", 10, 11) + sample_texts("de", 8, 12) + \
                sample_texts("ar", 6, 13)
            texts += ["Hello world's  12345 tests.\n\nNew", "a\tb\r\nc", "don't can't we'll I'm", "   ", "x"]
            rows = []
            for t in texts:
                ids = enc(t)
                pieces = [p[0] for p in tok._tokenizer.pre_tokenizer.pre_tokenize_str(t)]
                rows.append({"text": t, "pieces": pieces, "ids": ids, "decoded": dec(ids),
                             "default_len": len(tok.encode(t))})
            out[name] = rows
    dump("bytelevel_adapter.json.gz", out)


def main():
    dp, tu = ref_harness.load()
    make_known(dp)
    make_c0(dp)
    make_llama(tu)
    make_bytelevel(tu)


if __name__ == "__main__":
    main()
