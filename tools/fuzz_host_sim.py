"""Time-boxed fuzz campaign over the kernels' source under host emulation (tests/host_sim), CPU only.

    python tools/fuzz_host_sim.py [minutes] [seed]

Each trial draws documents from a wide Unicode soup (all general categories the split regexes distinguish: letters of many
scripts, Nd / Nl / No digits, combining marks, every kind of white space, 4-byte characters, apostrophes, punctuation)
and checks the emulated corpus pipeline against the installed `tokenizers` pre-tokenizer + the C oracle's DP (byte-level
rules GPT-2 / Llama-3 / BLOOM) or against the tokenizer-driven split of the oracle adapter (SentencePiece rule with merge
table).  Prints one line per failing trial with the seed that reproduces it; exit status 1 if anything failed.  The
checkers are the ones of tests/test_host_sim.py."""
import ctypes
import os
import random
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dp-tokenization_b200"), os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402

import conftest  # noqa: E402
import test_host_sim as ths  # noqa: E402
from helpers import make_sim_vocab, vocab_bytes  # noqa: E402
from oracle import adapters  # noqa: E402

RANGES = [(0x20, 0x7E), (0xA0, 0xFF), (0x100, 0x17F), (0x370, 0x3FF), (0x400, 0x4FF), (0x5D0, 0x5EA), (0x600, 0x6FF),
          (0x900, 0x97F), (0xE00, 0xE7F), (0x2000, 0x206F), (0x2150, 0x218F), (0x2460, 0x24FF), (0x3000, 0x303F),
          (0x3040, 0x30FF), (0x4E00, 0x4FFF), (0xAC00, 0xACFF), (0xFF00, 0xFFEF), (0x1F600, 0x1F64F), (0x1D400, 0x1D4FF),
          (0x300, 0x36F), (0x1E00, 0x1EFF), (0x2C60, 0x2C7F)]
SPECIAL = [" ", " ", " ", "  ", "\n", "\n\n", "\r\n", "\r", "\t", "\v", "\f", "", " ", " ", " ", " ",
           " ", " ", "　", "﻿", "​", "‍", "'", "'s", "'S", "'t", "'re", "'RE", "'ve", "'m", "'ll", "'LL",
           "'d", "'ſ", "’s", "''", ".", ",", "!", "?", "…", "。", "，", "、", "।", "۔", "،",
           "(", ")", "|", "[", "]", "1", "12", "123", "1234", "12345678", "٣٤", "²", "½", "Ⅷ", "〇",
           "a", "ab", "Hello", "world", "é", "naïve", "قُدَّام", "日本",
           "\U0001f600", "x" * 40, "7" * 11, " " * 9]


def rand_char(rng):
    while True:
        lo, hi = rng.choice(RANGES)
        c = rng.randint(lo, hi)
        if 0xD800 <= c <= 0xDFFF:
            continue
        return chr(c)


def rand_doc(rng, max_parts):
    parts = []
    for _ in range(rng.randint(1, max_parts)):
        r = rng.random()
        if r < 0.55:
            parts.append(rng.choice(SPECIAL))
        elif r < 0.85:
            parts.append("".join(rand_char(rng) for _ in range(rng.randint(1, 6))))
        else:
            c = rand_char(rng)
            parts.append(c * rng.randint(1, 50))
    return "".join(parts)


def main():
    minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    from dptok import assets
    sim = conftest.build_host_sim()
    suites = []
    for name, rule in (("gpt2_3k", 2), ("llama3_128k", 3), ("bloom_8k", 4)):
        tok = assets.load_tokenizer(name)
        v2i = {t: k for k, t in enumerate(assets.load_spec(name)["model"]["vocab"])}
        vb = vocab_bytes(v2i, "bytelevel")
        suites.append(("bl", name, rule, tok, vb, make_sim_vocab(sim, vb, 0)))
    tok, t2i, mid = ths._tokenizer_with_marker_run_tokens()
    vb = vocab_bytes(t2i, "spm")
    h = make_sim_vocab(sim, vb, 1)
    m = np.asarray(mid, dtype=np.int32)
    left, right, merged = (np.ascontiguousarray(m[:, k]) for k in range(3))
    assert sim.sim_vocab_set_merges(ctypes.c_void_p(h), left.ctypes.data, right.ctypes.data, merged.ctypes.data, len(m)) == 0
    suites.append(("spm", "llama2_2k+runs", 1, tok, vb, h))
    t_end = time.time() + 60 * minutes
    trial, failures = 0, 0
    while time.time() < t_end:
        seed = seed0 + trial
        rng = random.Random(seed)
        kind, name, rule, tok, vb, h = suites[trial % len(suites)]
        docs = [rand_doc(rng, rng.choice([3, 20, 120])) for _ in range(rng.choice([1, 4, 40]))]
        docs = [d for d in docs if d]
        kw = dict(nthreads=rng.choice([1, 3, 8]), n_slots=rng.choice([0, 0, 64]), n_ranges=rng.choice([1, 1, 3]))
        try:
            if kind == "bl":
                ths._check_bytelevel(sim, h, tok, vb, rule, [d.encode() for d in docs], **kw)
            else:
                # SentencePiece: the adapter is the oracle per document (ids), specials excluded like the product does
                docs = [d.replace("<s>", "s").replace("</s>", "s").replace("<unk>", "u") for d in docs]
                r = ths._run_fused(sim, h, 1, [d.encode() for d in docs], **kw)
                assert r["nout"][2] <= r["nout"][3] and r["nout"][4] <= r["nout"][5] and r["nout"][6] <= r["nout"][7]
                for k, d in enumerate(docs):
                    if r["dfl"][k]:
                        continue  # left to the host split (runs longer than the device handles): allowed, not a mismatch
                    want = adapters.llama_encode(tok, d)
                    got = r["ids"][r["dto"][k]:r["dto"][k + 1]].tolist()
                    assert got == want, (d, adapters.llama_words(tok, d))
        except Exception as e:  # noqa: BLE001
            failures += 1
            msg = traceback.format_exc().strip().splitlines()[-1][:300]
            print(f"FAIL seed={seed} suite={name} kw={kw} n_docs={len(docs)}: {type(e).__name__} {msg}", flush=True)
        trial += 1
        if trial % 200 == 0:
            print(f"... {trial} trials, {failures} failures", flush=True)
    print(f"done: {trial} trials, {failures} failures, seeds {seed0}..{seed0 + trial - 1}")
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
