"""The cooperative DP kernel's arithmetic (tests/coop_model.py, a lane-level transliteration of
dp-tokenization_b200/csrc/dpt_dp_coop.cuh) against the oracle's device-contract DP on random words."""
import random

from oracle import dp_oracle
import coop_model

MARK = "▁".encode("utf-8")


def _random_case(rng, spm):
    alpha = [bytes([c]) for c in b"abcde"] + (["é".encode(), "ß".encode(), "د".encode(), "€".encode()] if spm else
                                               [b"\xc3", b"\xa9", b"\x81"])
    n_chars = rng.randint(1, 9)
    chars = [rng.choice(alpha) for _ in range(n_chars)]
    body = b"".join(chars)
    vocab = {}
    singles = set(chars) if rng.random() < 0.7 else set(rng.sample(chars, max(1, len(chars) // 2)))
    for c in singles:
        vocab.setdefault(c, len(vocab))
    if spm and rng.random() < 0.8:
        vocab.setdefault(MARK, len(vocab))
    for _ in range(rng.randint(0, 10)):
        a = rng.randint(0, n_chars - 1)
        b = rng.randint(a + 1, n_chars)
        tokb = b"".join(chars[a:b])
        if spm and a == 0 and rng.random() < 0.6:
            tokb = MARK + tokb
        vocab.setdefault(tokb, len(vocab))
    return body, vocab


def _check(spm, seed, n_cases):
    rng = random.Random(seed)
    solved = deferred = untok = 0
    for _ in range(n_cases):
        body, vocab = _random_case(rng, spm)
        if len(body) + (1 if spm else 0) > 31:
            continue
        got = coop_model.solve(body, vocab, spm)
        data = (MARK + body) if spm else body
        if got is None:
            deferred += 1
            assert spm
            continue
        want = dp_oracle.dp_bytes(data, vocab, 1 if spm else 0)
        ids, wl, un = got
        assert un == want["untokenizable"], (body, vocab)
        assert wl == want["word_len"], (body, vocab, wl, want)
        assert ids == want["ids"], (body, vocab, ids, want)
        solved += 1
        untok += un
    assert solved > n_cases // 3 and untok > 0
    return solved, deferred, untok


def test_coop_model_bytes_mode():
    _check(False, 1, 6000)


def test_coop_model_codepoint_mode():
    _check(True, 2, 6000)
