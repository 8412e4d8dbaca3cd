"""Lane-level model of the cooperative DP kernel (dp-tokenization_b200/csrc/dpt_dp_coop.cuh) in plain Python.

TEST INFRASTRUCTURE.  The kernel keeps one word per tile of T lanes: lane t owns byte t (SPM rule: lane 0 owns the
word-initial U+2581 as ONE lane), records the ends of the vocabulary entries that start at t as a bit mask E_t, then
runs the forward relaxation with one min-reduction per position and selects the predecessors backwards with one
ballot + clz per token.  This model executes exactly those steps on Python ints (the same 32-bit keys, the same
masks), so the CPU suite can check the kernel's arithmetic against the oracle before GPU time is spent; the CUDA
source is a transliteration of ``solve`` below.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

NONE = 0xFFFFFFFF
MARK = "▁".encode("utf-8")


def _extend(kj: int, cl: int) -> int:
    lowj, lowe = kj & 0xFFFF, 0xFFFF - cl
    return ((kj & 0xFFFF0000) + (1 << 17) + min(lowj, lowe)) & 0xFFFFFFFF


def solve(body: bytes, vocab: Dict[bytes, int], spm: bool) -> Optional[Tuple[List[int], int, bool]]:
    """-> (ids, word_len, untokenizable), or None when the kernel defers the word (SPM: out-of-vocabulary character).

    ``body`` = the raw bytes of the word without the marker; SPM words are marker + body."""
    m = 1 if spm else 0
    n = len(body) + m
    assert 1 <= n <= 31
    # lane t: the byte string it owns
    own = [MARK if (spm and t == 0) else bytes([body[t - m]]) for t in range(n)]
    isstart = [True if t <= m else (not spm or (body[t - m] & 0xC0) != 0x80) for t in range(n)]
    Bm = sum(1 << t for t in range(n) if isstart[t]) | (1 << n)
    # walks: E_t = ends of vocabulary entries that start at t (the kernel finds them by walking the trie)
    E = [0] * n
    for t in range(n):
        if not isstart[t]:
            continue
        s = b""
        for i in range(t + 1, n + 1):
            s += own[i - 1]
            if s in vocab:
                E[t] |= 1 << i
        E[t] &= Bm
    if spm:
        for t in range(1, n):
            if isstart[t]:
                above = Bm >> (t + 1)
                nb = t + 1 + ((above & -above).bit_length() - 1)
                if not (E[t] >> nb) & 1:
                    return None  # a character that is no vocabulary entry: "<0xHH>" spelling, thread-per-word kernel
    popc = lambda x: bin(x).count("1")
    U = [popc(Bm & ((1 << p) - 1)) for p in range(n + 1)]
    best = [0xFFFF] + [NONE] * (n - 1)
    bestN = NONE
    for i in range(1, n + 1):
        cands = [(_extend(best[t], U[i] - U[t]) if (E[t] >> i) & 1 else NONE) for t in range(n)]
        kmin = min(cands)
        ph = ((U[i] << 17) | 0x1FFFF) if (Bm >> i) & 1 else NONE
        nb = min(kmin, ph)
        if i < n:
            best[i] = nb
        if i == n:
            bestN = nb
    wl = bestN >> 17
    reach = not (bestN & 0x10000)
    target = 0xFFFF - (bestN & 0xFFFF)
    if not reach:
        return [], wl, True
    ids: List[int] = [0] * wl
    i, o, got, cur = n, wl, False, bestN
    while i > 0:
        cands = [(_extend(best[t], U[i] - U[t]) if (E[t] >> i) & 1 else NONE) for t in range(n)]
        sel = [c != NONE and ((c >> 16) == (cur >> 16) if got else c == cur) for c in cands]
        mm = sum(1 << t for t in range(n) if sel[t])
        assert mm, "no predecessor on a reachable path"
        j = mm.bit_length() - 1
        o -= 1
        ids[o] = vocab[b"".join(own[j:i])]
        if not got and U[i] - U[j] == target:
            got = True
        cur = best[j]
        i = j
    assert o == 0
    return ids, wl, False
