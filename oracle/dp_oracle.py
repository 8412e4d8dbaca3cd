"""CPU oracle for the shortest-tokenization DP.  TEST INFRASTRUCTURE ONLY.

This module is the checker for the CUDA path: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  Nothing under ``dp-tokenization_b200/`` does.

Parity status: PINNED.  ``tests/test_oracle.py`` checks every function here
against (a) the reference's own offline golden vectors
(/root/reference/tests/test_tokenization_algorithms.py:14-48) and (b) fixtures
under ``tests/golden/`` produced by running the unmodified reference
(``oracle/ref_harness.py`` + ``tests/golden/make_golden.py``).

Three restatements of /root/reference/packages/dp_tokenize.py live here:

* ``enumerate_shortest``  - the literal algorithm (forward DP with predecessor
  lists, exhaustive DFS backtrace), dp_tokenize.py:24-70.
* ``pick_longest_token``  - the tie-break selector, dp_tokenize.py:72-84.
* ``select_shortest``     - closed-form O(n*Lmax) equivalent of
  ``pick_longest_token(enumerate_shortest(...))`` that never enumerates; this
  is the formulation the CUDA kernels implement.
* ``dp_bytes``            - the same closed form over a UTF-8 byte string with
  explicit unit boundaries: the exact contract of the device DP.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence, Tuple

INF_LEN = 1 << 30


# --------------------------------------------------------------------------
# literal restatement (dp_tokenize.py:24-70)
# --------------------------------------------------------------------------
def forward_tables(units: Sequence[str], vocab) -> Tuple[List[int], List[List[int]]]:
    """len_dp[0..n] and predecessor lists P(1..n) (index 0 unused).

    dp_tokenize.py:27-47.  ``len_dp[i]`` starts at ``i`` (phantom value, :28),
    a strict improvement resets the predecessor list (:40-43), equality appends
    (:44-46); predecessors are therefore ascending.
    """
    n = len(units)
    best = list(range(n + 1))
    preds: List[List[int]] = [[] for _ in range(n + 1)]
    for i in range(1, n + 1):
        plist: List[int] = []
        for j in range(i):
            if "".join(units[j:i]) in vocab:
                cand = best[j] + 1
                if cand < best[i]:
                    best[i] = cand
                    plist = [j]
                elif cand == best[i]:
                    plist.append(j)
        preds[i] = plist
    return best, preds


def enumerate_shortest(units: Sequence[str], vocab, strip_marker: bool = False,
                       marker=None) -> Tuple[List[List[str]], int]:
    """All optimal segmentations in the reference's DFS order, and len_dp[n].

    dp_tokenize.py:6-70.  ``strip_marker`` reproduces the dead
    ``disregard_word_initial_marker`` flag (:24-25, a character-set lstrip).
    Raises IndexError on empty input like the reference (:49).
    """
    if strip_marker:
        vocab = {t.lstrip(marker) for t in vocab}
    n = len(units)
    if n == 0:
        raise IndexError("list index out of range")
    best, preds = forward_tables(units, vocab)
    done: List[List[str]] = []
    # stack entries: (start j of the token being added, its end, tokens to the right)
    stack = [(j, n, []) for j in preds[n]]
    while stack:
        j, end, right = stack.pop()          # largest j first (:58)
        toks = ["".join(units[j:end])] + right
        if 0 in preds[end]:                   # completion test of :63
            done.append(toks)
        else:
            for k in preds[j]:
                stack.append((k, j, toks))
    return done, best[n]


def pick_longest_token(tokenizations: List[List[str]]) -> List[str]:
    """dp_tokenize.py:72-84: first segmentation whose longest token is longest."""
    scores = [max(len(t) for t in toks) for toks in tokenizations]
    return tokenizations[scores.index(max(scores))]


def min_tokens(units: Sequence[str], vocab) -> float:
    """Length-only DP with infinity init, inspect_tokenizer.py:77-86."""
    n = len(units)
    dp = [float("inf")] * (n + 1)
    dp[0] = 0
    for i in range(1, n + 1):
        for j in range(i):
            if "".join(units[j:i]) in vocab and dp[j] + 1 < dp[i]:
                dp[i] = dp[j] + 1
    return dp[n]


# --------------------------------------------------------------------------
# closed form on unit sequences
# --------------------------------------------------------------------------
def select_shortest(units: Sequence[str], vocab, max_token_units: int | None = None):
    """(selected tokens or None, len_dp[n], number of optimal segmentations).

    Equivalent to ``pick_longest_token(enumerate_shortest(units, vocab)[0])``
    without enumerating.  ``None`` when the reference would return an empty
    list (word not tokenizable).
    """
    n = len(units)
    if n == 0:
        raise IndexError("list index out of range")
    w = max_token_units or n
    best = list(range(n + 1))
    reach = [False] * (n + 1)
    longest = [0] * (n + 1)
    count = [0] * (n + 1)
    reach[0] = True
    count[0] = 1
    preds: List[List[Tuple[int, int]]] = [[] for _ in range(n + 1)]
    for i in range(1, n + 1):
        cands = []
        for j in range(max(0, i - w), i):
            s = "".join(units[j:i])
            if s in vocab:
                cands.append((j, len(s)))
                if best[j] + 1 < best[i]:
                    best[i] = best[j] + 1
        pl = [(j, cl) for j, cl in cands if best[j] + 1 == best[i]]
        preds[i] = pl
        for j, cl in pl:
            if reach[j]:
                reach[i] = True
                longest[i] = max(longest[i], longest[j], cl)
                count[i] += count[j]
    if not reach[n]:
        return None, best[n], 0
    target = longest[n]
    got = False
    i = n
    out: List[str] = []
    while i > 0:
        pick = None
        for j, cl in preds[i]:            # ascending j; keep the largest that qualifies
            if reach[j] and (got or cl == target or longest[j] == target):
                pick = (j, cl)
        j, cl = pick
        out.append("".join(units[j:i]))
        got = got or cl == target
        i = j
    out.reverse()
    return out, best[n], count[n]


# --------------------------------------------------------------------------
# byte-level contract of the device DP
# --------------------------------------------------------------------------
def utf8_boundaries(data: bytes) -> List[int]:
    """Byte positions where a code point starts, plus len(data)."""
    return [k for k, b in enumerate(data) if (b & 0xC0) != 0x80] + [len(data)]


def dp_bytes(data: bytes, vocab: Dict[bytes, int], unit_mode: int,
             boundaries: Iterable[int] | None = None, max_token_bytes: int | None = None):
    """Device-contract DP over bytes.

    unit_mode 0: every byte is a unit, token length counted in bytes (byte-level
    BPE after the inverse bytes_to_unicode map).  unit_mode 1: units are code
    points (or the explicit ``boundaries``), token length counted in code points
    (SentencePiece path; dp_tokenize.py:82 uses len() of the token string).

    Returns dict(ids, tokens, word_len, untokenizable, n_optimal).
    ``word_len`` is the phantom-initialised len_dp[n] of dp_tokenize.py:28,70.
    """
    nb = len(data)
    if boundaries is None:
        bnd = list(range(nb + 1)) if unit_mode == 0 else utf8_boundaries(data)
    else:
        bnd = sorted(set(boundaries))
    assert bnd and bnd[0] == 0 and bnd[-1] == nb and nb > 0
    index_of = {p: k for k, p in enumerate(bnd)}
    if unit_mode == 0:
        cp_before = list(range(nb + 1))
    else:
        cp_before = [0] * (nb + 1)
        for k in range(nb):
            cp_before[k + 1] = cp_before[k] + (1 if (data[k] & 0xC0) != 0x80 else 0)
    lmax = max_token_bytes or max((len(t) for t in vocab), default=1)
    n = len(bnd) - 1
    best = list(range(n + 1))
    reach = [False] * (n + 1)
    longest = [0] * (n + 1)
    count = [0] * (n + 1)
    reach[0] = True
    count[0] = 1
    preds: List[List[Tuple[int, int]]] = [[] for _ in range(n + 1)]
    for ui in range(1, n + 1):
        pi = bnd[ui]
        cands = []
        for uj in range(ui - 1, -1, -1):
            pj = bnd[uj]
            if pi - pj > lmax:
                break
            if data[pj:pi] in vocab:
                cands.append((uj, cp_before[pi] - cp_before[pj]))
        cands.reverse()
        for uj, _ in cands:
            if best[uj] + 1 < best[ui]:
                best[ui] = best[uj] + 1
        pl = [(uj, cl) for uj, cl in cands if best[uj] + 1 == best[ui]]
        preds[ui] = pl
        for uj, cl in pl:
            if reach[uj]:
                reach[ui] = True
                longest[ui] = max(longest[ui], longest[uj], cl)
                count[ui] += count[uj]
    if not reach[n]:
        return dict(ids=[], tokens=[], word_len=best[n], untokenizable=True, n_optimal=0)
    target = longest[n]
    got = False
    ui = n
    toks: List[bytes] = []
    while ui > 0:
        pick = None
        for uj, cl in preds[ui]:
            if reach[uj] and (got or cl == target or longest[uj] == target):
                pick = (uj, cl)
        uj, cl = pick
        toks.append(data[bnd[uj]:bnd[ui]])
        got = got or cl == target
        ui = uj
    toks.reverse()
    return dict(ids=[vocab[t] for t in toks], tokens=toks, word_len=best[n],
                untokenizable=False, n_optimal=count[n])
