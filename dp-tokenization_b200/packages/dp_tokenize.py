"""B200 implementation behind the call surface of the reference's ``packages/dp_tokenize.py``.

Same names, argument meaning, return types and error behaviour:

* ``compute_shortest_tokenizations(units, vocabulary, disregard_word_initial_marker,
  word_initial_marker) -> (List[List[str]], int)``   (dp_tokenize.py:6-11,70)
* ``obtain_longest_token(tokenizations) -> List[str]``   (dp_tokenize.py:72-84)

The forward DP (len_dp and the predecessor lists of dp_tokenize.py:27-47) runs on the GPU
(``dpt_lattice_word``); the enumeration of every optimal segmentation in the reference's DFS order
(dp_tokenize.py:49-69) is output formatting of unbounded size and is unrolled on the host from that
lattice.  The throughput adapters in ``tokenizer_utils`` never enumerate: they use the fused
select-on-device kernels.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

from dptok.engine import Engine
from dptok.vocab import CompiledVocab

_ENGINE_CACHE: dict = {}
_ENGINE_CACHE_MAX = 8


def _engine_for(vocabulary) -> Engine:
    """Compile (or reuse) the vocabulary.  Keyed by content and confirmed by EQUALITY (a hash collision cannot return
    another vocabulary's engine; mutating a set between calls is safe).  Fast path without the O(|V|) rebuild: the same
    immutable ``frozenset`` object as before."""
    ident = _ENGINE_CACHE.get(("id", id(vocabulary))) if isinstance(vocabulary, frozenset) else None
    if ident is not None and ident[0] is vocabulary:
        return ident[1]
    try:
        fs = frozenset(vocabulary)
    except TypeError:
        fs = frozenset(str(t) for t in vocabulary)
    ent = _ENGINE_CACHE.get(hash(fs))
    if ent is not None and ent[0] == fs:
        eng = ent[1]
    else:
        eng = Engine(CompiledVocab.from_strings(vocabulary))
        while len(_ENGINE_CACHE) >= 2 * _ENGINE_CACHE_MAX:
            _ENGINE_CACHE.pop(next(iter(_ENGINE_CACHE)))
        _ENGINE_CACHE[hash(fs)] = (fs, eng)
    if isinstance(vocabulary, frozenset):
        _ENGINE_CACHE[("id", id(vocabulary))] = (vocabulary, eng)
    return eng


def compute_shortest_tokenizations(base_representation_s: Sequence[str], vocabulary, disregard_word_initial_marker: bool,
                                   word_initial_marker: str, *_ignored) -> Tuple[List[List[str]], int]:
    """All minimum-token segmentations of the unit sequence, in the reference's order, and len_dp[n].

    A fifth positional argument is accepted and ignored because tokenizer_utils.py:71 passes one.
    Raises IndexError on empty input like dp_tokenize.py:49.
    """
    if disregard_word_initial_marker:
        # dp_tokenize.py:24-25: str.lstrip with a character SET (None strips whitespace)
        vocabulary = {token.lstrip(word_initial_marker) for token in vocabulary}
    units = base_representation_s
    n = len(units)
    if n == 0:
        raise IndexError("list index out of range")
    if isinstance(units, str):
        data = units.encode("utf-8")
        starts = None
        bounds_units = None
    else:
        pieces = [u.encode("utf-8") for u in units]
        if any(len(p) == 0 for p in pieces):
            raise ValueError("empty units are not supported by the device lattice")
        data = b"".join(pieces)
        starts, pos = [], 0
        for p in pieces:
            starts.append(pos)
            pos += len(p)
        bounds_units = units
    len_dp, preds = _engine_for(vocabulary).lattice(data, starts)
    assert len(len_dp) == n + 1, "device lattice and host unit count disagree"

    def piece(j, i):
        return units[j:i] if bounds_units is None else "".join(units[j:i])

    # DFS of dp_tokenize.py:57-69: predecessors pushed ascending, popped from the end
    complete: List[List[str]] = []
    stack = [(j, n, None) for j in preds[n]]
    while stack:
        j, end, right = stack.pop()
        node = (piece(j, end), right)
        if 0 in preds[end]:
            toks, cur = [], node
            while cur is not None:
                toks.append(cur[0])
                cur = cur[1]
            complete.append(toks)
        else:
            for k in preds[j]:
                stack.append((k, j, node))
    return complete, len_dp[n]


def obtain_longest_token(tokenizations: List[List[str]]) -> List[str]:
    """First tokenization whose longest token is longest (dp_tokenize.py:82-84); ValueError on []."""
    best_len, best = -1, None
    if not tokenizations:
        raise ValueError("max() arg is an empty sequence")
    for toks in tokenizations:
        m = max(len(t) for t in toks)
        if m > best_len:
            best_len, best = m, toks
    return best


def min_tokens_for_string(s, vocabulary):
    """Fewest vocabulary tokens that spell ``s`` (``str`` or list of units); ``float('inf')`` when there is no
    segmentation - the reference's ``inspect_tokenizer.min_tokens_for_string`` (inspect_tokenizer.py:77-86, the definition
    that shadows :62-74; vectors tests/test_tokenization_algorithms.py:14-30).  Unlike ``len_dp[n]`` of
    ``compute_shortest_tokenizations`` it starts from infinity, not from the phantom ``len_dp[i] = i``.  The DP runs on the
    GPU (``dpt_min_tokens_word``)."""
    if len(s) == 0:
        return 0
    if isinstance(s, str):
        data, starts = s.encode("utf-8"), None
    else:
        pieces = [u.encode("utf-8") for u in s]
        if any(len(p) == 0 for p in pieces):
            raise ValueError("empty units are not supported by the device DP")
        data = b"".join(pieces)
        starts, pos = [], 0
        for p in pieces:
            starts.append(pos)
            pos += len(p)
    return _engine_for(vocabulary).min_tokens(data, starts)
