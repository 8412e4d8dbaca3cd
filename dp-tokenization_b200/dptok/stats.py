"""Corpus statistics driver: DP vs default tokenization over a list of documents.

Restates ``step_probe_eval_dataset`` of the reference (main_analyze_s2orc.py:251-307): per document ``dp_length`` and
``default_length``, the number of improved documents, and - only for improved documents - the per-pre-token lists of
DP tokens vs default tokens where the DP is shorter (:278-290).  The DP side is one GPU pass over all documents plus one
over the pre-tokens of the improved ones; the default side is the tokenizer's own ``encode`` (host), as in the reference.
The result has the reference's JSON columns (:299-307).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence


def probe_dp_vs_default(tokenizer, abstracts: Sequence[str], domains: Optional[Sequence[str]] = None, dp_tokenize=None) -> Dict:
    from packages.tokenizer_utils import _BiMap, dp_tokenize_llama, pretokenize_with_llama
    if dp_tokenize is None:
        dp_tokenize, _ = dp_tokenize_llama(tokenizer)
    vocab = _BiMap(tokenizer.get_vocab())
    inv = vocab.inverse
    pretokenize_func = pretokenize_with_llama(tokenizer, vocab)
    abstracts = list(abstracts)
    dp_ids = dp_tokenize.batch(abstracts)                                  # one device pass (main_analyze_s2orc.py:271)
    dp_lengths = [len(x) for x in dp_ids]
    default_lengths = [len(tokenizer.encode(a)) for a in abstracts]       # :272
    improved = [k for k in range(len(abstracts)) if dp_lengths[k] < default_lengths[k]]
    # :277-290 - the pre-tokens of every improved document, each tokenized both ways as a string of its own
    pretoks: List[List[str]] = [pretokenize_func(abstracts[k]) for k in improved]
    flat = [t for doc in pretoks for t in doc]
    flat_dp = dp_tokenize.batch(flat) if flat else []
    improved_tokens: List[List[List[str]]] = [[] for _ in abstracts]
    worse_tokens: List[List[List[str]]] = [[] for _ in abstracts]
    pos = 0
    for k, doc in zip(improved, pretoks):
        for token in doc:
            dp_t = flat_dp[pos]
            pos += 1
            default_t = tokenizer.encode(token)
            if len(dp_t) < len(default_t):
                improved_tokens[k].append([inv[i] for i in dp_t])
                worse_tokens[k].append([inv[i] for i in default_t])
    return {"abstract": abstracts, "domain": list(domains) if domains is not None else [None] * len(abstracts),
            "dp_length": dp_lengths, "default_length": default_lengths, "improved_tokens": improved_tokens,
            "worse_tokens": worse_tokens, "total_improved": len(improved), "total": len(abstracts)}
