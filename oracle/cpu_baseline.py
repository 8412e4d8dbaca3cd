"""The reference's CPU path timed on host cores.  TEST/BENCH INFRASTRUCTURE ONLY (bench.py cpu_baseline leg).

The reference is pure Python and cannot travel to the GPU box, so this is the ORACLE PORT (kind "port"):
per document  default-tokenizer word split (tokenizer_utils.py:24-31)  ->  per word the literal algorithm of
dp_tokenize.py:24-84 (O(n^2) substring joins + set probes, exhaustive enumeration of tied optima, first-longest
selection)  ->  ids, i.e. ``oracle.dp_oracle.enumerate_shortest`` + ``pick_longest_token`` - the same
algorithmic structure and cost profile as the reference - under multiprocessing over documents like the
reference would be run on a multi-core host.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time

_STATE = {}


def _init(asset_name, root):
    for p in (root, os.path.join(root, "dp-tokenization_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")
    from dptok import assets
    from oracle import dp_oracle
    tok = assets.load_hf(asset_name)
    t2i = tok.get_vocab()
    _STATE.update(tok=tok, t2i=t2i, vocab=set(t2i), inv={i: t for t, i in t2i.items()}, dp=dp_oracle)


def _encode_doc(doc: str):
    tok, t2i, vocab, inv, dp = (_STATE[k] for k in ("tok", "t2i", "vocab", "inv", "dp"))
    toks = [inv[i] for i in tok.encode(doc)]
    words = []
    for k, t in enumerate(toks):
        if k == 0 or t.startswith("▁"):
            words.append(t)
        else:
            words[-1] += t
    n = 0
    for w in words:
        options, _ = dp.enumerate_shortest(w, vocab)
        n += len(dp.pick_longest_token(options))
    return len(doc.encode("utf-8")), n


def run(asset_name: str, docs, budget_s: float = 15.0, procs: int | None = None):
    """Tokenize ``docs`` (list of str) for about ``budget_s`` seconds on ``procs`` processes.
    -> dict(bytes_per_s, tokens_per_s, cores, bytes, tokens, docs, seconds)"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    done_b = done_t = done_d = 0
    with ctx.Pool(procs, initializer=_init, initargs=(asset_name, root)) as pool:
        list(pool.imap_unordered(_encode_doc, docs[:procs], chunksize=1))   # warm: imports, tokenizer load
        t0 = time.perf_counter()
        it = pool.imap_unordered(_encode_doc, docs, chunksize=4)
        for nb, nt in it:
            done_b += nb
            done_t += nt
            done_d += 1
            if time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
        pool.terminate()
    return dict(bytes_per_s=done_b / dt, tokens_per_s=done_t / dt, cores=procs, bytes=done_b, tokens=done_t,
                docs=done_d, seconds=dt)
