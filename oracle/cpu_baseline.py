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


class Runner:
    """One worker pool kept across several bounded runs (bench.py --impl reference: one run per step)."""

    def __init__(self, asset_name: str, procs: int | None = None):
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        self.procs = procs or os.cpu_count() or 1
        self.pool = mp.get_context("spawn").Pool(self.procs, initializer=_init, initargs=(asset_name, root))

    def warm(self, docs):
        list(self.pool.imap_unordered(_encode_doc, docs[:self.procs], chunksize=1))

    def run(self, docs, budget_s: float):
        """Tokenize documents from ``docs`` for about ``budget_s`` seconds.  Work is handed to the pool in small
        batches and each batch is waited for, so no task is left in flight when the step ends (imap over the whole
        list would keep running after a ``break`` and bleed into the next step's timing)."""
        done_b = done_t = done_d = 0
        batch = self.procs * 4
        pos = 0
        t0 = time.perf_counter()
        while pos < len(docs):
            for nb, nt in self.pool.map(_encode_doc, docs[pos:pos + batch], chunksize=4):
                done_b += nb
                done_t += nt
                done_d += 1
            pos += batch
            if time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
        return dict(bytes_per_s=done_b / dt, tokens_per_s=done_t / dt, cores=self.procs, bytes=done_b, tokens=done_t,
                    docs=done_d, seconds=dt)

    def close(self):
        self.pool.terminate()
