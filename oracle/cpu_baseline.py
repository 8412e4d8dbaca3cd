"""The reference's CPU path timed on host cores.  TEST/BENCH INFRASTRUCTURE ONLY (bench.py cpu_baseline / --impl reference).

Two kinds:
  * ``"reference"`` - the UNMODIFIED reference adapters (``dp_tokenize_llama(tok)`` / ``dp_tokenize_bloom(tok, cache)``,
    tokenizer_utils.py:52-96,98-181 over dp_tokenize.py:6-84) imported from ``oracle/_ref`` (staged by
    ``oracle/make_ref.py``; shims of ``oracle/ref_harness.py``), called per document exactly as
    main_analyze_s2orc.py:78 / main_biomed_translation.py:142-143 do;
  * ``"port"`` - when the staged copy is absent: the oracle port with the same algorithmic structure and cost profile
    (default-tokenizer word split, O(n^2) substring joins + set probes, exhaustive enumeration of tied optima,
    first-longest selection: ``oracle.dp_oracle.enumerate_shortest`` + ``pick_longest_token``).
Both run under multiprocessing over documents, like the reference would be run on a multi-core host.
"""
from __future__ import annotations

import json
import multiprocessing as mp
import os
import sys
import tempfile
import time

_STATE = {}
_HERE = os.path.dirname(os.path.abspath(__file__))
_BLOOM_SNAPSHOT = ("models--bigscience--bloom-3b", "snapshots", "52bc5b43010b4844513826b8be3f78c7344c37d7")


def kind() -> str:
    from . import make_ref
    return "reference" if make_ref.staged() else "port"


def _init(asset_name, family, root):
    for p in (root, os.path.join(root, "dp-tokenization_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")
    from dptok import assets
    from oracle import dp_oracle, make_ref
    tok = assets.load_hf(asset_name)
    _STATE.update(family=family, tok=tok, dp=dp_oracle, encode=None)
    if make_ref.staged():
        os.environ["DPT_REFERENCE_ROOT"] = make_ref.DST
        from oracle import ref_harness
        _dp, tu = ref_harness.load()
        if family == "spm":
            enc, _dec = tu.dp_tokenize_llama(tok)
        else:
            spec = assets.load_spec(asset_name)
            # the reference needs legacy "a b" merge strings at its hard-coded snapshot path (tokenizer_utils.py:111-119)
            spec["model"]["merges"] = [m if isinstance(m, str) else " ".join(m) for m in spec["model"]["merges"]]
            cache = tempfile.mkdtemp(prefix="dpt_ref_cache_")
            d = os.path.join(cache, *_BLOOM_SNAPSHOT)
            os.makedirs(d)
            with open(os.path.join(d, "tokenizer.json"), "w") as f:
                json.dump(spec, f)
            enc, _dec = tu.dp_tokenize_bloom(tok, cache)
        _STATE["encode"] = enc
        return
    if family == "spm":
        t2i = tok.get_vocab()
        _STATE.update(t2i=t2i, vocab=set(t2i), inv={i: t for t, i in t2i.items()})
    else:
        _STATE.update(v2i={t: k for k, t in enumerate(assets.load_spec(asset_name)["model"]["vocab"])})


def _encode_doc(doc: str):
    enc = _STATE["encode"]
    if enc is not None:  # the unmodified reference
        return len(doc.encode("utf-8")), len(enc(doc))
    dp, tok = _STATE["dp"], _STATE["tok"]
    n = 0
    if _STATE["family"] == "spm":
        inv, vocab = _STATE["inv"], _STATE["vocab"]
        toks = [inv[i] for i in tok.encode(doc)]
        words = []
        for k, t in enumerate(toks):
            if k == 0 or t.startswith("▁"):
                words.append(t)
            else:
                words[-1] += t
        for w in words:
            options, _ = dp.enumerate_shortest(w, vocab)
            n += len(dp.pick_longest_token(options))
    else:
        v2i = _STATE["v2i"]
        for piece, _span in tok._tokenizer.pre_tokenizer.pre_tokenize_str(doc):
            options, _ = dp.enumerate_shortest([c for c in piece], v2i)
            n += len(dp.pick_longest_token(options))
    return len(doc.encode("utf-8")), n


def _root():
    return os.path.dirname(_HERE)


def run(asset_name: str, docs, budget_s: float = 15.0, procs: int | None = None, family: str = "spm"):
    """Tokenize ``docs`` (list of str) for about ``budget_s`` seconds on ``procs`` processes.
    -> dict(bytes_per_s, tokens_per_s, cores, bytes, tokens, docs, seconds, kind)"""
    r = Runner(asset_name, procs, family)
    try:
        r.warm(docs)
        out = r.run(docs, budget_s)
    finally:
        r.close()
    return out


class Runner:
    """One worker pool kept across several bounded runs (bench.py --impl reference: one run per step)."""

    def __init__(self, asset_name: str, procs: int | None = None, family: str = "spm"):
        self.procs = procs or os.cpu_count() or 1
        self.kind = kind()
        self.pool = mp.get_context("spawn").Pool(self.procs, initializer=_init, initargs=(asset_name, family, _root()))

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
                    docs=done_d, seconds=dt, kind=self.kind)

    def close(self):
        self.pool.terminate()
