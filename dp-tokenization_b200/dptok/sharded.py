"""Document-sharded multi-GPU driver: one process per GPU, no data-path collective.

Replaces the per-document loops that sum lengths in main_analyze_s2orc.py:269-298 and
main_biomed_translation.py:71-82.  Documents (and words inside them) are independent, so each rank
tokenizes a contiguous, byte-balanced range of documents with its own replica of the compiled vocabulary;
the only exchange is ONE all_reduce(SUM) of the int64[4] counter vector {bytes, words, tokens,
untokenizable} (32 bytes over NCCL/NVLink; gloo in the CPU tests), from which every rank derives the same
compression ratio.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(doc_offs: np.ndarray, world: int) -> np.ndarray:
    """Document index boundaries int64[world+1]: contiguous ranges with (nearly) equal byte totals."""
    doc_offs = np.asarray(doc_offs, dtype=np.int64)
    n_docs = len(doc_offs) - 1
    total = int(doc_offs[-1] - doc_offs[0])
    targets = doc_offs[0] + (np.arange(1, world, dtype=np.int64) * total) // world
    cuts = np.searchsorted(doc_offs, targets, side="left")
    # choose the nearer document boundary
    for k, (c, t) in enumerate(zip(cuts, targets)):
        if 0 < c <= n_docs and abs(int(doc_offs[c - 1]) - int(t)) < abs(int(doc_offs[min(c, n_docs)]) - int(t)):
            cuts[k] = c - 1
    bounds = np.concatenate([[0], np.clip(cuts, 0, n_docs), [n_docs]]).astype(np.int64)
    return np.maximum.accumulate(bounds)


def take_shard(text: np.ndarray, doc_offs: np.ndarray, world: int, rank: int):
    """(text slice, rebased doc offsets) of this rank's documents."""
    b = shard_bounds(doc_offs, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    offs = np.asarray(doc_offs[lo:hi + 1], dtype=np.int64)
    return text[offs[0]:offs[-1]], offs - offs[0]


@dataclass
class CorpusStats:
    bytes: int
    words: int
    tokens: int
    untokenizable: int

    @property
    def bytes_per_token(self) -> float:
        return self.bytes / max(self.tokens, 1)


def reduce_counters(counters: torch.Tensor, group=None) -> CorpusStats:
    """Sum the int64[4] counter vector over all ranks (in place) and return the global statistics."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
    c = counters.cpu().tolist()
    return CorpusStats(*[int(x) for x in c])


class ShardedTokenizer:
    """Runs ``encode_fn(text, doc_offs) -> result with .counters`` on this rank's shard and reduces."""

    def __init__(self, encode_fn: Callable, world: Optional[int] = None, rank: Optional[int] = None, group=None):
        self.encode_fn = encode_fn
        self.group = group
        init = dist.is_available() and dist.is_initialized()
        self.world = world if world is not None else (dist.get_world_size(group) if init else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if init else 0)

    def run_global(self, text: np.ndarray, doc_offs: np.ndarray):
        """Every rank holds the same (text, doc_offs); each encodes its own shard.  -> (local result, stats)."""
        t, o = take_shard(text, doc_offs, self.world, self.rank)
        if len(o) <= 1 or len(t) == 0:
            # more ranks than documents, or one document that outweighs a whole share: this rank has nothing to encode.
            # It must still enter the reduction (the other ranks wait in it) - with a zero counter vector.
            dev = torch.device("cuda", torch.cuda.current_device()) if (torch.cuda.is_available() and dist.is_available() and
                                                                       dist.is_initialized() and
                                                                       dist.get_backend(self.group) == "nccl") else "cpu"
            return None, reduce_counters(torch.zeros(4, dtype=torch.int64, device=dev), self.group)
        res = self.encode_fn(t, o)
        return res, reduce_counters(res.counters.clone(), self.group)
