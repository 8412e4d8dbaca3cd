"""Multi-GPU plumbing on CPU: byte-balanced document sharding and the counter reduction over a real
world_size-2 process group (gloo).  The data path has no collective; the only exchange is the int64[4] sum."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def test_shard_bounds_balanced_and_contiguous():
    from dptok.sharded import shard_bounds, take_shard
    rng = np.random.default_rng(0)
    lens = rng.integers(1, 5000, 10_000)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    text = np.zeros(offs[-1], dtype=np.uint8)
    for world in (1, 2, 3, 4, 8):
        b = shard_bounds(offs, world)
        assert b[0] == 0 and b[-1] == len(lens) and (np.diff(b) >= 0).all()
        sizes = [int(offs[b[r + 1]] - offs[b[r]]) for r in range(world)]
        assert sum(sizes) == offs[-1]
        assert max(sizes) - min(sizes) <= 2 * 5000
        total = 0
        for r in range(world):
            t, o = take_shard(text, offs, world, r)
            assert o[0] == 0 and o[-1] == len(t)
            total += len(t)
        assert total == len(text)
    # degenerate: fewer documents than ranks
    b = shard_bounds(np.array([0, 10, 20], dtype=np.int64), 8)
    assert b[0] == 0 and b[-1] == 2 and (np.diff(b) >= 0).all()


def test_shard_bounds_properties():
    """Any document-length profile (one giant document, more ranks than documents, equal documents) and any world size:
    the bounds are monotone, cover every document exactly once, and no rank gets more than its byte share plus one document
    (cut at the nearer boundary)."""
    from hypothesis import given, settings, strategies as st
    from dptok.sharded import shard_bounds, take_shard

    @settings(max_examples=300, deadline=None)
    @given(st.lists(st.one_of(st.integers(1, 50), st.integers(1, 100_000)), min_size=1, max_size=200), st.integers(1, 17),
           st.integers(0, 1000))
    def check(lens, world, base):
        offs = base + np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        b = shard_bounds(offs, world)
        assert len(b) == world + 1 and b[0] == 0 and b[-1] == len(lens) and (np.diff(b) >= 0).all()
        share = (offs[-1] - offs[0]) / world
        sizes = [int(offs[b[r + 1]] - offs[b[r]]) for r in range(world)]
        assert sum(sizes) == offs[-1] - offs[0]
        assert max(sizes) <= share + max(lens) + 1
        if base == 0:
            text = np.zeros(int(offs[-1]), dtype=np.uint8)
            got = 0
            for r in range(world):
                t, o = take_shard(text, offs, world, r)
                assert o[0] == 0 and o[-1] == len(t) and (np.diff(o) > 0).all()
                got += len(t)
            assert got == len(text)

    check()


def _worker(rank, world, port, out):
    for p in (ROOT, PKG):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dptok.sharded import ShardedTokenizer

    class Fake:
        def __init__(self, counters):
            self.counters = counters

    def encode_fn(text, doc_offs):
        # stand-in for Engine.encode_corpus: what matters here is the sharding + reduction plumbing
        words = int((text == 32).sum()) + (len(doc_offs) - 1)
        return Fake(torch.tensor([len(text), words, words * 2, rank], dtype=torch.int64))

    rng = np.random.default_rng(7)
    lens = rng.integers(50, 400, 1000)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    text = rng.choice(np.array([32, 97, 98], dtype=np.uint8), offs[-1])
    res, stats = ShardedTokenizer(encode_fn).run_global(text, offs)
    out[rank] = (stats.bytes, stats.words, stats.tokens, stats.untokenizable, int(res.counters[0]))
    dist.destroy_process_group()


def test_counter_reduction_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    a, b = out[0], out[1]
    assert a[:4] == b[:4], "every rank must derive the same global statistics"
    rng = np.random.default_rng(7)
    lens = rng.integers(50, 400, 1000)
    total = int(lens.sum())
    assert a[0] == total and a[4] + b[4] == total and a[3] == 1
    assert abs(a[4] - b[4]) <= 800


def _worker_empty(rank, world, port, out):
    for p in (ROOT, PKG):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dptok.sharded import ShardedTokenizer

    class Fake:
        def __init__(self, counters):
            self.counters = counters

    def encode_fn(text, doc_offs):
        assert len(text) > 0 and len(doc_offs) > 1, "an empty shard must not reach the encoder"
        return Fake(torch.tensor([len(text), len(doc_offs) - 1, 7, 0], dtype=torch.int64))

    offs = np.array([0, 1000], dtype=np.int64)       # ONE document, two ranks: one rank has nothing to do
    text = np.full(1000, 97, dtype=np.uint8)
    res, stats = ShardedTokenizer(encode_fn).run_global(text, offs)
    out[rank] = (stats.bytes, stats.words, stats.tokens, res is None)
    dist.destroy_process_group()


def test_empty_shard_still_enters_the_reduction_world2_gloo():
    """More ranks than documents: the rank without documents contributes zeros instead of raising (the others would wait
    in all_reduce forever)."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_empty, args=(2, port, out), nprocs=2, join=True)
    a, b = out[0], out[1]
    assert a[:3] == b[:3] == (1000, 1, 7)
    assert sorted([a[3], b[3]]) == [False, True]
