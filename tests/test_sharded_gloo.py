"""CPU, world_size 2, gloo: the host side of the N>1 path — shard bounds, pool_base offsets, the [world, Q, K]
all-gather layout of rag4dyg_b200.sharded.gather_candidates — checked end to end with the oracle standing in for the
per-shard GPU scorer (the scorer itself has no CPU implementation; its shard invariance is a GPU test)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import random_sets, to_csr
from oracle import jaccard_oracle as jo
from rag4dyg_b200 import sharded


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 1000, 1_000_000):
        for w in (1, 2, 3, 8):
            b = sharded.shard_bounds(n, w)
            assert b[0] == 0 and b[-1] == n and len(b) == w + 1
            sizes = np.diff(b)
            assert sizes.min() >= 0 and sizes.max() - sizes.min() <= 1


def _oracle_merge(gi, gu, gx, k):
    """numpy restatement of the merge: exact rational compare via cross-multiplication, index tiebreak."""
    world, nq, kin = gx.shape
    oi, ou, ox = np.zeros((nq, k), np.int64), np.ones((nq, k), np.int64), np.full((nq, k), 0x7FFFFFFF, np.int64)
    for q in range(nq):
        cands = [(int(gi[w, q, t]), int(gu[w, q, t]), int(gx[w, q, t])) for w in range(world) for t in range(kin)
                 if gx[w, q, t] != 0x7FFFFFFF]
        import functools

        def cmp(a, b):
            l, r = a[0] * b[1], b[0] * a[1]
            if l != r:
                return -1 if l > r else 1
            return -1 if a[2] < b[2] else (1 if a[2] > b[2] else 0)
        cands.sort(key=functools.cmp_to_key(cmp))
        for t, c in enumerate(cands[:k]):
            oi[q, t], ou[q, t], ox[q, t] = c
    return oi, ou, ox


def _worker(rank, world, port, q, p, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharded.my_shard(len(p))
        li, lu, lx = jo.c_topk(*to_csr(q), *to_csr(p[lo:hi]), k, pool_base=lo)      # stand-in for the GPU scorer
        parts = (torch.from_numpy(li.astype(np.int32)), torch.from_numpy(lu.astype(np.int32)), torch.from_numpy(lx))
        gi, gu, gx = sharded.gather_candidates(parts)
        assert gx.shape == (world, len(q), k)
        assert torch.equal(gx[rank], parts[2])                                      # own slot holds own candidates
        oi, ou, ox = _oracle_merge(gi.numpy(), gu.numpy(), gx.numpy(), k)
        if rank == 0:
            ret.put((oi, ou, ox))
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_and_merge_matches_unsharded_oracle():
    rng = np.random.default_rng(42)
    q = random_sets(rng, 40, 200, mean=3, p_empty=0.05)
    p = random_sets(rng, 501, 200, mean=3, p_empty=0.05)
    k = 10
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, p, k, ret)) for r in range(2)]
    for pr in procs:
        pr.start()
    oi, ou, ox = ret.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    ri, ru, rx = jo.c_topk(*to_csr(q), *to_csr(p), k)
    assert np.array_equal(ox, rx) and np.array_equal(oi, ri) and np.array_equal(ou, ru)


class _OraclePool:
    """Stand-in for JaccardPool on CPU (the scorer has no CPU implementation): same .topk signature, oracle inside."""

    def __init__(self, p):
        self.p = to_csr(p)

    def topk(self, q_ids, q_off, k, zero_diag=False, query_base=0, out=None):
        li, lu, lx = jo.c_topk(q_ids.numpy(), q_off.numpy(), *self.p, k, zero_diag=zero_diag, query_base=query_base)
        return (torch.from_numpy(li.astype(np.int32)), torch.from_numpy(lu.astype(np.int32)), torch.from_numpy(lx))


def _qworker(rank, world, port, q, p, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        qi, qo = to_csr(q)
        qi, qo = torch.from_numpy(qi), torch.from_numpy(qo)
        ids, off, first = sharded.query_shard(qi, qo)
        a, b = sharded.my_shard(len(q))
        assert first == a and off.numel() == b - a + 1 and int(off[0]) == 0 and int(off[-1]) == ids.numel()
        # train x train with the diagonal zeroed: the slice must carry its global query row (query_base)
        (gi, gu, gx), base = sharded.jaccard_topk_query_sharded(_OraclePool(p), qi, qo, k, zero_diag=True, gather=True)
        assert base == 0 and gx.shape == (len(q), k)
        if rank == 0:
            ret.put((gi.numpy(), gu.numpy(), gx.numpy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_query_sharding_matches_unsharded_oracle():
    rng = np.random.default_rng(43)
    p = random_sets(rng, 300, 150, mean=3, p_empty=0.05)
    q = p[:64]                                   # queries = pool rows: zero_diag needs the global row of every query
    k = 10
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_qworker, args=(r, 2, port, q, p, k, ret)) for r in range(2)]
    for pr in procs:
        pr.start()
    gi, gu, gx = ret.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    ri, ru, rx = jo.c_topk(*to_csr(q), *to_csr(p), k, zero_diag=True)
    assert np.array_equal(gx, rx) and np.array_equal(gi, ri) and np.array_equal(gu, ru)
