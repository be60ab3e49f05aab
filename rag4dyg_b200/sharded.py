"""Sharding across the GPUs of one node.

One process per GPU.  Two ways to split the path:
* QUERY sharding (`query_shard`, `jaccard_topk_query_sharded`): the pool state is replicated (C4: 2.56 GB of bitsets +
  a 28 MB postings index per GPU), every rank scores its own contiguous slice of the queries; there is no data-path
  collective at all (an optional all-gather returns the whole [Q, K] result to every rank).  The choice for Q >> N/G.
* POOL sharding (everything else in this file): for pools that do not fit one GPU.  The pool is split row-wise into
`world` contiguous shards (pool_base = first global row of the shard); queries are replicated; every rank runs the fused scorer + top-K on its shard; ONE exchange follows: an
all-gather of the per-shard [Q, K] candidate lists (NCCL over NVLink/NVSwitch, gloo in CPU tests), then a local merge
with the same exact comparator and (score desc, global index asc) tie rule, so the result does not depend on `world`.
The reference has no multi-GPU code on this path (its DDP/DataParallel paths only wrap model training).
"""
import ctypes

import torch
import torch.distributed as dist

from . import engine


def shard_bounds(n_rows, world):
    """Contiguous, balanced row ranges: shard r owns [bounds[r], bounds[r+1])."""
    return [r * n_rows // world for r in range(world + 1)]


def my_shard(n_rows, rank=None, world=None):
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    b = shard_bounds(n_rows, world)
    return b[rank], b[rank + 1]


def gather_candidates(parts, group=None):
    """all-gather a tuple of [Q, K] tensors into [world, Q, K] tensors (works for CUDA/nccl and CPU/gloo)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return tuple(p.unsqueeze(0).contiguous() for p in parts)
    world = dist.get_world_size(group)
    out = []
    for p in parts:
        p = p.contiguous()
        # concatenated-along-dim-0 layout is accepted by both NCCL and gloo; viewed as [world, Q, K]
        g = torch.empty((world * p.shape[0],) + tuple(p.shape[1:]), dtype=p.dtype, device=p.device)
        dist.all_gather_into_tensor(g, p, group=group)
        out.append(g.view((world,) + tuple(p.shape)))
    return tuple(out)


class P2PExchange:
    """Fused exchange over NVLink peer memory (second step): symmetric-memory gather buffers
    [n_planes][world][nq][k] (4-byte elements), double buffered.  Each rank's final merge kernel stores its lists into
    slot `rank` of EVERY peer's buffer (r4d_*_topk_scatter), one cross-GPU barrier follows, then every rank merges its
    own buffer locally.  Replaces the three (two) NCCL all-gathers of the plain path.

    Double buffering makes one barrier per step sufficient: a peer can only overwrite buffer b again two steps later,
    which is behind the next step's barrier, which this rank reaches only after its merge of buffer b was enqueued."""

    def __init__(self, nq, k, n_planes, group=None, device=None):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        if hasattr(symm, "enable_symm_mem_for_group"):  # needed by older torch, a deprecated no-op in newer ones
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                try:
                    symm.enable_symm_mem_for_group(self.group.group_name)
                except Exception:
                    pass
        self.bufs, self.hdls, self.ptrs = [], [], []
        for _ in range(2):
            t = symm.empty((n_planes, self.world, nq, k), dtype=torch.int32, device=device)
            h = symm.rendezvous(t, self.group.group_name)
            off = int(getattr(h, "offset", 0) or 0)
            self.bufs.append(t)
            self.hdls.append(h)
            self.ptrs.append((ctypes.c_void_p * self.world)(*[int(p) + off for p in h.buffer_ptrs]))
        self.step = 0

    def next(self):
        b = self.step & 1
        self.step += 1
        return self.bufs[b], self.hdls[b], self.ptrs[b]


def query_shard(q_ids, q_off, rank=None, world=None):
    """This rank's contiguous slice of CSR queries: (ids, offsets rebased to 0, first global query row)."""
    nq = q_off.numel() - 1
    a, b = my_shard(nq, rank, world)
    o = q_off[a:b + 1]
    return q_ids[int(o[0]):int(o[-1])].contiguous(), (o - o[0]).contiguous(), a


def jaccard_topk_query_sharded(pool, q_ids, q_off, k, zero_diag=False, gather=False, group=None):
    """Query-sharded Jaccard top-K.  pool: a JaccardPool holding the WHOLE pool on this rank; q_ids/q_off: the FULL
    query set (CSR, CUDA tensors, the same on every rank).  Every rank scores rows my_shard(nq); returns its
    (inter, union, idx) [Q/world, K] and the first global query row, or — gather=True — the whole [Q, K] result on
    every rank (one all-gather of the outputs; needs equal slices, i.e. nq divisible by world)."""
    ids, off, first = query_shard(q_ids, q_off, None if dist.is_initialized() else 0, None if dist.is_initialized() else 1)
    parts = pool.topk(ids, off, k, zero_diag=zero_diag, query_base=first)
    if not gather or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return parts, first
    world = dist.get_world_size(group)
    if (q_off.numel() - 1) % world:
        raise engine.R4DError("jaccard_topk_query_sharded(gather=True): the query count must be divisible by the world size")
    g = gather_candidates(parts, group)
    return tuple(t.reshape(-1, t.shape[-1]) for t in g), 0


def jaccard_pool_topk_sharded(pool_shard, q_ids, q_off, k, zero_diag=False, query_base=0, group=None, exchange=None):
    """Pool-sharded Jaccard top-K over a JaccardPool shard (pool_shard.pool_base = first global row of the shard),
    queries as CSR id lists replicated on every rank.  Local fused top-K (postings path when the shard has an index),
    ONE exchange of [Q, K] candidates, merge.  exchange: a P2PExchange(nq, k, 3) selects the fused NVLink path (the
    top-K kernel itself stores into every peer's gather buffer)."""
    nq = q_off.numel() - 1
    if exchange is not None and pool_shard.index is not None:
        buf, hdl, ptrs = exchange.next()
        engine.jaccard_topk_postings_scatter(q_ids, q_off, pool_shard.index, k, ptrs, exchange.world, exchange.rank,
                                             zero_diag=zero_diag, query_base=query_base, pool_base=pool_shard.pool_base,
                                             workspace=pool_shard.workspace(nq, k))
        hdl.barrier()
        return engine.jaccard_topk_merge(buf[0], buf[1], buf[2], k)
    parts = pool_shard.topk(q_ids, q_off, k, zero_diag=zero_diag, query_base=query_base)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return parts
    gi, gu, gx = gather_candidates(parts, group)
    return engine.jaccard_topk_merge(gi, gu, gx, k)


def jaccard_topk_sharded(q, p_shard, k, pool_base, zero_diag=False, query_base=0, group=None, workspace=None,
                         exchange=None):
    """q: replicated query bitsets; p_shard: this rank's pool rows.  Returns the GLOBAL (inter, union, idx) [Q, K].
    exchange: a P2PExchange(nq, k, 3) selects the fused NVLink path instead of NCCL all-gathers."""
    if exchange is not None:
        buf, hdl, ptrs = exchange.next()
        engine.jaccard_topk_scatter(q, p_shard, k, ptrs, exchange.world, exchange.rank, zero_diag=zero_diag,
                                    query_base=query_base, pool_base=pool_base, workspace=workspace)
        hdl.barrier()
        return engine.jaccard_topk_merge(buf[0], buf[1], buf[2], k)
    parts = engine.jaccard_topk(q, p_shard, k, zero_diag=zero_diag, query_base=query_base, pool_base=pool_base,
                                workspace=workspace)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return parts
    gi, gu, gx = gather_candidates(parts, group)
    return engine.jaccard_topk_merge(gi, gu, gx, k)


def dense_topk_sharded(q, p_shard, k, pool_base, mode=engine.DENSE_HALF_COS, q_time=None, p_time=None, lam=0.0,
                       group=None, workspace=None, exchange=None):
    """Dense analogue: local tcgen05 top-K on the shard, all-gather (or fused P2PExchange(nq, k, 2)), merge."""
    if exchange is not None:
        buf, hdl, ptrs = exchange.next()
        engine.dense_topk_scatter(q, p_shard, k, ptrs, exchange.world, exchange.rank, mode, q_time, p_time, lam,
                                  pool_base=pool_base, workspace=workspace)
        hdl.barrier()
        return engine.dense_topk_merge(buf[0].view(torch.float32), buf[1], k)
    parts = engine.dense_topk(q, p_shard, k, mode, q_time, p_time, lam, pool_base=pool_base, workspace=workspace)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return parts
    gs, gi = gather_candidates(parts, group)
    return engine.dense_topk_merge(gs, gi, k)
