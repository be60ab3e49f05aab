"""Pool sharding across the GPUs of one node.

One process per GPU.  The pool is split row-wise into `world` contiguous shards (pool_base = first global row of the
shard); queries are replicated; every rank runs the fused scorer + top-K on its shard; ONE exchange follows: an
all-gather of the per-shard [Q, K] candidate lists (NCCL over NVLink/NVSwitch, gloo in CPU tests), then a local merge
with the same exact comparator and (score desc, global index asc) tie rule, so the result does not depend on `world`.
The reference has no multi-GPU code on this path (its DDP/DataParallel paths only wrap model training).
"""
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


def jaccard_topk_sharded(q, p_shard, k, pool_base, zero_diag=False, query_base=0, group=None, workspace=None):
    """q: replicated query bitsets; p_shard: this rank's pool rows.  Returns the GLOBAL (inter, union, idx) [Q, K]."""
    parts = engine.jaccard_topk(q, p_shard, k, zero_diag=zero_diag, query_base=query_base, pool_base=pool_base,
                                workspace=workspace)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return parts
    gi, gu, gx = gather_candidates(parts, group)
    return engine.jaccard_topk_merge(gi, gu, gx, k)


def dense_topk_sharded(q, p_shard, k, pool_base, mode=engine.DENSE_HALF_COS, q_time=None, p_time=None, lam=0.0,
                       group=None, workspace=None):
    """Dense analogue: local tcgen05 top-K on the shard, all-gather, merge.  Returns (score, idx) [Q, K]."""
    parts = engine.dense_topk(q, p_shard, k, mode, q_time, p_time, lam, pool_base=pool_base, workspace=workspace)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return parts
    gs, gi = gather_candidates(parts, group)
    return engine.dense_topk_merge(gs, gi, k)
