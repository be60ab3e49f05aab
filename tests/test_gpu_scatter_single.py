"""GPU, ONE device: the fused-exchange kernels (r4d_jaccard_topk_scatter, r4d_jaccard_topk_postings_scatter,
r4d_dense_topk_scatter) exercised as a world of 2 whose two "peer" gather buffers both live on this GPU.  Rank 0 and
rank 1 run back to back on their pool shards, each storing its final lists into slot `rank` of BOTH buffers; every
buffer is then merged (r4d_*_topk_merge) and must equal the oracle / the unsharded call.  The same kernels and the same
peer-pointer arithmetic run over NVLink at N >= 2 (tools/multi_gpu_check.py, logs under profiles/)."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import random_sets, to_csr
from oracle import dense_oracle
from oracle import jaccard_oracle as jo

pytestmark = pytest.mark.gpu

from rag4dyg_b200 import engine, set_encoder  # noqa: E402
from rag4dyg_b200.jaccard_pool import JaccardPool  # noqa: E402

WORLD = 2


def _peer_bufs(n_planes, nq, k):
    bufs = [torch.full((n_planes, WORLD, nq, k), -1, dtype=torch.int32, device="cuda") for _ in range(WORLD)]
    ptrs = (ctypes.c_void_p * WORLD)(*[b.data_ptr() for b in bufs])
    return bufs, ptrs


@pytest.mark.parametrize("path", ["bitsets", "postings"])
def test_jaccard_fused_exchange_kernels_world2_on_one_gpu(path):
    rng = np.random.default_rng(77)
    n_bits, nq, npool, k = 20000, 900, 40001, 10
    q = random_sets(rng, nq, n_bits, mean=2.2, p_empty=0.02, dup=True)
    p = random_sets(rng, npool, n_bits, mean=2.2)
    bounds = [0, npool // 2, npool]
    bufs, ptrs = _peer_bufs(3, nq, k)
    qi, qo = to_csr(q)
    dqi, dqo = torch.as_tensor(qi).cuda(), torch.as_tensor(qo).cuda()
    bq = set_encoder.encode_csr(qi, qo, n_bits)
    for rank in range(WORLD):
        lo, hi = bounds[rank], bounds[rank + 1]
        shard = JaccardPool.from_csr(*to_csr(p[lo:hi]), n_bits, pool_base=lo, postings=(path == "postings"))
        if path == "postings":
            engine.jaccard_topk_postings_scatter(dqi, dqo, shard.index, k, ptrs, WORLD, rank, pool_base=lo)
        else:
            engine.jaccard_topk_scatter(bq, shard.bits, k, ptrs, WORLD, rank, pool_base=lo)
    oi, ou, ox = jo.c_topk(qi, qo, *to_csr(p), k)
    for b in bufs:                                   # every "peer" received both ranks' lists
        mi, mu, mx = engine.jaccard_topk_merge(b[0], b[1], b[2], k)
        assert np.array_equal(mx.cpu().numpy(), ox)
        assert np.array_equal(mi.cpu().numpy(), oi) and np.array_equal(mu.cpu().numpy(), ou)


@pytest.mark.parametrize("prec", [engine.PREC_BF16, engine.PREC_BF16X3])
def test_dense_fused_exchange_kernels_world2_on_one_gpu(prec):
    g = torch.Generator().manual_seed(11)
    npool, nq, d, k = 30000, 300, 256, 10
    pe, qe = torch.randn(npool, d, generator=g), torch.randn(nq, d, generator=g)
    tp, tq = (torch.rand(npool, generator=g) * 110).cuda(), (torch.rand(nq, generator=g) * 110).cuda()
    qp = engine.dense_prepare(qe.cuda(), prec)
    pa = engine.dense_prepare(pe.cuda(), prec)
    bounds = [0, npool // 2, npool]
    bufs, ptrs = _peer_bufs(2, nq, k)
    for rank in range(WORLD):
        lo, hi = bounds[rank], bounds[rank + 1]
        engine.dense_topk_scatter(qp, pa.rows(lo, hi), k, ptrs, WORLD, rank, engine.DENSE_COS_DECAY, tq,
                                  tp[lo:hi].contiguous(), 0.01, pool_base=lo)
    rs, ri = engine.dense_topk(qp, pa, k, engine.DENSE_COS_DECAY, tq, tp, 0.01)
    ref = dense_oracle.scores(qe, pe, 1, tq.cpu(), tp.cpu(), 0.01).numpy()
    tol = 1e-5 if prec == engine.PREC_BF16X3 else 3e-3
    for b in bufs:
        ms, mi = engine.dense_topk_merge(b[0].view(torch.float32), b[1], k)
        assert torch.equal(mi, ri) and torch.equal(ms, rs)          # same arithmetic per pair: sharding changes nothing
        assert not dense_oracle.topk_tolerance_ok(ref, mi.cpu().numpy(), ms.cpu().numpy(), k, tol)
