"""GPU parity: tcgen05 dense scorer vs the torch-CPU fp32 oracle (oracle/dense_oracle.py).
Floating point -> tolerances are explicit:
  BF16X3 (hi/lo split, fp32 accumulate): |score - fp32 reference| <= 5e-6, ranking tolerance-aware at the same tol
  BF16   (single pass)                 : |score - fp32 reference| <= 3e-3
  decay factor uses ex2.approx (rel. error ~1e-7), covered by the same tolerances."""
import numpy as np
import pytest
import torch

from oracle import dense_oracle as do

pytestmark = pytest.mark.gpu

from rag4dyg_b200 import engine  # noqa: E402
from rag4dyg_b200.engine import DENSE_COS_DECAY, DENSE_HALF_COS, DENSE_HALF_COS_DECAY, PREC_BF16, PREC_BF16X3  # noqa: E402

TOL = {PREC_BF16X3: 5e-6, PREC_BF16: 3e-3}


def make(nq, npool, d, seed, correlated=True):
    g = torch.Generator().manual_seed(seed)
    p = torch.randn(npool, d, generator=g)
    q = torch.randn(nq, d, generator=g)
    if correlated:  # GPT-2-like embeddings share a large mean component: cosines crowd near 1 (SURVEY hard part 5)
        mean = torch.randn(1, d, generator=g) * 2.0
        p, q = p + mean, q + mean
    tq = torch.rand(nq, generator=g) * 110.0
    tp = torch.rand(npool, generator=g) * 110.0
    return q, p, tq, tp


@pytest.mark.parametrize("prec", [PREC_BF16X3, PREC_BF16])
@pytest.mark.parametrize("nq,npool,d", [(110, 1708, 512), (1, 1, 64), (130, 300, 100), (257, 1000, 768), (64, 5000, 256)])
def test_dense_full_scores(nq, npool, d, prec):
    q, p, _, _ = make(nq, npool, d, nq + npool + d)
    ref = do.score_block(q, p).numpy()
    got = engine.dense_full(engine.dense_prepare(q.cuda(), prec), engine.dense_prepare(p.cuda(), prec)).cpu().numpy()
    err = np.abs(got - ref).max()
    assert err <= TOL[prec], f"max |score err| {err:.3g} > {TOL[prec]}"


@pytest.mark.parametrize("mode,lam", [(DENSE_COS_DECAY, 1e-4), (DENSE_COS_DECAY, 0.1), (DENSE_HALF_COS_DECAY, 0.1)])
def test_dense_decay_epilogue(mode, lam):
    q, p, tq, tp = make(200, 3000, 512, 99)
    ref = do.scores(q, p, mode, tq, tp, lam).numpy()
    got = engine.dense_full(engine.dense_prepare(q.cuda(), PREC_BF16X3), engine.dense_prepare(p.cuda(), PREC_BF16X3), mode,
                            tq.cuda(), tp.cuda(), lam).cpu().numpy()
    err = np.abs(got - ref).max()
    assert err <= TOL[PREC_BF16X3], f"max |score err| {err:.3g}"


@pytest.mark.parametrize("prec", [PREC_BF16X3, PREC_BF16])
@pytest.mark.parametrize("nq,npool,d,k,mode", [
    (110, 1708, 512, 10, DENSE_HALF_COS), (300, 20000, 768, 10, DENSE_HALF_COS), (129, 9000, 256, 7, DENSE_COS_DECAY),
    (5, 3, 64, 10, DENSE_HALF_COS), (64, 4097, 128, 32, DENSE_HALF_COS_DECAY),
])
def test_dense_topk_tolerance_aware(nq, npool, d, k, mode, prec):
    q, p, tq, tp = make(nq, npool, d, nq * 3 + npool)
    lam = 0.05
    ref = do.scores(q, p, mode, tq, tp, lam).numpy()
    ts, ti = engine.dense_topk(engine.dense_prepare(q.cuda(), prec), engine.dense_prepare(p.cuda(), prec), k, mode,
                               tq.cuda(), tp.cuda(), lam)
    ts, ti = ts.cpu().numpy(), ti.cpu().numpy()
    kk = min(k, npool)
    assert np.all(ti[:, kk:] == engine.R4D_IDX_NONE)
    bad = do.topk_tolerance_ok(ref, ti, ts, kk, TOL[prec])
    assert not bad, bad[:5]


def test_dense_topk_equals_own_full_ranking():
    """Internal consistency, exact: fused top-K == stable ranking of the kernel's own full score rows."""
    q, p, _, _ = make(200, 6000, 512, 5)
    bq, bp = engine.dense_prepare(q.cuda(), PREC_BF16X3), engine.dense_prepare(p.cuda(), PREC_BF16X3)
    full = engine.dense_full(bq, bp)
    ts, ti = engine.dense_topk(bq, bp, 10)
    order = engine.rank_rows(full)[:, :10]
    assert torch.equal(ti, order)
    assert torch.equal(ts, torch.gather(full, 1, order.long()))


def test_dense_sharded_merge_is_shard_invariant():
    q, p, _, _ = make(100, 5000, 256, 8)
    bq, bp = engine.dense_prepare(q.cuda(), PREC_BF16X3), engine.dense_prepare(p.cuda(), PREC_BF16X3)
    ref_s, ref_i = engine.dense_topk(bq, bp, 10)
    bounds = [0, 1300, 2600, 5000]
    parts = [engine.dense_topk(bq, bp.rows(a, b), 10, pool_base=a) for a, b in zip(bounds[:-1], bounds[1:])]
    s = torch.stack([x[0] for x in parts]).contiguous()
    i = torch.stack([x[1] for x in parts]).contiguous()
    ms, mi = engine.dense_topk_merge(s, i, 10)
    assert torch.equal(mi, ref_i) and torch.equal(ms, ref_s)
