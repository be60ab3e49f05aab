"""GPU parity on stress shapes: very wide vocabularies (hundreds of TMA chunks), saturated bitsets, odd word counts,
embedding widths from 64 to 1024 (resident-query and streaming variants of the CTA-pair kernel, general kernel)."""
import numpy as np
import pytest
import torch

from conftest import random_sets, to_csr
from oracle import dense_oracle as do
from oracle import jaccard_oracle as jo

pytestmark = pytest.mark.gpu

from rag4dyg_b200 import engine, set_encoder  # noqa: E402


@pytest.mark.parametrize("n_bits,mean,max_len", [(200_000, 5, 64), (65_537, 30, 200), (1025, 300, 1025), (4097, 2, 8)])
def test_jaccard_wide_and_saturated(n_bits, mean, max_len):
    rng = np.random.default_rng(n_bits)
    q = random_sets(rng, 140, n_bits, mean=mean, max_len=max_len, p_empty=0.05, dup=True)
    p = random_sets(rng, 700, n_bits, mean=mean, max_len=max_len, p_empty=0.05, dup=True)
    if n_bits <= 4097:
        p[3] = list(range(n_bits))            # a saturated row: inter == |q| for every query
        q[5] = list(range(n_bits))
    bq = set_encoder.encode_csr(*to_csr(q), n_bits)
    bp = set_encoder.encode_csr(*to_csr(p), n_bits)
    inter, score = engine.jaccard_full(bq, bp)
    ci, cu = jo.c_counts(*to_csr(q), *to_csr(p))
    assert np.array_equal(inter.cpu().numpy(), ci)
    assert np.array_equal(score.cpu().numpy(), jo.scores_from_counts(ci.astype(np.int64), cu.astype(np.int64)))
    for k in (1, 10, 32):
        ti, tu, tx = engine.jaccard_topk(bq, bp, k)
        oi, ou, ox = jo.c_topk(*to_csr(q), *to_csr(p), k)
        assert np.array_equal(tx.cpu().numpy(), ox) and np.array_equal(ti.cpu().numpy(), oi)
        assert np.array_equal(tu.cpu().numpy(), ou)


@pytest.mark.parametrize("d", [64, 200, 512, 768, 1024])
@pytest.mark.parametrize("prec,k", [(engine.PREC_BF16, 10), (engine.PREC_BF16, 32), (engine.PREC_BF16X3, 10)])
def test_dense_widths(d, prec, k):
    g = torch.Generator().manual_seed(d + k)
    q, p = torch.randn(300, d, generator=g) + 0.2, torch.randn(20_000, d, generator=g) + 0.2
    tq, tp = torch.rand(300, generator=g) * 50, torch.rand(20_000, generator=g) * 50
    tol = 3e-3 if prec == engine.PREC_BF16 else 1e-5
    ref = do.scores(q, p, 2, tq, tp, 0.02).numpy()
    qp, pp = engine.dense_prepare(q.cuda(), prec), engine.dense_prepare(p.cuda(), prec)
    full = engine.dense_full(qp, pp, engine.DENSE_HALF_COS_DECAY, tq.cuda(), tp.cuda(), 0.02)
    assert np.abs(full.cpu().numpy() - ref).max() <= tol
    ts, ti = engine.dense_topk(qp, pp, k, engine.DENSE_HALF_COS_DECAY, tq.cuda(), tp.cuda(), 0.02)
    assert not do.topk_tolerance_ok(ref, ti.cpu().numpy(), ts.cpu().numpy(), k, tol)
    # exact internal consistency: top-K == stable ranking of the kernel's own full rows (same arithmetic per pair)
    if prec == engine.PREC_BF16X3 or k > 16:
        order = engine.rank_rows(full)[:, :k]
        assert torch.equal(ti, order)


def test_negative_lambda_disables_bound_skip():
    """lambda < 0 (the reference's argparse default is -1) makes the 'decay' factor > 1: the raw-cosine upper bound is
    invalid there and the kernel must fall back to evaluating every score."""
    g = torch.Generator().manual_seed(9)
    q, p = torch.randn(200, 256, generator=g), torch.randn(30_000, 256, generator=g)
    tq, tp = torch.rand(200, generator=g), torch.rand(30_000, generator=g)
    ref = do.scores(q, p, 1, tq, tp, -1.0).numpy()
    qp, pp = engine.dense_prepare(q.cuda(), engine.PREC_BF16), engine.dense_prepare(p.cuda(), engine.PREC_BF16)
    ts, ti = engine.dense_topk(qp, pp, 10, engine.DENSE_COS_DECAY, tq.cuda(), tp.cuda(), -1.0)
    assert not do.topk_tolerance_ok(ref, ti.cpu().numpy(), ts.cpu().numpy(), 10, 8e-3)
