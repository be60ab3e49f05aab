"""GPU parity at BASELINE.json's full pool sizes (synthetic configs 4 and 5): size-independent properties + an
oracle-checked sub-block, since the CPU oracle cannot score 1e11 pairs."""
import math

import numpy as np
import pytest
import torch

from oracle import dense_oracle as do
from oracle import jaccard_oracle as jo

pytestmark = pytest.mark.gpu

from rag4dyg_b200 import engine, set_encoder  # noqa: E402


def synth_sets(n, seed, mean=1 / 0.45, vocab=20000, max_len=64):
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(n, generator=g, dtype=torch.float64).clamp_(min=1e-300)
    lens = (torch.floor(torch.log(u) / math.log(1 - 1 / mean)).to(torch.int64) + 1).clamp_(1, max_len)
    off = torch.zeros(n + 1, dtype=torch.int64)
    torch.cumsum(lens, 0, out=off[1:])
    return torch.randint(0, vocab, (int(off[-1]),), generator=g, dtype=torch.int32), off


def rows(ids, off, a, b):
    o = off[a:b + 1]
    return ids[int(o[0]):int(o[-1])].contiguous(), (o - o[0]).contiguous()


def test_jaccard_1m_pool_properties_and_oracle_block():
    """1,000,000-set pool, V = 20,000 (W = 625), K = 10; 2,048 queries = pool rows + 2,048 fresh queries."""
    n_pool, k = 1_000_000, 10
    p_ids, p_off = synth_sets(n_pool, 1234)
    q_ids, q_off = synth_sets(2048, 5678)
    bp = set_encoder.encode_csr(p_ids, p_off, 20000)
    card = bp.card.cpu().numpy()
    # (1) self retrieval: a pool row's best match has inter == union == |set| and index <= its own
    ti, tu, tx = [t.cpu().numpy() for t in engine.jaccard_topk(bp.rows(0, 2048), bp, k)]
    assert np.array_equal(ti[:, 0], card[:2048]) and np.array_equal(tu[:, 0], card[:2048])
    assert np.all(tx[:, 0] <= np.arange(2048))
    s = ti / tu
    assert np.all(np.diff(s, axis=1) <= 0) and np.all(np.diff(tx, axis=1)[np.diff(s, axis=1) == 0] > 0)
    # (2) zero_diag removes exactly the self match
    zi, zu, zx = [t.cpu().numpy() for t in engine.jaccard_topk(bp.rows(0, 2048), bp, k, zero_diag=True)]
    assert not np.any(zx == np.arange(2048)[:, None]) or np.all((zi / zu)[zx == np.arange(2048)[:, None]] == 0)
    # (3) shard invariance at full size: 8 pool shards merged == unsharded
    bq = set_encoder.encode_csr(q_ids, q_off, 20000)
    ref = engine.jaccard_topk(bq, bp, k)
    b = [i * n_pool // 8 for i in range(9)]
    parts = [engine.jaccard_topk(bq, bp.rows(b[i], b[i + 1]), k, pool_base=b[i]) for i in range(8)]
    merged = engine.jaccard_topk_merge(*[torch.stack([pp[j] for pp in parts]).contiguous() for j in range(3)], k)
    assert all(torch.equal(x, y) for x, y in zip(ref, merged))
    # (4) oracle on a 48-query sub-block against the WHOLE pool (48e6 exact pair merges on the CPU)
    qi, qo = rows(q_ids, q_off, 0, 48)
    oi, ou, ox = jo.c_topk(qi.numpy(), qo.numpy(), p_ids.numpy(), p_off.numpy(), k)
    assert np.array_equal(ref[2][:48].cpu().numpy(), ox)
    assert np.array_equal(ref[0][:48].cpu().numpy(), oi) and np.array_equal(ref[1][:48].cpu().numpy(), ou)


def test_dense_2m_pool_properties_and_oracle_block():
    """2,000,000 x 768 pool (the 10M config's shape per GPU at 4-8 GPUs), bf16 pair kernel, decay epilogue."""
    n_pool, d, k, lam = 2_000_000, 768, 10, 1e-4
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(4321)
    hi = torch.empty((n_pool, d), dtype=torch.bfloat16, device=dev)
    keep = {}
    for a in range(0, n_pool, 500_000):
        x = torch.randn((500_000, d), generator=g, device=dev)
        hi[a:a + 500_000] = engine.dense_prepare(x, engine.PREC_BF16).hi
        if a == 0:
            keep["head"] = x[:4096].clone()          # fp32 rows for the oracle / self-retrieval queries
    pool = engine.DensePlanes(hi, None, d, d, engine.PREC_BF16)
    p_time = torch.rand(n_pool, generator=g, device=dev) * 110.0
    # (1) self retrieval with zero time distance: query = pool row i with its own time -> index i on top, score ~ 1
    q = engine.dense_prepare(keep["head"][:1024], engine.PREC_BF16)
    ts, ti = engine.dense_topk(q, pool, k, engine.DENSE_COS_DECAY, p_time[:1024].contiguous(), p_time, lam)
    assert torch.equal(ti[:, 0].cpu(), torch.arange(1024, dtype=torch.int32))
    assert float((ts[:, 0] - 1).abs().max()) < 3e-3
    assert bool((ts[:, 1:] <= ts[:, :-1]).all())
    # (2) shard invariance: 4 shards merged == unsharded, bit for bit
    b = [i * n_pool // 4 for i in range(5)]
    parts = [engine.dense_topk(q, pool.rows(b[i], b[i + 1]), k, engine.DENSE_COS_DECAY, p_time[:1024].contiguous(),
                               p_time[b[i]:b[i + 1]].contiguous(), lam, pool_base=b[i]) for i in range(4)]
    ms, mi = engine.dense_topk_merge(torch.stack([x[0] for x in parts]).contiguous(),
                                     torch.stack([x[1] for x in parts]).contiguous(), k)
    assert torch.equal(mi, ti) and torch.equal(ms, ts)
    # (3) oracle on a sub-block: 16 fresh queries x the first 200,000 pool rows in fp32 on the CPU (tolerance 3e-3)
    gq = torch.Generator().manual_seed(8765)
    qf = torch.randn(16, d, generator=gq)
    tq = torch.rand(16, generator=gq) * 110.0
    sub = 200_000
    # the pool rows as the kernel sees them (bf16-rounded, normalised): exact reference for the stored operand
    pf = hi[:sub].float().cpu()
    ref = do.scores(qf, pf, 1, tq, p_time[:sub].cpu(), lam).numpy()
    ts2, ti2 = engine.dense_topk(engine.dense_prepare(qf.cuda(), engine.PREC_BF16), pool.rows(0, sub), k,
                                 engine.DENSE_COS_DECAY, tq.cuda(), p_time[:sub].contiguous(), lam)
    bad = do.topk_tolerance_ok(ref, ti2.cpu().numpy(), ts2.cpu().numpy(), k, 3e-3)
    assert not bad, bad[:3]


@pytest.mark.parametrize("nq,npool", [(0, 10), (10, 0), (0, 0)])
def test_empty_inputs(nq, npool):
    def sets(n):
        return [[1, 2, 3]] * n
    from conftest import to_csr
    bq = set_encoder.encode_csr(*to_csr(sets(nq)), 100)
    bp = set_encoder.encode_csr(*to_csr(sets(npool)), 100)
    inter, score = engine.jaccard_full(bq, bp)
    assert inter.shape == (nq, npool) and score.shape == (nq, npool)
    ti, tu, tx = engine.jaccard_topk(bq, bp, 5)
    assert tx.shape == (nq, 5)
    if nq:
        assert bool((tx == engine.R4D_IDX_NONE).all())
    qe = engine.dense_prepare(torch.randn(max(nq, 0), 64).cuda(), engine.PREC_BF16)
    pe = engine.dense_prepare(torch.randn(max(npool, 0), 64).cuda(), engine.PREC_BF16)
    assert engine.dense_full(qe, pe).shape == (nq, npool)
    ds, di = engine.dense_topk(qe, pe, 5)
    assert di.shape == (nq, 5)
    if nq:
        assert bool((di == engine.R4D_IDX_NONE).all())
    assert engine.rank_rows(torch.zeros((nq, npool), dtype=torch.float64).cuda()).shape == (nq, npool)
