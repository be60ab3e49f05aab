"""GPU parity: set encoder, Jaccard full / fused top-K / merge vs the CPU oracle.  Bit-exact (integer work)."""
import numpy as np
import pytest
import torch

from conftest import random_sets, to_csr
from oracle import jaccard_oracle as jo

pytestmark = pytest.mark.gpu

from rag4dyg_b200 import engine, set_encoder  # noqa: E402


def encode(sets, n_bits):
    bp, ro = to_csr(sets)
    return set_encoder.encode_csr(bp, ro, n_bits)


def bits_to_sets(bm):
    bits = bm.bits.cpu().numpy().view(np.uint32)
    out = []
    for r in range(bits.shape[0]):
        ids = []
        for w in np.nonzero(bits[r])[0]:
            v = int(bits[r, w])
            ids.extend(int(w) * 32 + b for b in range(32) if (v >> b) & 1)
        out.append(ids)
    return out


@pytest.mark.parametrize("n_bits", [1, 31, 32, 33, 1794, 4749, 20000])
def test_encoder_matches_python_sets(n_bits):
    rng = np.random.default_rng(n_bits)
    sets = random_sets(rng, 300, n_bits, mean=4, max_len=min(64, n_bits), p_empty=0.1, dup=True)
    bm = encode(sets, n_bits)
    assert bm.words == (n_bits + 31) // 32 and bm.pitch_words % 32 == 0
    got = bits_to_sets(bm)
    assert [sorted(set(s)) for s in sets] == got
    assert bm.card.cpu().tolist() == [len(set(s)) for s in sets]
    # padding words are zero
    assert int(bm.bits[:, bm.words:].abs().sum()) == 0


def test_encoder_full_vocab_row_and_empty_input():
    n_bits = 777
    bm = encode([list(range(n_bits)), [], [5, 5, 5]], n_bits)
    assert bm.card.cpu().tolist() == [n_bits, 0, 1]
    empty = encode([], 100)
    assert empty.n_rows == 0


def oracle_counts(q, p):
    ci, cu = jo.c_counts(*to_csr(q), *to_csr(p))
    return ci.astype(np.int64), cu.astype(np.int64)


@pytest.mark.parametrize("nq,npool,n_bits,mean", [
    (1, 1, 40, 3), (5, 7, 40, 3), (128, 128, 1000, 2.2), (129, 257, 1794, 8), (300, 1000, 4749, 2.2),
    (200, 700, 20000, 2.2), (150, 500, 20000, 20), (64, 300, 33, 10),
])
def test_full_matrix_bit_exact(nq, npool, n_bits, mean):
    rng = np.random.default_rng(nq * 1000 + npool)
    q = random_sets(rng, nq, n_bits, mean=mean, max_len=min(64, n_bits), p_empty=0.05, dup=True)
    p = random_sets(rng, npool, n_bits, mean=mean, max_len=min(64, n_bits), p_empty=0.05, dup=True)
    inter, score = engine.jaccard_full(encode(q, n_bits), encode(p, n_bits))
    ci, cu = oracle_counts(q, p)
    assert np.array_equal(inter.cpu().numpy().astype(np.int64), ci)
    ref = jo.scores_from_counts(ci, cu)
    assert np.array_equal(score.cpu().numpy(), ref), "float64 scores must be bit-identical to int/int division"


def test_full_matrix_matches_python_set_loop_and_zero_diag():
    rng = np.random.default_rng(5)
    p = random_sets(rng, 90, 60, mean=4, p_empty=0.1, dup=True)
    ps = [list(map(str, s)) for s in p]
    ref = jo.occurrence_matrix(ps, ps)          # reference algorithm verbatim (Python sets)
    np.fill_diagonal(ref, 0)
    bm = encode(p, 60)
    _, score = engine.jaccard_full(bm, bm, zero_diag=True)
    assert np.array_equal(score.cpu().numpy(), ref)


@pytest.mark.parametrize("nq,npool,n_bits,mean,k,zero_diag", [
    (3, 5, 40, 3, 10, False),          # pool shorter than k -> padded with R4D_IDX_NONE
    (130, 1000, 1000, 2.2, 10, False),
    (257, 3000, 20000, 2.2, 10, False),
    (100, 2000, 20000, 20, 7, False),
    (500, 500, 300, 3, 10, True),      # train x train with the diagonal zeroed
    (64, 5000, 4749, 2.2, 32, False),
    (40, 700, 64, 12, 1, False),
])
def test_fused_topk_bit_exact(nq, npool, n_bits, mean, k, zero_diag):
    rng = np.random.default_rng(nq + npool + k)
    p = random_sets(rng, npool, n_bits, mean=mean, max_len=min(64, n_bits), p_empty=0.05, dup=True)
    q = p[:nq] if zero_diag else random_sets(rng, nq, n_bits, mean=mean, max_len=min(64, n_bits), p_empty=0.05)
    ti, tu, tx = engine.jaccard_topk(encode(q, n_bits), encode(p, n_bits), k, zero_diag=zero_diag)
    oi, ou, ox = jo.c_topk(*to_csr(q), *to_csr(p), k, zero_diag=zero_diag)
    assert np.array_equal(tx.cpu().numpy(), ox)
    assert np.array_equal(ti.cpu().numpy().astype(np.int64), oi)
    assert np.array_equal(tu.cpu().numpy().astype(np.int64), ou)


def test_topk_equals_full_matrix_stable_ranking():
    rng = np.random.default_rng(11)
    n_bits = 500
    q = random_sets(rng, 200, n_bits, mean=3)
    p = random_sets(rng, 4000, n_bits, mean=3)
    bq, bp = encode(q, n_bits), encode(p, n_bits)
    _, score = engine.jaccard_full(bq, bp)
    order, vals = jo.topk_stable(score.cpu().numpy(), 10)
    ti, tu, tx = engine.jaccard_topk(bq, bp, 10)
    assert np.array_equal(tx.cpu().numpy(), order)
    assert np.array_equal(ti.cpu().numpy() / tu.cpu().numpy(), vals)


@pytest.mark.parametrize("n_shards", [2, 3, 8])
def test_sharded_pool_merge_is_shard_invariant(n_shards):
    rng = np.random.default_rng(n_shards)
    n_bits, npool, k = 2000, 3001, 10
    q = random_sets(rng, 150, n_bits, mean=2.2)
    p = random_sets(rng, npool, n_bits, mean=2.2)
    bq, bp = encode(q, n_bits), encode(p, n_bits)
    ref = engine.jaccard_topk(bq, bp, k)
    bounds = np.linspace(0, npool, n_shards + 1).astype(int)
    parts = [engine.jaccard_topk(bq, bp.rows(int(a), int(b)), k, pool_base=int(a)) for a, b in zip(bounds[:-1], bounds[1:])]
    stacked = [torch.stack([pp[i] for pp in parts]).contiguous() for i in range(3)]
    merged = engine.jaccard_topk_merge(*stacked, k)
    for a, b in zip(ref, merged):
        assert torch.equal(a, b)


def test_synthetic_full_width_properties():
    """BASELINE config 4 shape at reduced row counts (V=20000, W=625): self-retrieval and symmetry properties that
    need no oracle, then an oracle check on a sub-block."""
    rng = np.random.default_rng(1234)
    n_bits = 20000
    p = random_sets(rng, 20000, n_bits, mean=2.2, max_len=64)
    bp = encode(p, n_bits)
    bq = bp.rows(0, 2000)
    ti, tu, tx = engine.jaccard_topk(bq, bp, 10)
    ti, tu, tx = ti.cpu().numpy(), tu.cpu().numpy(), tx.cpu().numpy()
    card = bp.card.cpu().numpy()
    # a set's best match is a set identical to it (score 1); the first such index is <= its own index
    assert np.array_equal(ti[:, 0], card[:2000]) and np.array_equal(tu[:, 0], card[:2000])
    assert np.all(tx[:, 0] <= np.arange(2000))
    # scores are non-increasing, indices strictly increasing inside equal-score runs
    s = ti / tu
    assert np.all(np.diff(s, axis=1) <= 0)
    same = np.diff(s, axis=1) == 0
    assert np.all(np.diff(tx, axis=1)[same] > 0)
    oi, ou, ox = jo.c_topk(*to_csr(p[:300]), *to_csr(p), 10)
    assert np.array_equal(tx[:300], ox) and np.array_equal(ti[:300], oi) and np.array_equal(tu[:300], ou)


@pytest.mark.parametrize("warps", [8, 16])
def test_zero_span_skipping_and_warp_variants_are_exact(warps):
    """Every kernel variant (8/16 consumer warps, zero-span skipping on/off) returns identical bits."""
    from rag4dyg_b200 import _lib
    rng = np.random.default_rng(77)
    n_bits = 20000
    q = random_sets(rng, 300, n_bits, mean=2.2) + random_sets(rng, 60, n_bits, mean=40, max_len=64)
    p = random_sets(rng, 3000, n_bits, mean=2.2) + random_sets(rng, 500, n_bits, mean=40, max_len=64)
    bq, bp = encode(q, n_bits), encode(p, n_bits)
    prev_w = _lib.set_option("jaccard_warps", warps)
    try:
        outs = []
        for skip, sparse in ((1, 1), (1, 0), (0, 0)):      # sparse-query path, dense kernel with skipping, plain dense
            _lib.set_option("jaccard_skip_zero", skip)
            _lib.set_option("jaccard_sparse_q", sparse)
            top = engine.jaccard_topk(bq, bp, 10)
            inter, score = engine.jaccard_full(bq, bp)
            outs.append([t.clone() for t in top] + [inter, score])
        for other in outs[1:]:
            for a, b in zip(outs[0], other):
                assert torch.equal(a, b)
        oi, ou, ox = jo.c_topk(*to_csr(q), *to_csr(p), 10)
        assert np.array_equal(outs[0][2].cpu().numpy(), ox) and np.array_equal(outs[0][0].cpu().numpy(), oi)
    finally:
        _lib.set_option("jaccard_skip_zero", 1)
        _lib.set_option("jaccard_sparse_q", 1)
        _lib.set_option("jaccard_warps", prev_w)


def _index_case(case):
    import zlib
    rng = np.random.default_rng(zlib.crc32(case.encode()) % 1000)
    if case == "hot_small_universe":       # ~18 % of all pairs intersect (~2 700 per query): the queries' 1 024-slot arrays
        n_bits, k, zd = 50, 10, False      # overflow and the (stripe, query) lists behind them overflow k all the time
        q = random_sets(rng, 300, n_bits, mean=3, max_len=8, p_empty=0.05)
        p = random_sets(rng, 15000, n_bits, mean=3, max_len=8, p_empty=0.05)
    elif case == "hot_no_overflow":        # same data density, but every query's candidates fit its own array
        n_bits, k, zd = 50, 10, False
        q = random_sets(rng, 300, n_bits, mean=3, max_len=8, p_empty=0.05)
        p = random_sets(rng, 4000, n_bits, mean=3, max_len=8, p_empty=0.05)
    elif case == "duplicates":             # pool drawn from 40 distinct sets: long runs of equal scores, ties by index
        n_bits, k, zd = 2000, 10, False
        base = random_sets(rng, 40, n_bits, mean=3, max_len=8)
        p = [base[int(i)] for i in rng.integers(0, 40, 6000)]
        q = [base[int(i)] for i in rng.integers(0, 40, 200)] + random_sets(rng, 100, n_bits, mean=3)
    elif case == "multiword":              # several intersecting words per pair: each pair must be emitted once
        n_bits, k, zd = 400, 10, False
        q = random_sets(rng, 260, n_bits, mean=6, max_len=7)
        p = random_sets(rng, 3000, n_bits, mean=6, max_len=12)
    elif case == "batches":                # more than one 8 192-row query batch
        n_bits, k, zd = 5000, 10, False
        q = random_sets(rng, 8192 + 300, n_bits, mean=2.2, max_len=16, p_empty=0.02)
        p = random_sets(rng, 2000, n_bits, mean=2.2, max_len=16)
    elif case == "mixed_dense_tiles":      # tiles above the per-tile entry limit go to the dense kernel
        n_bits, k, zd = 20000, 10, False
        q = (random_sets(rng, 128, n_bits, mean=40, max_len=64) + random_sets(rng, 200, n_bits, mean=2.2) +
             random_sets(rng, 100, n_bits, mean=40, max_len=64))
        p = random_sets(rng, 3000, n_bits, mean=2.2) + random_sets(rng, 300, n_bits, mean=40, max_len=64)
    elif case == "tiny_pool":
        n_bits, k, zd = 300, 10, False
        q = random_sets(rng, 140, n_bits, mean=3, p_empty=0.2)
        p = random_sets(rng, 3, n_bits, mean=3)
    elif case == "zero_diag_self":         # train x train: the diagonal is a forced zero, hence only ever a filler
        n_bits, k, zd = 700, 10, True
        p = random_sets(rng, 1500, n_bits, mean=2.2, p_empty=0.05)
        q = p[:700]
    elif case == "wide_vocab_shared_buckets":   # 50 000 bits: index buckets shared by 4 bit ids, one 6 KB row per ring slot
        n_bits, k, zd = 50_000, 10, False
        q = random_sets(rng, 400, n_bits, mean=2.5, max_len=12, p_empty=0.02)
        p = random_sets(rng, 5000, n_bits, mean=2.5, max_len=12, p_empty=0.02)
        for i in range(0, 400, 7):             # plant real matches: with 50 000 ids random sets hardly ever intersect
            q[i] = list(p[(i * 13) % 5000]) + q[i][:1]
    elif case == "multi_group":                 # > 20 480 set bits in one 8 192-row batch: the pool is streamed per group
        n_bits, k, zd = 20000, 10, False
        q = random_sets(rng, 8192, n_bits, mean=5, max_len=7)
        p = random_sets(rng, 3000, n_bits, mean=5, max_len=7)
    elif case == "k32_hot":                     # k = 32 (the largest list) with lists that overflow all the time
        n_bits, k, zd = 64, 32, False
        q = random_sets(rng, 200, n_bits, mean=3, max_len=8)
        p = random_sets(rng, 20000, n_bits, mean=3, max_len=8)
    elif case == "all_empty_queries":
        n_bits, k, zd = 300, 5, False
        q = [[] for _ in range(130)]
        p = random_sets(rng, 500, n_bits, mean=3, p_empty=0.3)
    else:
        raise AssertionError(case)
    return n_bits, k, zd, q, p


@pytest.mark.parametrize("case", ["hot_small_universe", "duplicates", "multiword", "batches", "mixed_dense_tiles",
                                  "tiny_pool", "zero_diag_self", "all_empty_queries", "wide_vocab_shared_buckets",
                                  "multi_group", "k32_hot", "hot_no_overflow"])
def test_query_index_path_hard_cases(case):
    """The query-index kernel (jaccard_sparse.cu) against the oracle and against the dense-bitset kernel."""
    from rag4dyg_b200 import _lib
    n_bits, k, zd, q, p = _index_case(case)
    bq, bp = encode(q, n_bits), encode(p, n_bits)
    got = engine.jaccard_topk(bq, bp, k, zero_diag=zd)
    _lib.set_option("jaccard_sparse_q", 0)
    try:
        dense = engine.jaccard_topk(bq, bp, k, zero_diag=zd)
    finally:
        _lib.set_option("jaccard_sparse_q", 1)
    for a, b in zip(got, dense):
        assert torch.equal(a, b)
    oi, ou, ox = jo.c_topk(*to_csr(q), *to_csr(p), k, zero_diag=zd)
    assert np.array_equal(got[2].cpu().numpy(), ox)
    assert np.array_equal(got[0].cpu().numpy().astype(np.int64), oi)
    assert np.array_equal(got[1].cpu().numpy().astype(np.int64), ou)
