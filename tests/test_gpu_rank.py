"""GPU parity: stable radix ranking, top-k of explicit rows, triplet mining vs numpy (bit-exact index work)."""
import numpy as np
import pytest
import torch

from oracle import jaccard_oracle as jo

pytestmark = pytest.mark.gpu

from rag4dyg_b200 import engine  # noqa: E402


def tied_matrix(rng, nq, n, dtype):
    # few distinct values -> long tie runs, plus exact zeros (Jaccard-like)
    vals = np.array([0.0, 0.0, 0.0, 1.0, 0.5, 1 / 3, 2 / 3, 0.25, 0.06666666666666667, 1e-9, 0.9999999], dtype=dtype)
    return vals[rng.integers(0, len(vals), size=(nq, n))]


@pytest.mark.parametrize("nq,n", [(1, 1), (3, 31), (7, 256), (5, 2048), (4, 2049), (9, 7464), (2, 20011)])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_rank_rows_equals_stable_argsort(nq, n, dtype):
    rng = np.random.default_rng(nq * 7 + n)
    m = tied_matrix(rng, nq, n, dtype)
    got = engine.rank_rows(torch.from_numpy(m).cuda()).cpu().numpy()
    assert np.array_equal(got, np.argsort(-m, axis=1, kind="stable"))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_rank_rows_random_negative_nan(dtype):
    rng = np.random.default_rng(3)
    m = rng.standard_normal((6, 5000)).astype(dtype)
    m[0, 17] = np.nan
    m[1, :] = 0.0
    m[2, ::2] = -0.0
    got = engine.rank_rows(torch.from_numpy(m).cuda()).cpu().numpy()
    assert np.array_equal(got, np.argsort(-m, axis=1, kind="stable"))


@pytest.mark.parametrize("k", [1, 10, 32])
def test_topk_rows(k):
    rng = np.random.default_rng(k)
    m = tied_matrix(rng, 50, 3001, np.float64)
    ts, ti = engine.topk_rows(torch.from_numpy(m).cuda(), k)
    order, vals = jo.topk_stable(m, k)
    assert np.array_equal(ti.cpu().numpy(), order) and np.array_equal(ts.cpu().numpy(), vals)


def ref_mine(out_m, in_m, thr, neg_num):
    n = out_m.shape[0]
    n_pos = np.zeros(n, np.int32)
    negs = []
    for i in range(n):
        pos = set(np.where(out_m[i] > thr)[0].tolist())
        n_pos[i] = len(pos)
        order = np.argsort(-in_m[i], kind="stable")
        lst = [j for j in order if j not in pos and out_m[i, j] > 0][:neg_num]
        if len(lst) < neg_num:
            lst += [j for j in order if j not in pos and out_m[i, j] == 0][: neg_num - len(lst)]
        negs.append(lst)
    return n_pos, negs


@pytest.mark.parametrize("n,neg_num", [(64, 5), (333, 5), (1000, 3)])
def test_triplet_mining(n, neg_num):
    rng = np.random.default_rng(n)
    out_m = tied_matrix(rng, n, n, np.float64)
    in_m = tied_matrix(rng, n, n, np.float64)
    out_m[1, :] = 1.0          # everything positive -> no negatives at all
    out_m[2, :] = 0.0          # no positives
    out_m[3, :] = 0.0
    out_m[3, 5] = 1.0          # no hard negatives -> fill only
    np.fill_diagonal(out_m, 0)
    np.fill_diagonal(in_m, 0)
    n_pos, neg, n_neg = engine.triplet_mine(torch.from_numpy(out_m).cuda(), torch.from_numpy(in_m).cuda(), 0.8, neg_num)
    rp, rn = ref_mine(out_m, in_m, 0.8, neg_num)
    assert np.array_equal(n_pos.cpu().numpy(), rp)
    neg, n_neg = neg.cpu().numpy(), n_neg.cpu().numpy()
    for i in range(n):
        assert n_neg[i] == len(rn[i])
        assert neg[i, : n_neg[i]].tolist() == rn[i], i


@pytest.mark.parametrize("n,vo,vi,mean_o,mean_i,thr,neg_num", [
    (50, 40, 60, 2.0, 6.0, 0.8, 5),        # one anchor block + a ragged tile
    (333, 120, 300, 2.2, 8.0, 0.8, 5),
    (1000, 300, 2500, 1.5, 20.0, 0.8, 3),  # several 64-word chunks of history sets
    (700, 30, 80, 2.0, 5.0, 0.3, 5),       # small universe: many positives and ties
    (257, 64, 64, 3.0, 3.0, 0.0, 7),       # thr = 0: every intersecting pair is a positive -> fill negatives only
])
def test_triplet_mining_from_bitsets(n, vo, vi, mean_o, mean_i, thr, neg_num):
    """r4d_triplet_mine (OUT / IN bitsets, no matrix) == the reference's walk on the float64 matrices (:53-71)."""
    from conftest import random_sets, to_csr
    from rag4dyg_b200 import set_encoder
    rng = np.random.default_rng(n + neg_num)
    s_out = random_sets(rng, n, vo, mean=mean_o, max_len=min(64, vo), p_empty=0.03, dup=True)
    s_in = random_sets(rng, n, vi, mean=mean_i, max_len=min(64, vi), p_empty=0.01, dup=True)
    s_out[5] = list(s_out[4])                  # identical label sets: score 1.0 both ways
    out_m = jo.scores_from_counts(*jo.counts_matrix(s_out, s_out))
    in_m = jo.scores_from_counts(*jo.counts_matrix(s_in, s_in))
    np.fill_diagonal(out_m, 0)
    np.fill_diagonal(in_m, 0)
    b_out = set_encoder.encode_csr(*to_csr(s_out), vo)
    b_in = set_encoder.encode_csr(*to_csr(s_in), vi)
    for cap in (None, 7):                      # 7: the positives overflow the first buffer and the call is repeated
        m = engine.triplet_mine_bits(b_out, b_in, thr, neg_num, zero_diag=True, pos_cap=cap)
        rp, rn = ref_mine(out_m, in_m, thr, neg_num)
        assert np.array_equal(m["n_pos"].cpu().numpy(), rp)
        rows, cols = np.nonzero(out_m > thr)       # row-major, like np.where per row
        assert np.array_equal(m["pos_row"].cpu().numpy(), rows) and np.array_equal(m["pos_col"].cpu().numpy(), cols)
        pi, pu = m["pos_inter"].cpu().numpy(), m["pos_union"].cpu().numpy()
        assert np.array_equal(pi / pu, out_m[rows, cols])                      # bit-identical float64 scores
        neg, n_neg = m["neg"].cpu().numpy(), m["n_neg"].cpu().numpy()
        ni, nu = m["neg_inter"].cpu().numpy(), m["neg_union"].cpu().numpy()
        for i in range(n):
            assert n_neg[i] == len(rn[i])
            assert neg[i, : n_neg[i]].tolist() == rn[i], i
            assert np.array_equal(ni[i, : n_neg[i]] / nu[i, : n_neg[i]], out_m[i, rn[i]])
            assert (neg[i, n_neg[i]:] == -1).all()
