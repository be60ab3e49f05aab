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
