"""GPU tests of the SURVEY.md 8f "next" rows: pool-embedding producer, device-side triplet sampling."""
import numpy as np
import pytest
import torch

from oracle import dense_oracle as do

pytestmark = pytest.mark.gpu

from rag4dyg_b200 import engine  # noqa: E402
from rag4dyg_b200 import retrieval_data_annotation as rda  # noqa: E402


@pytest.mark.parametrize("b,l,d", [(32, 512, 512), (5, 37, 100), (64, 128, 768), (1, 1, 64)])
def test_meanpool_prepare_matches_torch_mean_then_prepare(b, l, d):
    g = torch.Generator().manual_seed(b * l + d)
    h = torch.randn(b, l, d, generator=g) + 0.5
    planes, mean = engine.meanpool_prepare(h.cuda(), engine.PREC_BF16X3, want_mean=True)
    ref_mean = torch.mean(h, dim=1)                       # train/train_retriever.py:420 (pads included)
    assert torch.allclose(mean.cpu(), ref_mean, rtol=0, atol=2e-6)
    # planes == what r4d_dense_prepare gives for that mean (same normalisation + split); compare reconstructed rows
    want = engine.dense_prepare(mean, engine.PREC_BF16X3)
    assert torch.equal(planes.hi, want.hi) and torch.equal(planes.lo, want.lo)
    recon = planes.hi.float()[:, :d] + planes.lo.float()[:, :d]
    ref = ref_mean / ref_mean.norm(dim=1, keepdim=True)
    assert (recon.cpu() - ref).abs().max() <= 1e-5


def test_meanpool_feeds_scorer_end_to_end():
    g = torch.Generator().manual_seed(3)
    hq, hp = torch.randn(40, 64, 256, generator=g) + 0.3, torch.randn(900, 64, 256, generator=g) + 0.3
    q, _ = engine.meanpool_prepare(hq.cuda(), engine.PREC_BF16X3)
    p, _ = engine.meanpool_prepare(hp.cuda(), engine.PREC_BF16X3)
    got = engine.dense_full(q, p).cpu().numpy()
    ref = do.score_block(hq.mean(1), hp.mean(1)).numpy()  # :420, :433-438
    assert np.abs(got - ref).max() <= 1e-5


def test_counter_sampler_is_deterministic_and_valid(tmp_path):
    rng = np.random.default_rng(0)
    n = 400
    vals = np.array([0.0, 0.0, 0.0, 1.0, 0.5, 1 / 3, 0.9])
    out_m, in_m = vals[rng.integers(0, len(vals), (n, n))], vals[rng.integers(0, len(vals), (n, n))]
    np.fill_diagonal(out_m, 0)
    np.fill_diagonal(in_m, 0)
    rda.dataset = "UCI_13"
    files = []
    for rep in range(2):
        fi, fs = str(tmp_path / f"i{rep}"), str(tmp_path / f"s{rep}")
        rda.save_train_annotation(out_m, in_m, fi, fs, threshold=0.8, neg_num=5, sampler="counter", seed=123)
        files.append((open(fi).read(), open(fs).read()))
    assert files[0] == files[1], "same seed -> identical files"
    fi2, fs2 = str(tmp_path / "i2"), str(tmp_path / "s2")
    rda.save_train_annotation(out_m, in_m, fi2, fs2, threshold=0.8, neg_num=5, sampler="counter", seed=124)
    assert open(fi2).read() != files[0][0], "different seed -> different negatives"
    # same anchors/positives as the reference rule; every negative comes from the mined candidate list
    n_pos, neg, n_neg = engine.triplet_mine(torch.from_numpy(out_m).cuda(), torch.from_numpy(in_m).cuda(), 0.8, 5)
    neg, n_neg = neg.cpu().numpy(), n_neg.cpu().numpy()
    trip = np.array([[int(x) for x in ln.split()] for ln in files[0][0].splitlines()])
    pos_ref = np.argwhere(out_m > 0.8)
    assert np.array_equal(trip[:, :2], pos_ref)
    for i, _, ng in trip:
        assert ng in neg[i, : n_neg[i]]
    for ln_i, ln_s in zip(files[0][0].splitlines(), files[0][1].splitlines()):
        i, p, ng = map(int, ln_i.split())
        assert ln_s == f"{i} {out_m[i, p]} {out_m[i, ng]}"
