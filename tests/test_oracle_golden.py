"""CPU: pin the oracle (oracle/jaccard_oracle.py, .c) against golden vectors produced by the unmodified reference
(oracle/make_golden.py: stable argsort + np.random.seed(0)).  sha256 values equal SURVEY.md section 8c."""
import hashlib
import lzma
import os

import numpy as np
import pytest

from conftest import DATASETS, GOLD, random_sets, to_csr
from oracle import jaccard_oracle as jo


def sha256_file(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def check_outputs(root, ds, manifest):
    for rel, info in manifest[ds]["files"].items():
        p = os.path.join(root, rel)
        assert os.path.exists(p), rel
        assert os.path.getsize(p) == info["bytes"], rel
        assert sha256_file(p) == info["sha256"], f"{ds}: {rel} differs from the reference golden"


def _run_oracle(ds, dataset_dir, manifest):
    root = dataset_dir(ds)
    n_pos = jo.annotate(ds, DATASETS[ds], 0.8, root=root, seed=0)
    assert n_pos == manifest[ds]["files"][f"resources/{ds}/{DATASETS[ds]}/train_retrieval/train_index.retrieval"]["lines"]
    check_outputs(root, ds, manifest)


def test_oracle_reproduces_reference_files_uci(dataset_dir, manifest):
    _run_oracle("UCI_13", dataset_dir, manifest)


def test_oracle_reproduces_reference_files_hepth(dataset_dir, manifest):
    _run_oracle("hepth", dataset_dir, manifest)


@pytest.mark.skipif(os.environ.get("R4D_SLOW") != "1", reason="dialog oracle run takes minutes; set R4D_SLOW=1")
def test_oracle_reproduces_reference_files_dialog(dataset_dir, manifest):
    _run_oracle("dialog", dataset_dir, manifest)


def test_known_answers(manifest):
    # positives written; first three UCI triplets share anchor 0 with positives 34, 132, 167, score 1.0
    assert manifest["UCI_13"]["files"]["resources/UCI_13/12/train_retrieval/train_index.retrieval"]["lines"] == 9578
    assert manifest["hepth"]["files"]["resources/hepth/11/train_retrieval/train_index.retrieval"]["lines"] == 8250
    assert manifest["dialog"]["files"]["resources/dialog/15/train_retrieval/train_index.retrieval"]["lines"] == 10762
    idx = lzma.open(os.path.join(GOLD, "UCI_13", "train_index.retrieval.xz"), "rt").read().splitlines()[:3]
    sc = lzma.open(os.path.join(GOLD, "UCI_13", "train_score.retrieval.xz"), "rt").read().splitlines()[:3]
    assert [ln.split()[:2] for ln in idx] == [["0", "34"], ["0", "132"], ["0", "167"]]
    assert all(ln.split()[1] == "1.0" for ln in sc)


def test_set_loop_equals_integer_formulation_on_real_lines(dataset_dir):
    root = dataset_dir("UCI_13")
    base = os.path.join(root, "resources", "UCI_13", "12")
    train = jo._read(os.path.join(base, "train.link_prediction"))[:120]
    tin, tout = jo.get_inout_list(train, train)
    for seqs in (tin, tout):
        ref = jo.occurrence_matrix(seqs[:40], seqs)           # Python sets, reference :5-15, :36-41
        fast = jo.scores_from_counts(*jo.counts_matrix(seqs[:40], seqs))
        assert np.array_equal(ref, fast)
    assert any("time" in t for s in tin for t in s), "history sets keep <|timeN|> markers (SURVEY fact 5)"
    assert not any("time" in t for s in tout for t in s)


def test_c_oracle_equals_python_oracle():
    rng = np.random.default_rng(7)
    q = random_sets(rng, 37, 50, mean=3, p_empty=0.1, dup=True)
    p = random_sets(rng, 91, 50, mean=3, p_empty=0.1, dup=True)
    qs, ps = [list(map(str, s)) for s in q], [list(map(str, s)) for s in p]
    m = jo.occurrence_matrix(qs, ps)
    ci, cu = jo.c_counts(*to_csr(q), *to_csr(p))
    assert np.array_equal(jo.scores_from_counts(ci.astype(np.int64), cu.astype(np.int64)), m)
    ti, tu, tx = jo.c_topk(*to_csr(q), *to_csr(p), 10)
    order, vals = jo.topk_stable(m, 10)
    assert np.array_equal(tx, order)
    assert np.array_equal(ti / tu, vals)
    # zero_diag on a square problem == np.fill_diagonal(m, 0) before ranking
    m2 = jo.occurrence_matrix(ps, ps)
    np.fill_diagonal(m2, 0)
    ti, tu, tx = jo.c_topk(*to_csr(p), *to_csr(p), 5, zero_diag=True)
    order, vals = jo.topk_stable(m2, 5)
    assert np.array_equal(tx, order) and np.array_equal(ti / tu, vals)
