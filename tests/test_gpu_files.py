"""GPU parity at file level: the drop-in CLI / functions reproduce the reference's eight output files byte for byte
(golden sha256 from the unmodified reference, stable ties, np.random.seed(0))."""
import os

import numpy as np
import pytest

from conftest import DATASETS
from oracle import jaccard_oracle as jo
from test_oracle_golden import check_outputs, sha256_file

pytestmark = pytest.mark.gpu

from rag4dyg_b200 import retrieval_data_annotation as rda  # noqa: E402


@pytest.mark.parametrize("ds", ["UCI_13", "hepth", "dialog"])
def test_cli_reproduces_reference_files(ds, dataset_dir, manifest, monkeypatch, capsys):
    root = dataset_dir(ds)
    monkeypatch.chdir(root)
    np.random.seed(0)
    rda.main(["retrieval_data_annotation.py", ds, DATASETS[ds], "0.8"])
    out = capsys.readouterr().out
    n_lines = manifest[ds]["files"][f"resources/{ds}/{DATASETS[ds]}/train_retrieval/train_index.retrieval"]["lines"]
    assert f"Number of positive samples: {n_lines}" in out and "Done!" in out
    check_outputs(root, ds, manifest)


def test_function_level_drop_in_uci(dataset_dir, manifest, monkeypatch):
    """The reference's module-level functions, called the way its __main__ calls them (:162-198)."""
    ds, T = "UCI_13", "12"
    root = dataset_dir(ds)
    monkeypatch.chdir(root)
    base = os.path.join("resources", ds, T)
    train = rda._read_lines(os.path.join(base, "train.link_prediction"))
    test = rda._read_lines(os.path.join(base, "test.link_prediction"))
    test_gt = rda._read_lines(os.path.join(base, "test_gt.link_prediction"))
    tin, tout = rda.get_inout_list(train, train)
    _, te_out = rda.get_inout_list(test, test_gt)
    assert (tin, tout) == jo.get_inout_list(train, train)
    m_out = rda.occurrence_matrix(tout, tout)
    m_in = rda.occurrence_matrix(tin, tin)
    m_te = rda.occurrence_matrix(te_out, tout)
    assert m_out.dtype == np.float64 and m_out.shape == (1708, 1708)
    np.fill_diagonal(m_out, 0)
    np.fill_diagonal(m_in, 0)
    os.makedirs("o", exist_ok=True)
    rda.dataset = ds   # the reference reads a module global inside save_train_annotation (:73)
    np.random.seed(0)
    rda.save_train_annotation(m_out, m_in, "o/train_index.retrieval", "o/train_score.retrieval", threshold=0.8, neg_num=5)
    rda.save_index_score(m_te, "o/test_index.retrieval", "o/test_score.retrieval")
    rda.save_score_file_train(m_out, "o/train_index.gen", "o/train_score.gen", topk=10)
    files = manifest[ds]["files"]
    for name in ["train_index.retrieval", "train_score.retrieval", "test_index.retrieval", "test_score.retrieval"]:
        assert sha256_file(f"o/{name}") == files[f"resources/{ds}/{T}/train_retrieval/{name}"]["sha256"], name
    for name in ["train_index.gen", "train_score.gen"]:
        assert sha256_file(f"o/{name}") == files[f"resources/train_generator/{ds}/{T}/train_gt_topk/{name}"]["sha256"], name


def test_co_occurrence_ratio_semantics():
    assert rda.co_occurrence_ratio(["a", "b", "b"], ["b", "c"]) == 1 / 3
    assert rda.co_occurrence_ratio([], ["b"]) == 0 and rda.co_occurrence_ratio(None, ["b"]) == 0
    assert rda.co_occurrence_ratio(["b"], "b") == 1.0      # non-list seq_j is wrapped (:6-7)
