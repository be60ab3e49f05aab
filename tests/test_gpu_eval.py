"""GPU: the evaluation loop of test() after the embeddings exist (train/train_retriever.py:425-474) — score, write the
.gen files, hit@1/@3 — on the reference's own embeddings and the reference's own Jaccard rows (golden files)."""
import lzma
import os

import numpy as np
import pytest
import torch

from conftest import GOLD

pytestmark = pytest.mark.gpu

from rag4dyg_b200 import dense_retrieval as dr  # noqa: E402


def ref_hit_rates(scores, jaccard_rows, batch=32):
    """Restatement of :458-479 with the canonical (stable) argsort."""
    h1 = h3 = 0.0
    steps = 0
    for b0 in range(0, scores.shape[0], batch):
        s, j = scores[b0:b0 + batch], jaccard_rows[b0:b0 + batch].astype(np.float32)
        n = s.shape[0]
        b1 = b3 = 0
        for i in range(n):
            gt = np.argsort(-j[i], kind="stable")[:3]
            pred = np.argsort(-s[i], kind="stable")
            b1 += 1 if set(pred[:1]) & set(gt) else 0
            b3 += 1 if set(pred[:3]) & set(gt) else 0
        h1 += b1 / n
        h3 += b3 / n
        steps += 1
    return round(h1 / steps, 4), round(h3 / steps, 4)


def test_evaluate_batches_on_reference_embeddings(tmp_path):
    z = np.load(os.path.join(GOLD, "dense_UCI13.npz"))
    rows = lzma.open(os.path.join(GOLD, "UCI_13", "test_score.retrieval.xz"), "rt").read().splitlines()
    jac = np.array([[float(x) for x in ln.split()] for ln in rows])         # the score targets of :402-410
    assert jac.shape == (110, 1708)
    index = dr.DenseIndex(torch.from_numpy(z["pool_emb"]).cuda())
    q = torch.from_numpy(z["query_emb"]).cuda()
    hit1, hit3 = dr.evaluate_batches(index, q, jac, "UCI_13", evaluate=False, root=str(tmp_path))
    assert (hit1, hit3) == ref_hit_rates(z["ref_scores"], jac)
    out = tmp_path / "resources" / "retrieval_result" / "UCI_13"
    assert (out / "test_index.gen").exists() and (out / "test_score.gen").exists()
    assert len(open(out / "test_index.gen").read().splitlines()) == 110
