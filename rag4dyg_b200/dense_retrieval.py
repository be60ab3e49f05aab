"""Drop-in for the dense scoring / ranking block of the retriever's test() (train/train_retriever.py:425-456) and
its file writer save_index_score (train/train_retriever.py:357-368), computed by libr4d.so on a B200.

Reference block (per eval batch):
    h_egos_norm = h_egos / h_egos.norm(dim=1, keepdim=True)                     :433
    train_embeddings_norm = train_embeddings / train_embeddings.norm(...)       :436   (redone every batch)
    dot_products = (h_egos_norm @ train_embeddings_norm.t() + 1) / 2            :437-438
    np.argsort(-dot_products, axis=1)  -> {val,test}_index.gen / _score.gen     :358, :455-456
Here the pool is normalised ONCE into the scorer's bf16 layout (DenseIndex) and stays in HBM; queries are scored by
the tcgen05 kernel; rankings come from the device radix ranker (full files) or the fused top-K (demonstrations).
The optional exp(-lambda*|dt|) query-time factor restates CLtime_loss (train/train_retriever.py:50-55); the
reference has no inference-time decay, so it is off by default.
"""
import os

import numpy as np
import torch

from . import engine, writers
from .engine import DENSE_COS_DECAY, DENSE_HALF_COS, DENSE_HALF_COS_DECAY, PREC_BF16, PREC_BF16X3  # noqa: F401


class DenseIndex:
    """Pool embeddings prepared once: L2-normalised bf16 planes resident in HBM (+ optional query times)."""

    def __init__(self, train_embeddings, times=None, prec=PREC_BF16X3):
        if not train_embeddings.is_cuda:
            raise engine.R4DError("DenseIndex needs CUDA embeddings (no CPU path)")
        self.prec = prec
        self.planes = engine.dense_prepare(train_embeddings.float().contiguous(), prec)
        self.times = None if times is None else times.to(train_embeddings.device, torch.float32).contiguous()

    @property
    def n(self):
        return self.planes.n_rows

    def _q(self, h_egos):
        return engine.dense_prepare(h_egos.float().contiguous(), self.prec)

    def scores(self, h_egos, mode=DENSE_HALF_COS, q_time=None, lam=0.0):
        """[B, N] float32 score block == dot_products of :437-438 (mode 0)."""
        return engine.dense_full(self._q(h_egos), self.planes, mode, q_time, self.times, lam)

    def topk(self, h_egos, k, mode=DENSE_HALF_COS, q_time=None, lam=0.0):
        """Top-k demonstrations per query: (scores f32 [B,k], idx int32 [B,k]), order (score desc, idx asc)."""
        return engine.dense_topk(self._q(h_egos), self.planes, k, mode, q_time, self.times, lam)


def score_block(h_egos, train_embeddings, prec=PREC_BF16X3):
    """One-shot equivalent of train/train_retriever.py:433-438 for CUDA tensors."""
    return DenseIndex(train_embeddings, prec=prec).scores(h_egos)


def save_index_score(score_matrix, save_index_file, save_score_file, steps):
    """train/train_retriever.py:357-368: full descending ranking + "%.4f" scores; 'w' on the first batch, 'a' after.
    score_matrix: float32 numpy array or CUDA tensor [B, N]."""
    if isinstance(score_matrix, torch.Tensor):
        s = score_matrix.to("cuda", torch.float32).contiguous()
    else:
        s = torch.from_numpy(np.ascontiguousarray(score_matrix, dtype=np.float32)).to("cuda")
    order = engine.rank_rows(s)
    mode = "w" if steps == 0 else "a"
    writers.write_int_rows(save_index_file, order.cpu().numpy(), mode)
    writers.write_float_rows(save_score_file, s.cpu().numpy(), writers.fmt_4f, mode)


def hit_rate_at_k(predictions, targets, k=1):
    """train/train_retriever.py:31-38."""
    return 1 if set(predictions[:k]) & set(targets) else 0


def evaluate_batches(index, query_embeddings, jaccard_rows, dataset, evaluate=True, batch_size=32, write=True,
                     root="."):
    """The per-batch loop of test() after the embeddings exist (:425-474): score, write .gen files, hit@1/@3.
    GT = first 3 of the stable descending ranking of the Jaccard row (:460-461, canonical ties)."""
    out_dir = os.path.join(root, "resources", "retrieval_result", dataset)
    if write:
        os.makedirs(out_dir, exist_ok=True)
    tag = "val" if evaluate else "test"
    hit1 = hit3 = 0.0
    steps = 0
    for b0 in range(0, query_embeddings.shape[0], batch_size):
        h = query_embeddings[b0:b0 + batch_size]
        s = index.scores(h)
        if write:
            save_index_score(s, os.path.join(out_dir, f"{tag}_index.gen"), os.path.join(out_dir, f"{tag}_score.gen"),
                             steps)
        pred = engine.rank_rows(s)[:, :3].cpu().numpy()
        gt_rows = torch.as_tensor(jaccard_rows[b0:b0 + batch_size], dtype=torch.float32).to("cuda").contiguous()
        gt = engine.rank_rows(gt_rows)[:, :3].cpu().numpy()
        n = h.shape[0]
        hit1 += sum(hit_rate_at_k(pred[i], gt[i], 1) for i in range(n)) / n
        hit3 += sum(hit_rate_at_k(pred[i], gt[i], 3) for i in range(n)) / n
        steps += 1
    return round(hit1 / max(steps, 1), 4), round(hit3 / max(steps, 1), 4)
