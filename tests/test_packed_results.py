"""Host half of the packed result lists (r4d_jaccard_topk_postings_packed, include/r4d.h): the decode a consumer runs on
pair = inter << 16 | |pool set|, idx, q_card = |query set| must give back the (inter, union, idx) planes of
r4d_jaccard_topk_postings, i.e. the operands of the reference's score len(a & b) / len(a | b)
(retrieval_data_annotation.py:12-15).  CPU only; the kernels' side is tests/test_gpu_postings.py."""
import numpy as np
import torch

from conftest import random_sets, to_csr
from oracle import jaccard_oracle as jo
from rag4dyg_b200 import engine

IDX_NONE = 0x7FFFFFFF


def pack_like_the_kernel(inter, union, idx, q_card):
    """pj_finish (csrc/jaccard_postings.cu): |pool set| = union + inter - |query set|; a padding entry packs as 0."""
    card = union + inter - q_card[:, None]
    pair = np.where(idx == IDX_NONE, 0, (inter << 16) | card)
    return pair.astype(np.uint32).view(np.int32)


def test_unpack_restores_the_oracle_planes():
    rng = np.random.default_rng(5)
    n_bits, k = 300, 10
    p = random_sets(rng, 400, n_bits, mean=4, p_empty=0.1, dup=True)
    q = random_sets(rng, 60, n_bits, mean=4, p_empty=0.1, dup=True)
    for pool in (p, p[:6], [[], [], []]):                   # a pool shorter than k pads with (0, 1, IDX_NONE)
        oi, ou, ox = jo.c_topk(*to_csr(q), *to_csr(pool), k)
        q_card = np.array([len({t for t in row if 0 <= t < n_bits}) for row in q], np.int64)
        pair = pack_like_the_kernel(oi.astype(np.int64), ou.astype(np.int64), ox.astype(np.int64), q_card)
        ui, uu, ux = engine.unpack_topk(torch.as_tensor(pair), torch.as_tensor(ox.astype(np.int32)),
                                        torch.as_tensor(q_card.astype(np.int32)))
        assert np.array_equal(ux.numpy(), ox) and np.array_equal(ui.numpy(), oi) and np.array_equal(uu.numpy(), ou)


def test_unpack_handles_counts_that_reach_the_sign_bit():
    """inter and |pool set| go up to 65 535 (n_bits <= 65 535): the packed word then has its top bit set."""
    inter = np.array([[65535, 40000, 1, 0]], np.int64)
    card = np.array([[65535, 50000, 65535, 0]], np.int64)
    q_card = np.array([65535], np.int64)
    idx = np.array([[3, 9, 11, IDX_NONE]], np.int32)
    pair = np.where(idx == IDX_NONE, 0, (inter << 16) | card).astype(np.uint32).view(np.int32)
    ui, uu, ux = engine.unpack_topk(torch.as_tensor(pair), torch.as_tensor(idx), torch.as_tensor(q_card.astype(np.int32)))
    assert ui.tolist() == [[65535, 40000, 1, 0]]
    assert uu.tolist() == [[65535, 65535 + 50000 - 40000, 65535 + 65535 - 1, 1]]
    assert ux.tolist() == idx.tolist()
