"""GPU property tests (hypothesis): set encoder + Jaccard against Python set algebra on adversarial small inputs
(empty sets, duplicate tokens, full-vocabulary sets, word-boundary vocab sizes)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from conftest import to_csr

pytestmark = pytest.mark.gpu

from rag4dyg_b200 import engine, set_encoder  # noqa: E402


@st.composite
def problem(draw):
    n_bits = draw(st.sampled_from([1, 2, 31, 32, 33, 64, 65, 255, 256, 257, 1000]))
    ids = st.integers(min_value=0, max_value=n_bits - 1)
    sets = st.lists(st.lists(ids, min_size=0, max_size=min(3 * n_bits, 70)), min_size=1, max_size=40)
    q, p = draw(sets), draw(sets)
    if draw(st.booleans()):
        p.append(list(range(n_bits)))          # a full-vocabulary row
    return n_bits, q, p, draw(st.integers(min_value=1, max_value=12))


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(problem())
def test_counts_and_topk_match_python_sets(prob):
    n_bits, q, p, k = prob
    bq = set_encoder.encode_csr(*to_csr(q), n_bits)
    bp = set_encoder.encode_csr(*to_csr(p), n_bits)
    inter, score = engine.jaccard_full(bq, bp)
    inter, score = inter.cpu().numpy(), score.cpu().numpy()
    ref = np.zeros((len(q), len(p)))
    for i, a in enumerate(q):
        for j, b in enumerate(p):
            sa, sb = set(a), set(b)
            assert inter[i, j] == len(sa & sb)
            ref[i, j] = len(sa & sb) / len(sa | sb) if sa and sb else 0      # retrieval_data_annotation.py:10-14
    assert np.array_equal(score, ref)
    ti, tu, tx = engine.jaccard_topk(bq, bp, k)
    order = np.argsort(-ref, axis=1, kind="stable")[:, :k]
    kk = min(k, len(p))
    assert np.array_equal(tx.cpu().numpy()[:, :kk], order[:, :kk])
    assert np.all(tx.cpu().numpy()[:, kk:] == engine.R4D_IDX_NONE)
    got = ti.cpu().numpy()[:, :kk] / np.maximum(tu.cpu().numpy()[:, :kk], 1)
    assert np.array_equal(got, np.take_along_axis(ref, order[:, :kk], axis=1))
