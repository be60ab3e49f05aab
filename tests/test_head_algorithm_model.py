"""A plain-Python model of postings_head_kernel's ranking argument (csrc/jaccard_postings.cu), checked against the oracle:

    a pool row holding exactly ONE of the query's m ids scores 1 / (m + |pool set| - 1), so among single-hit rows the
    order is (|pool set| asc, row asc) = the order of the per-id best lists; the top-K over single-hit rows is the merge
    of the (at most HEAD-entry) list heads with multi-hit and forced-zero rows removed, PROVIDED every head that lost
    entries still holds k of them or is the whole list; multi-hit rows are ranked on their own with their exact counts.

The model follows the kernel's decisions (hand over when a head is depleted) but none of its machinery (no filter, no
probes: multi-hit rows come from exact counting), so it pins the ALGORITHM on the CPU — random sets, repeat-heavy pools,
forced-zero diagonals, heads depleted on purpose.  The kernels themselves are checked in tests/test_gpu_postings.py."""
from fractions import Fraction

import numpy as np
import pytest

from conftest import random_sets, to_csr
from oracle import jaccard_oracle as jo

HEAD = 32          # PJ_BEST: entries per best list
IDX_NONE = 0x7FFFFFFF


def build_index(pool):
    """id -> posting list sorted by (|pool set|, row): the order postings_best_kernel keeps the heads in."""
    sets = [set(x) for x in pool]
    lists = {}
    for row, s in enumerate(sets):
        for t in s:
            lists.setdefault(t, []).append((len(s), row))
    return sets, {t: sorted(v) for t, v in lists.items()}


def head_topk(query, sets, lists, k, diag=None):
    """Returns ([(inter, union, row)] * <= k in canonical order, handed_over)."""
    ids = set(query)
    m = len(ids)
    hits = {}
    for t in ids:
        for _, row in lists.get(t, []):
            hits[row] = hits.get(row, 0) + 1
    multi = {row: c for row, c in hits.items() if c >= 2 and row != diag}
    cands = []
    for t in ids:
        full = lists.get(t, [])
        alive = [(card, row) for card, row in full[:HEAD] if row not in multi and row != diag]
        if len(alive) < k and len(full) > HEAD:
            return None, True                     # a depleted head cannot vouch for its id's k best single-hit rows
        cands += [(Fraction(1, m + card - 1), 1, m + card - 1, row) for card, row in alive[:k]]
    cands += [(Fraction(c, m + len(sets[row]) - c), c, m + len(sets[row]) - c, row) for row, c in multi.items()]
    cands.sort(key=lambda e: (-e[0], e[3]))
    return [(c, u, row) for _, c, u, row in cands[:k]], False


def check(queries, pool, k, zero_diag=False):
    sets, lists = build_index(pool)
    oi, ou, ox = jo.c_topk(*to_csr(queries), *to_csr(pool), k, zero_diag=zero_diag)
    served = 0
    for qi, q in enumerate(queries):
        got, handed = head_topk(q, sets, lists, k, diag=qi if zero_diag else None)
        if handed:
            continue
        served += 1
        # the oracle's list continues with zero-score fillers (lowest indices) / padding: compare the scoring prefix
        n_pos = int((oi[qi] > 0).sum())
        assert len(got) >= min(n_pos, k) and len(got) <= k
        assert [g[2] for g in got[:n_pos]] == ox[qi, :n_pos].tolist(), (qi, q)
        assert [g[0] for g in got[:n_pos]] == oi[qi, :n_pos].tolist() and [g[1] for g in got[:n_pos]] == ou[qi, :n_pos].tolist()
        assert len(got) == n_pos or n_pos == k    # the model lists positives only
    return served


@pytest.mark.parametrize("npool,n_bits,mean,k", [(3000, 2000, 2.2, 10), (2000, 60, 3, 10), (4000, 300, 5, 16), (500, 40, 2, 1)])
def test_head_ranking_equals_the_oracle(npool, n_bits, mean, k):
    rng = np.random.default_rng(npool + k)
    pool = random_sets(rng, npool, n_bits, mean=mean, max_len=min(32, n_bits), p_empty=0.05, dup=True)
    queries = random_sets(rng, 150, n_bits, mean=mean, max_len=min(32, n_bits), p_empty=0.05, dup=True)
    assert check(queries, pool, k) > 100
    assert check([list(x) for x in pool[:150]], pool, k, zero_diag=True) > 50   # forced-zero rows sit inside the heads


def test_depleted_heads_are_handed_over_not_guessed():
    # 40 rows {0, 1} head id 0's list (smallest sets), 400 longer rows follow: for the query {0, 1} every head entry of
    # id 0 is a multi-hit row, so the head says nothing about id 0's single-hit rows
    pool = [[0, 1] for _ in range(40)] + [[0, 5 + i % 300, 6 + i % 290] for i in range(400)] + [[1, 7 + i % 200] for i in range(25)]
    sets, lists = build_index(pool)
    assert head_topk([0, 1], sets, lists, 10)[1] is True
    # ... while queries that leave the head enough entries are served, and exactly
    assert check([[0], [1], [0, 399], [1, 8]], pool, 10) == 4
