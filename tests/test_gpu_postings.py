"""GPU parity of the postings path (r4d_postings_build + r4d_jaccard_topk_postings) vs the CPU oracle and vs the
bitset path.  Bit-exact: integer counts, canonical (score desc, index asc) order."""
import numpy as np
import pytest
import torch

from conftest import random_sets, to_csr
from oracle import jaccard_oracle as jo

pytestmark = pytest.mark.gpu

from rag4dyg_b200 import engine, set_encoder  # noqa: E402


def run_postings(q, p, n_bits, k, zero_diag=False, query_base=0, pool_base=0):
    bp = set_encoder.encode_csr(*to_csr(p), n_bits)
    index = engine.build_postings(bp)
    qi, qo = to_csr(q)
    if qi.size == 0:
        qi = np.zeros(1, np.int32)
    dqi, dqo = torch.as_tensor(qi).cuda(), torch.as_tensor(qo).cuda()
    ti, tu, tx = engine.jaccard_topk_postings(dqi, dqo, index, k, zero_diag=zero_diag, query_base=query_base, pool_base=pool_base)
    # the packed form of the same call (8 bytes per entry) must decode to exactly these planes, whatever kernel served a row
    packed = engine.jaccard_topk_postings_packed(dqi, dqo, index, k, zero_diag=zero_diag, query_base=query_base,
                                                 pool_base=pool_base)
    ui, uu, ux = engine.unpack_topk(*packed)
    assert torch.equal(ux, tx) and torch.equal(ui, ti) and torch.equal(uu, tu), "packed results differ from the planes"
    return ti.cpu().numpy().astype(np.int64), tu.cpu().numpy().astype(np.int64), tx.cpu().numpy(), bp, index


def check(q, p, n_bits, k, zero_diag=False, query_base=0, pool_base=0):
    ti, tu, tx, bp, index = run_postings(q, p, n_bits, k, zero_diag, query_base, pool_base)
    oi, ou, ox = jo.c_topk(*to_csr(q), *to_csr(p), k, zero_diag=zero_diag, query_base=query_base, pool_base=pool_base)
    assert np.array_equal(tx, ox), f"indices differ in rows {np.nonzero((tx != ox).any(1))[0][:5]}"
    assert np.array_equal(ti, oi)
    assert np.array_equal(tu, ou)
    return bp, index


@pytest.mark.parametrize("nq,npool,n_bits,mean,k,zero_diag", [
    (3, 5, 40, 3, 10, False),            # pool shorter than k -> padded with R4D_IDX_NONE
    (130, 1000, 1000, 2.2, 10, False),
    (257, 3000, 20000, 2.2, 10, False),
    (300, 20000, 20000, 2.2, 10, False),  # three row windows
    (100, 2000, 20000, 20, 7, False),
    (500, 500, 300, 3, 10, True),        # train x train with the diagonal zeroed
    (64, 5000, 4749, 2.2, 32, False),
    (40, 700, 64, 12, 1, False),         # tiny vocabulary: long posting lists, several passes / heavy queries
    (200, 30000, 500, 3, 10, False),     # ~180 postings per id over 4 windows: multi-pass light queries
])
def test_postings_topk_bit_exact(nq, npool, n_bits, mean, k, zero_diag):
    rng = np.random.default_rng(nq + npool + k)
    p = random_sets(rng, npool, n_bits, mean=mean, max_len=min(64, n_bits), p_empty=0.05, dup=True)
    q = p[:nq] if zero_diag else random_sets(rng, nq, n_bits, mean=mean, max_len=min(64, n_bits), p_empty=0.05, dup=True)
    check(q, p, n_bits, k, zero_diag)


def test_postings_equals_bitset_path():
    rng = np.random.default_rng(21)
    n_bits = 20000
    p = random_sets(rng, 50000, n_bits, mean=2.2)
    q = random_sets(rng, 3000, n_bits, mean=2.2, p_empty=0.02, dup=True)
    ti, tu, tx, bp, _ = run_postings(q, p, n_bits, 10)
    bq = set_encoder.encode_csr(*to_csr(q), n_bits)
    ri, ru, rx = engine.jaccard_topk(bq, bp, 10)
    assert np.array_equal(tx, rx.cpu().numpy())
    assert np.array_equal(ti, ri.cpu().numpy()) and np.array_equal(tu, ru.cpu().numpy())


def test_postings_heavy_queries_and_hot_ids():
    """Queries with more than 64 ids, a hot id held by a third of the pool, and light queries next to them."""
    rng = np.random.default_rng(5)
    n_bits, npool = 3000, 20000
    p = random_sets(rng, npool, n_bits, mean=4, max_len=64)
    for i in range(0, npool, 3):
        p[i] = p[i] + [7]                      # hot id: ~6 700 postings
    q = random_sets(rng, 60, n_bits, mean=3, max_len=64, p_empty=0.05, dup=True)
    q[3] = list(rng.choice(n_bits, 200, replace=False))          # > 64 ids -> heavy kernel
    q[10] = [7, 7, 11]                                             # hot id -> several passes
    q[11] = list(range(0, 2600))                                   # > 2 048 distinct ids: heavy kernel without the id list
    q[12] = list(rng.choice(n_bits, 70, replace=True)) + [7]
    check(q, p, n_bits, 10)


def test_postings_hot_window_falls_back_exactly():
    """All postings of an id inside ONE row window: the light kernel's table overflows and hands the query over."""
    rng = np.random.default_rng(9)
    n_bits, npool = 5000, 9000
    p = random_sets(rng, npool, n_bits, mean=2.2)
    for i in range(100, 900):
        p[i] = [42] + p[i][:1]                 # 800 postings of id 42, all in window 0
    q = [[42], [42, 43], [1, 2, 3], []]
    check(q, p, n_bits, 10)


def test_postings_bases_and_diag_offsets():
    rng = np.random.default_rng(13)
    n_bits = 800
    p = random_sets(rng, 1200, n_bits, mean=3)
    check(p[100:400], p, n_bits, 10, zero_diag=True, query_base=1100, pool_base=1000)
    check(random_sets(rng, 50, n_bits, mean=3), p, n_bits, 5, pool_base=123456)


def test_postings_empty_inputs():
    n_bits = 100
    bp = set_encoder.encode_csr(*to_csr([[1, 2], [3]]), n_bits)
    index = engine.build_postings(bp)
    ti, tu, tx = engine.jaccard_topk_postings(torch.zeros(1, dtype=torch.int32).cuda(), torch.zeros(1, dtype=torch.int64).cuda(),
                                              index, 4)
    assert ti.shape == (0, 4)
    # a pool of empty sets: every score 0, fillers in index order
    check([[1], []], [[], [], []], n_bits, 2)


def test_postings_shard_invariance():
    rng = np.random.default_rng(17)
    n_bits, npool, k = 2000, 20001, 10
    p = random_sets(rng, npool, n_bits, mean=2.2)
    q = random_sets(rng, 150, n_bits, mean=2.2)
    oi, ou, ox = jo.c_topk(*to_csr(q), *to_csr(p), k)
    bounds = [0, 7000, 7001, 15000, npool]
    parts = []
    qi, qo = to_csr(q)
    dq, do = torch.as_tensor(qi).cuda(), torch.as_tensor(qo).cuda()
    for a, b in zip(bounds[:-1], bounds[1:]):
        index = engine.build_postings(set_encoder.encode_csr(*to_csr(p[a:b]), n_bits))
        parts.append(engine.jaccard_topk_postings(dq, do, index, k, pool_base=a))
    mi, mu, mx = engine.jaccard_topk_merge(torch.stack([x[0] for x in parts]), torch.stack([x[1] for x in parts]),
                                           torch.stack([x[2] for x in parts]), k)
    assert np.array_equal(mx.cpu().numpy(), ox)
    assert np.array_equal(mi.cpu().numpy(), oi) and np.array_equal(mu.cpu().numpy(), ou)


@pytest.mark.parametrize("zero_diag", [False, True])
def test_host_topk_row_ranges_equal_one_call(zero_diag):
    """HostTopK scores a step in row ranges (offsets stay absolute, query_base moves with the range) and overlaps their
    device->host copies: the result must equal one device-resident call, for two steps in flight."""
    from rag4dyg_b200.jaccard_pool import HostTopK, JaccardPool
    rng = np.random.default_rng(9)
    n_bits, npool, nq, k = 5000, 30000, 18000, 10
    p = random_sets(rng, npool, n_bits, mean=2.2, p_empty=0.02, dup=True)
    q = p[:nq] if zero_diag else random_sets(rng, nq, n_bits, mean=2.2, p_empty=0.02, dup=True)
    pool = JaccardPool.from_csr(*to_csr(p), n_bits)
    qi, qo = to_csr(q)
    tq, to = torch.as_tensor(qi).pin_memory(), torch.as_tensor(qo).pin_memory()
    ref = pool.topk(tq.cuda(), to.cuda(), k, zero_diag=zero_diag)
    # copy-stream ranges / kernel stores to pinned host / the same with packed lists
    for kw in (dict(direct=False, chunks=4), dict(direct=True), dict(direct=True, packed=True)):
        hk = HostTopK(pool, k, nq, qi.size, depth=2, **kw)
        t0 = hk.submit(tq, to, zero_diag=zero_diag)
        t1 = hk.submit(tq, to, zero_diag=zero_diag)
        for t in (t0, t1):
            got = hk.result(t)
            if hk.packed:
                assert got[2].shape == (nq,) and hk.bytes_per_step(nq, qi.size)[1] == nq * k * 8 + nq * 4
                got = HostTopK.unpack(got)
            assert all(torch.equal(g, r.cpu()) for g, r in zip(got, ref)), kw
    oi, ou, ox = jo.c_topk(qi, qo, *to_csr(p), k, zero_diag=zero_diag)
    assert np.array_equal(ref[2].cpu().numpy(), ox) and np.array_equal(ref[0].cpu().numpy(), oi)


def test_packed_results_from_every_kernel_of_the_chain():
    """Packed lists through the hash-table kernel as first stage (option postings_kernel = 1), with hand-overs to the
    heavy kernel, a pool shorter than k (padding entries) and empty-vs-empty pairs (union forced to 1)."""
    from rag4dyg_b200 import _lib
    rng = np.random.default_rng(33)
    n_bits = 600
    p = random_sets(rng, 4000, n_bits, mean=3, p_empty=0.1, dup=True)
    q = random_sets(rng, 300, n_bits, mean=3, p_empty=0.1, dup=True)
    q[5] = list(range(0, 500))                 # heavy kernel
    q[6] = list(rng.choice(n_bits, 100))       # > 64 ids
    _lib.set_option("postings_kernel", 1)
    try:
        check(q, p, n_bits, 10)
        check(q[:40], p[:6], n_bits, 10)       # padding entries
        check([[], [1]], [[], [], [2]], n_bits, 3)
    finally:
        _lib.set_option("postings_kernel", 0)
    check(q[:40], p[:6], n_bits, 10)
    check([[], [1]], [[], [], [2]], n_bits, 3)


def test_packed_lists_relayed_to_host_in_blocks():
    """Packed lists bound for pinned host memory leave the head kernel in whole 64-query blocks (relay through a
    staging copy in HBM): a last partial block, long queries served first, queries handed to the later stages (which
    write straight to the host after the relays), the forced-zero diagonal — equal to the device-resident call, with
    the relay on and off."""
    from rag4dyg_b200 import _lib
    from rag4dyg_b200.jaccard_pool import JaccardPool
    rng = np.random.default_rng(41)
    n_bits, npool, k = 3000, 20000, 10
    p = random_sets(rng, npool, n_bits, mean=2.5, max_len=40, p_empty=0.03, dup=True)
    for i in range(0, 3000, 2):
        p[i] = p[i] + [7]                       # a hot (id, window) bucket: probes of id 7 hand the query over
    for nq, zero_diag in ((1000 + 37, False), (64, False), (5, False), (2000, True)):
        q = [list(x) for x in p[:nq]] if zero_diag else random_sets(rng, nq, n_bits, mean=2.5, max_len=40, p_empty=0.03, dup=True)
        if not zero_diag:
            q[3] = list(rng.choice(n_bits, 60, replace=False))      # > 32 ids: handed over
            q[4] = [7, 8, 9]
            q[-1] = list(rng.choice(n_bits, 20, replace=False))     # a long query (served first)
        pool = JaccardPool.from_csr(*to_csr(q and p), n_bits)
        qi, qo = to_csr(q)
        dq, do = torch.as_tensor(qi).cuda(), torch.as_tensor(qo).cuda()
        ref = pool.topk(dq, do, k, zero_diag=zero_diag)
        oi, ou, ox = jo.c_topk(qi, qo, *to_csr(p), k, zero_diag=zero_diag)
        assert np.array_equal(ref[2].cpu().numpy(), ox) and np.array_equal(ref[0].cpu().numpy(), oi)
        for relay in (1, 0):
            prev = _lib.set_option("postings_relay", relay)
            try:
                host = (torch.full((nq, k), -1, dtype=torch.int32).pin_memory(), torch.full((nq, k), -1, dtype=torch.int32).pin_memory(),
                        torch.full((nq,), -1, dtype=torch.int32).pin_memory())
                pool.topk_packed(dq, do, k, zero_diag=zero_diag, out=host)
                torch.cuda.synchronize()
            finally:
                _lib.set_option("postings_relay", prev)
            got = engine.unpack_topk(*host)
            assert all(torch.equal(g, r.cpu()) for g, r in zip(got, ref)), (nq, zero_diag, relay)


def test_graph_topk_replays_with_refreshed_queries():
    """GraphTopK: the captured launch sequence re-run on new query CONTENTS (same buffers) equals a plain call."""
    from rag4dyg_b200.jaccard_pool import GraphTopK, JaccardPool
    rng = np.random.default_rng(17)
    n_bits, npool, nq, k = 3000, 20000, 2000, 10
    p = random_sets(rng, npool, n_bits, mean=2.2)
    pool = JaccardPool.from_csr(*to_csr(p), n_bits)
    qa = random_sets(rng, nq, n_bits, mean=2.2, max_len=8)
    qb = random_sets(rng, nq, n_bits, mean=2.2, max_len=8)
    cap = 8 * nq
    ids = torch.zeros(cap, dtype=torch.int32, device="cuda")
    off = torch.zeros(nq + 1, dtype=torch.int64, device="cuda")

    def load(q):
        qi, qo = to_csr(q)
        ids[:qi.size].copy_(torch.as_tensor(qi))
        off.copy_(torch.as_tensor(qo))
        return qi, qo
    qi, qo = load(qa)
    g = GraphTopK(pool, ids, off, k)
    for q in (qa, qb, qa):
        qi, qo = load(q)
        got = [t.cpu().numpy() for t in g.replay()]
        oi, ou, ox = jo.c_topk(qi, qo, *to_csr(p), k)
        assert np.array_equal(got[2], ox) and np.array_equal(got[0], oi) and np.array_equal(got[1], ou)


@pytest.mark.parametrize("npool,n_bits,mean,nq,k,self_queries", [
    (3000, 200, 6, 400, 10, False),      # small vocabulary: many rows share several ids (true repeats), 2-3 passes
    (30000, 30000, 20, 300, 10, True),   # queries = pool rows with up to 64 ids: one row repeats in EVERY list; > 32 ids
    (30000, 30000, 20, 300, 12, False),  # general list search (5..64 ids), staged output at the widest staged k
    (9000, 1500, 3, 2000, 7, True),      # two windows, forced-zero diagonal next to repeats
])
def test_register_kernel_repeats_and_passes(npool, n_bits, mean, nq, k, self_queries):
    """The register-resident light kernel (label-like regime): repeat filter + resolution, few-list and searched list
    lookup, several passes over row windows — bit-identical to the oracle and to the hash-table kernel."""
    from rag4dyg_b200 import _lib
    rng = np.random.default_rng(npool + k)
    p = random_sets(rng, npool, n_bits, mean=mean, max_len=min(64, n_bits), p_empty=0.02, dup=True)
    q = [list(x) for x in p[:nq]] if self_queries else random_sets(rng, nq, n_bits, mean=mean, max_len=min(64, n_bits),
                                                                 p_empty=0.02, dup=True)
    zero_diag = self_queries and k == 7
    ti, tu, tx, bp, index = run_postings(q, p, n_bits, k, zero_diag=zero_diag)
    oi, ou, ox = jo.c_topk(*to_csr(q), *to_csr(p), k, zero_diag=zero_diag)
    assert np.array_equal(tx, ox) and np.array_equal(ti, oi) and np.array_equal(tu, ou)
    for first_stage in (1, 2):        # the hash-table kernel / the register kernel as first stage of the same call
        prev = _lib.set_option("postings_kernel", first_stage)
        try:
            hi, hu, hx, _, _ = run_postings(q, p, n_bits, k, zero_diag=zero_diag)
        finally:
            _lib.set_option("postings_kernel", prev)
        assert np.array_equal(hx, tx) and np.array_equal(hi, ti) and np.array_equal(hu, tu), first_stage


def test_head_kernel_hand_overs_and_kills():
    """The head kernel (first stage for label-like sets, k <= 16): list heads that lose entries to multi-hit rows and
    to the forced-zero row, heads depleted below k (hand-over), more noted rows than it resolves (hand-over), long
    lists of a tiny vocabulary (many false positives of the filter), every k the merge network serves."""
    rng = np.random.default_rng(77)
    n_bits = 400
    # ids 0 (X) and 1 (Y): 40 rows {X, Y} head X's list (smallest sets) -> all of X's head is multi-hit for {X, Y}
    p = [[0, 1] for _ in range(40)]
    p += [[0, 5 + i % 300, 6 + i % 290] for i in range(400)]          # 400 more rows with X (3 ids each)
    p += [[1, 7 + i % 200] for i in range(25)]                        # a few more rows with Y
    p += random_sets(rng, 3000, n_bits, mean=3, p_empty=0.05, dup=True)
    q = [[0, 1], [0], [1], [0, 1, 5], [1, 0, 0, 1], [0, 399], [2, 3, 4, 5, 6, 7, 8], list(range(2, 34)), []]
    q += random_sets(rng, 200, n_bits, mean=3, p_empty=0.05, dup=True)
    for k in (1, 3, 10, 16):
        check(q, p, n_bits, k)
    # > 64 rows hold both ids of the query: more noted rows than the kernel resolves
    p2 = [[10, 11] for _ in range(150)] + random_sets(rng, 2000, n_bits, mean=2.2)
    check([[10, 11], [10], [11, 12], [10, 11, 12, 13]], p2, n_bits, 10)
    # forced-zero rows inside the heads, long lists (vocabulary of 50 ids: ~200 postings per id)
    p3 = random_sets(rng, 5000, 50, mean=2, p_empty=0.02, dup=True)
    check([list(x) for x in p3[:400]], p3, 50, 10, zero_diag=True)
    check([list(x) for x in p3[1000:1300]], p3, 50, 16, zero_diag=True, query_base=1000)
    # sharded pool: global indices and a diagonal that lies outside / inside the shard
    check([list(x) for x in p3[:300]], p3[200:], 50, 10, zero_diag=True, query_base=0, pool_base=200)


def test_head_kernel_key_overflow_hands_over():
    """The head kernel ranks list heads as 32-bit keys |pool set| << bits(rows) | row: in a pool of 2.1 M rows a set of
    more than 2 046 ids does not fit the key, and the query must travel on to the register kernel — same result."""
    n_bits, npool, k = 3000, (1 << 21) + 5, 10
    rng = np.random.default_rng(3)
    ids = rng.integers(10, n_bits, npool).astype(np.int32)           # one id per row (ids 0..9 stay rare)
    lens = np.ones(npool, np.int64)
    big = np.arange(0, 2500, dtype=np.int32)                         # row 7: a set of 2 500 ids, among them the rare ones
    lens[7] = big.size
    off = np.zeros(npool + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    flat = np.empty(int(off[-1]), np.int32)
    flat[off[:-1]] = ids
    flat[off[7]:off[8]] = big
    few = [11, 12, npool - 3, 100000, 2000000]                       # other rows that hold the rare ids 3 and 4
    for r in few:
        flat[off[r]] = 3 if r % 2 else 4
    bp = set_encoder.encode_csr(torch.as_tensor(flat), torch.as_tensor(off), n_bits)
    index = engine.build_postings(bp)
    q = [[3], [4, 3], [3, 2999], [5], [2998, 2997, 4]]
    qi, qo = to_csr(q)
    got = engine.jaccard_topk_postings(torch.as_tensor(qi).cuda(), torch.as_tensor(qo).cuda(), index, k)
    oi, ou, ox = jo.c_topk(qi, qo, flat, off, k)
    assert np.array_equal(got[2].cpu().numpy(), ox)
    assert np.array_equal(got[0].cpu().numpy(), oi) and np.array_equal(got[1].cpu().numpy(), ou)
