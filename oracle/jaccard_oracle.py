"""CPU ORACLE for the Jaccard annotation path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
Nothing under rag4dyg_b200/ imports it; the product path has no CPU fallback.

It restates, on the CPU, the algorithm of /root/reference/retrieval_data_annotation.py (file:line cited per
function).  Parity pinning: tests/test_oracle_golden.py runs `annotate` on the shipped datasets and compares the
output files with golden vectors produced by running the unmodified reference in the build container
(oracle/make_golden.py; sha256 in tests/golden/manifest.json, identical to SURVEY.md section 8c).

The one deliberate deviation from the reference: every ranking uses the canonical tie rule (score descending, index
ascending) == np.argsort(-x, kind='stable'), because the reference's default argsort is unstable and its tie order
depends on the CPU's SIMD level.  The goldens were generated with the same switch.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


# ----------------------------------------------------------------------------- reference :5-15
def co_occurrence_ratio(seq_i, seq_j):
    if type(seq_j) is not list:
        seq_j = [seq_j]
    if seq_i is None or seq_j is None:
        return 0
    if len(seq_i) == 0 or len(seq_j) == 0:
        return 0
    a, b = set(seq_i), set(seq_j)
    return len(a & b) / len(a | b)


# ----------------------------------------------------------------------------- reference :17-34
def get_input_seq(line):
    body = line.split("<|history|>")[1].split("<|endofhistory|>")[0]
    return [t for t in body.split(" ") if t != ""]


def get_output_seq(line):
    body = line.split("<|pre|>")[1].split("<|endofpre|>")[0]
    return [t for t in body.split(" ") if t != "" and "time" not in t]


def get_inout_list(data, gt):
    return [get_input_seq(d) for d in data], [get_output_seq(g) for g in gt[: len(data)]]


# ----------------------------------------------------------------------------- reference :36-41
def occurrence_matrix(target, source):
    """Pure-Python double loop over Python sets — the reference's hot loop, used for small cases and as the
    timed CPU baseline ("port")."""
    m = np.zeros((len(target), len(source)))
    for i, a in enumerate(target):
        for j, b in enumerate(source):
            m[i, j] = co_occurrence_ratio(a, b)
    return m


def counts_matrix(target, source):
    """Exact integer formulation for mid-size cases: (inter, union) int64 [Q, N] via a 0/1 incidence product.
    Validated against `occurrence_matrix` in tests/test_oracle.py."""
    toks = {}
    for seqs in (target, source):
        for s in seqs:
            for t in s:
                toks.setdefault(t, len(toks))
    v = max(1, len(toks))

    def incidence(seqs):
        a = np.zeros((len(seqs), v), dtype=np.float32)
        for i, s in enumerate(seqs):
            for t in s:
                a[i, toks[t]] = 1.0
        return a

    qa, pa = incidence(target), incidence(source)
    assert v < (1 << 24)  # 0/1 sums stay exact in float32
    inter = np.rint(qa @ pa.T).astype(np.int64)
    union = qa.sum(1, dtype=np.float64).astype(np.int64)[:, None] + pa.sum(1, dtype=np.float64).astype(np.int64)[None, :] - inter
    return inter, union


def scores_from_counts(inter, union):
    """len(intersection)/len(union) in float64 (reference :14); 0 where either set is empty (:11)."""
    out = np.zeros(inter.shape, dtype=np.float64)
    np.divide(inter.astype(np.float64), union.astype(np.float64), out=out, where=inter > 0)
    return out


# ----------------------------------------------------------------------------- canonical ranking
def rank_stable(scores):
    """np.argsort(-scores) of reference :56,:89,:101 with the canonical tie rule."""
    return np.argsort(-np.asarray(scores), axis=-1, kind="stable")


def topk_stable(scores, k):
    order = rank_stable(scores)[..., :k]
    return order, np.take_along_axis(np.asarray(scores), order, axis=-1)


# ----------------------------------------------------------------------------- reference :43-85
def train_annotation_lines(out_m, in_m, threshold, neg_num, dataset):
    """Returns (index_lines, score_lines, n_positive).  Consumes the legacy global numpy RNG exactly like :79."""
    idx_lines, score_lines = [], []
    for i in range(out_m.shape[0]):
        pos = np.where(out_m[i] > threshold)[0].tolist()
        if not pos:
            continue
        pos_set = set(pos)
        order = rank_stable(in_m[i])
        negs = []
        for j in order:
            if j not in pos_set and out_m[i, j] > 0:
                negs.append(j)
                if len(negs) == neg_num:
                    break
        if len(negs) < neg_num:
            for j in order:
                if j not in pos_set and out_m[i, j] == 0:
                    negs.append(j)
                    if len(negs) == neg_num:
                        break
        if "dialog" in dataset:
            pos = pos[:4]
        for p in pos:
            n = np.random.choice(negs)
            idx_lines.append(f"{i} {p} {n}")
            score_lines.append(f"{i} {out_m[i, p]} {out_m[i, n]}")
    return idx_lines, score_lines, len(idx_lines)


# ----------------------------------------------------------------------------- reference :88-103
def index_score_lines(m):
    order = rank_stable(m)
    return ([" ".join(str(x) for x in order[i]) for i in range(m.shape[0])],
            [" ".join(str(x) for x in m[i]) for i in range(m.shape[0])])


def topk_lines(m, topk=10):
    order, vals = topk_stable(m, topk)
    return ([" ".join(map(str, order[i])) for i in range(m.shape[0])],
            [" ".join(map(str, vals[i])) for i in range(m.shape[0])])


def _read(path):
    with open(path) as f:
        return [ln for ln in f.read().splitlines() if len(ln) > 0 and not ln.isspace()]


def _write(path, lines):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        for ln in lines:
            f.write(ln + "\n")


def annotate(dataset, timestamp, threshold, root=".", neg_num=5, topk=10, seed=None, fast=True):
    """Restatement of reference __main__ (:109-200); writes the eight files under `root`.  fast=True uses the
    integer incidence product instead of the Python set loop (identical values, see tests)."""
    base = os.path.join(root, "resources", dataset, timestamp)
    train = _read(os.path.join(base, "train.link_prediction"))
    test, test_gt = _read(os.path.join(base, "test.link_prediction")), _read(os.path.join(base, "test_gt.link_prediction"))
    val, val_gt = _read(os.path.join(base, "val.link_prediction")), _read(os.path.join(base, "val_gt.link_prediction"))
    tr_in, tr_out = get_inout_list(train, train)
    _, te_out = get_inout_list(test, test_gt)
    _, va_out = get_inout_list(val, val_gt)

    def matrix(a, b):
        return scores_from_counts(*counts_matrix(a, b)) if fast else occurrence_matrix(a, b)

    m_out, m_in = matrix(tr_out, tr_out), matrix(tr_in, tr_in)
    m_te, m_va = matrix(te_out, tr_out), matrix(va_out, tr_out)
    np.fill_diagonal(m_out, 0)
    np.fill_diagonal(m_in, 0)
    if seed is not None:
        np.random.seed(seed)
    r = os.path.join(root, "resources", dataset, timestamp, "train_retrieval")
    g = os.path.join(root, "resources", "train_generator", dataset, timestamp, "train_gt_topk")
    il, sl, n_pos = train_annotation_lines(m_out, m_in, threshold, neg_num, dataset)
    _write(os.path.join(r, "train_index.retrieval"), il)
    _write(os.path.join(r, "train_score.retrieval"), sl)
    for name, m in (("test", m_te), ("val", m_va)):
        il, sl = index_score_lines(m)
        _write(os.path.join(r, f"{name}_index.retrieval"), il)
        _write(os.path.join(r, f"{name}_score.retrieval"), sl)
    il, sl = topk_lines(m_out, topk)
    _write(os.path.join(g, "train_index.gen"), il)
    _write(os.path.join(g, "train_score.gen"), sl)
    return n_pos


# ----------------------------------------------------------------------------- C restatement (oracle/jaccard_oracle.c)
_clib = None


def build_c(force=False):
    """gcc -O2 the plain-C restatement into oracle/libjoracle.so (building the checker is not using it)."""
    src, so = os.path.join(HERE, "jaccard_oracle.c"), os.path.join(HERE, "libjoracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, src])
    return so


def _c():
    global _clib
    if _clib is None:
        lib = ctypes.CDLL(build_c())
        lib.joracle_topk.restype = ctypes.c_int
        lib.joracle_topk.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int64, ctypes.c_int64,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.joracle_counts.restype = ctypes.c_int
        lib.joracle_counts.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
        _clib = lib
    return _clib


def _csr_args(bit_pos, row_off):
    bp = np.ascontiguousarray(bit_pos, dtype=np.int32)
    ro = np.ascontiguousarray(row_off, dtype=np.int64)
    return bp, ro


def c_topk(q_pos, q_off, p_pos, p_off, k, zero_diag=False, query_base=0, pool_base=0):
    """Top-k of every query against the pool from CSR id lists, canonical order.  Returns (inter, union, idx)
    int64/int64/int32 arrays [nq, k]; rows short of k padded with (0, 1, 0x7fffffff)."""
    qp, qo = _csr_args(q_pos, q_off)
    pp, po = _csr_args(p_pos, p_off)
    nq, npool = len(qo) - 1, len(po) - 1
    inter = np.zeros((nq, k), dtype=np.int64)
    union = np.ones((nq, k), dtype=np.int64)
    idx = np.full((nq, k), 0x7FFFFFFF, dtype=np.int32)
    rc = _c().joracle_topk(qp.ctypes.data, qo.ctypes.data, nq, pp.ctypes.data, po.ctypes.data, npool, k,
                           int(bool(zero_diag)), query_base, pool_base, inter.ctypes.data, union.ctypes.data,
                           idx.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"joracle_topk rc={rc}")
    return inter, union, idx


def c_counts(q_pos, q_off, p_pos, p_off):
    """(inter, union) int32 [nq, np] from CSR id lists (duplicates allowed)."""
    qp, qo = _csr_args(q_pos, q_off)
    pp, po = _csr_args(p_pos, p_off)
    nq, npool = len(qo) - 1, len(po) - 1
    inter = np.zeros((nq, npool), dtype=np.int32)
    union = np.zeros((nq, npool), dtype=np.int32)
    rc = _c().joracle_counts(qp.ctypes.data, qo.ctypes.data, nq, pp.ctypes.data, po.ctypes.data, npool,
                             inter.ctypes.data, union.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"joracle_counts rc={rc}")
    return inter, union
