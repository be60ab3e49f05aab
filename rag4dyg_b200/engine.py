"""Thin torch-tensor front end of the C ABI (include/r4d.h).

Every function here takes CUDA tensors, passes raw device pointers + the current CUDA stream to libr4d.so and returns
CUDA tensors.  Nothing is computed in Python/torch: a missing library or device raises (no CPU fallback).
"""
import functools
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import (DENSE_COS_DECAY, DENSE_HALF_COS, DENSE_HALF_COS_DECAY, PREC_BF16, PREC_BF16X3, R4D_IDX_NONE,
                   R4D_TOPK_MAX, R4DError, check)

__all__ = [
    "BitsetMatrix", "encode_bitsets", "jaccard_full", "jaccard_topk", "jaccard_topk_merge", "jaccard_topk_scatter", "dense_topk_scatter", "rank_rows", "topk_rows",
    "PostingsIndex", "build_postings", "jaccard_topk_postings", "jaccard_topk_postings_scatter",
    "triplet_mine", "triplet_mine_bits", "triplet_sample", "DensePlanes", "dense_prepare", "dense_topk", "dense_full", "dense_topk_merge", "meanpool_prepare", "R4D_IDX_NONE",
    "R4D_TOPK_MAX", "DENSE_HALF_COS", "DENSE_COS_DECAY", "DENSE_HALF_COS_DECAY", "PREC_BF16", "PREC_BF16X3",
    "launch_count", "reset_launch_count",
]

_launch_base = 0  # library launch counter at the last reset (bench.py reports the difference as gpu_launches)


def launch_count():
    """Kernels of OURS the library has enqueued since reset_launch_count() (counted inside libr4d.so at every <<<>>>)."""
    return int(_lib.load().r4d_kernel_launches()) - _launch_base


def reset_launch_count():
    global _launch_base
    _launch_base = int(_lib.load().r4d_kernel_launches())


def _count(n):
    """Kept for call-site symmetry; the library counts its own launches."""


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _device_of(x):
    if isinstance(x, torch.Tensor):
        return x.device
    for attr in ("bits", "hi", "blob"):          # BitsetMatrix / DensePlanes / PostingsIndex
        t = getattr(x, attr, None)
        if isinstance(t, torch.Tensor):
            return t.device
    return None


def _on_input_device(fn):
    """Run `fn` with the CUDA device of its tensor arguments current (kernels, streams and the per-device shared-memory
    opt-ins all follow the current device); inputs on different GPUs raise instead of launching on the wrong one."""
    @functools.wraps(fn)
    def wrapped(*args, **kw):
        dev = None
        for a in list(args) + list(kw.values()):
            d = _device_of(a)
            if d is not None and d.type == "cuda":
                if dev is None:
                    dev = d
                elif d != dev:
                    raise R4DError(f"{fn.__name__}: inputs live on different devices ({dev} vs {d})")
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kw)
        with torch.cuda.device(dev):
            return fn(*args, **kw)
    return wrapped


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _dev_tensor(t, dtype, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise R4DError(f"{name}: expected a CUDA tensor (this path has no CPU implementation)")
    if t.dtype != dtype:
        raise R4DError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise R4DError(f"{name}: tensor must be contiguous")
    return t


@dataclass
class BitsetMatrix:
    """Rows of fixed-width uint32 bitsets in HBM: bits[n_rows, pitch_words] (int32 storage), card[n_rows] = |set|."""
    bits: torch.Tensor
    card: torch.Tensor
    n_bits: int
    words: int
    pitch_words: int

    @property
    def n_rows(self):
        return self.bits.shape[0]

    def rows(self, start, stop):
        """Contiguous row slice (a pool shard); shares storage."""
        return BitsetMatrix(self.bits[start:stop], self.card[start:stop], self.n_bits, self.words, self.pitch_words)


@_on_input_device
def encode_bitsets(bit_pos, row_off, n_bits):
    """CSR (int32 bit positions, int64 row offsets; CUDA tensors) -> BitsetMatrix.  r4d_bitset_encode."""
    lib = _lib.load()
    bit_pos = _dev_tensor(bit_pos, torch.int32, "bit_pos")
    row_off = _dev_tensor(row_off, torch.int64, "row_off")
    n_rows = row_off.numel() - 1
    if n_rows < 0:
        raise R4DError("row_off must hold n_rows+1 offsets")
    words = lib.r4d_bitset_words(n_bits)
    pitch = lib.r4d_bitset_pitch_words(n_bits)
    dev = row_off.device
    bits = torch.empty((n_rows, pitch), dtype=torch.int32, device=dev)
    card = torch.empty((n_rows,), dtype=torch.int32, device=dev)
    check(lib.r4d_bitset_encode(_ptr(bit_pos), _ptr(row_off), n_rows, n_bits, pitch, _ptr(bits), _ptr(card), _stream()),
          "r4d_bitset_encode")
    _count(1 if n_rows else 0)
    return BitsetMatrix(bits, card, n_bits, words, pitch)


def _check_pair(q, p):
    if q.words != p.words or q.pitch_words != p.pitch_words:
        raise R4DError("query and pool bitsets must share the same universe (words/pitch differ)")


@_on_input_device
def jaccard_full(q, p, zero_diag=False, want_score=True, query_base=0, pool_base=0):
    """All-pairs intersection counts (int32 [nq, np]) and float64 Jaccard scores.  r4d_jaccard_full."""
    lib = _lib.load()
    _check_pair(q, p)
    nq, np_ = q.n_rows, p.n_rows
    dev = q.bits.device
    inter = torch.empty((nq, np_), dtype=torch.int32, device=dev)
    score = torch.empty((nq, np_), dtype=torch.float64, device=dev) if want_score else None
    check(lib.r4d_jaccard_full(_ptr(q.bits), _ptr(q.card), nq, _ptr(p.bits), _ptr(p.card), np_, q.words, q.pitch_words,
                               int(bool(zero_diag)), query_base, pool_base, _ptr(inter), np_, _ptr(score), np_,
                               _stream()), "r4d_jaccard_full")
    _count(1 if nq and np_ else 0)
    return inter, score


@_on_input_device
def jaccard_topk(q, p, k, zero_diag=False, query_base=0, pool_base=0, workspace=None):
    """Fused scorer + top-K: (inter, union, idx) int32 [nq, k], order (score desc, idx asc).  r4d_jaccard_topk."""
    lib = _lib.load()
    _check_pair(q, p)
    nq, np_ = q.n_rows, p.n_rows
    dev = q.bits.device
    need = lib.r4d_jaccard_topk_workspace_bytes(nq, np_, k)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty((need,), dtype=torch.uint8, device=dev)
    top_inter = torch.empty((nq, k), dtype=torch.int32, device=dev)
    top_union = torch.empty((nq, k), dtype=torch.int32, device=dev)
    top_idx = torch.empty((nq, k), dtype=torch.int32, device=dev)
    check(lib.r4d_jaccard_topk(_ptr(q.bits), _ptr(q.card), nq, _ptr(p.bits), _ptr(p.card), np_, q.words, q.pitch_words,
                               k, int(bool(zero_diag)), query_base, pool_base, _ptr(top_inter), _ptr(top_union),
                               _ptr(top_idx), _ptr(workspace), workspace.numel(), _stream()), "r4d_jaccard_topk")
    _count(2 if nq and np_ else (1 if nq else 0))
    return top_inter, top_union, top_idx


@_on_input_device
def jaccard_topk_scatter(q, p, k, peer_ptrs, world, rank, zero_diag=False, query_base=0, pool_base=0, workspace=None):
    """Fused exchange: local fused top-K whose final lists are stored straight into slot `rank` of every peer's
    gather buffer [3][world][nq][k] over NVLink.  peer_ptrs: ctypes array of `world` device pointers.
    r4d_jaccard_topk_scatter."""
    lib = _lib.load()
    _check_pair(q, p)
    nq, np_ = q.n_rows, p.n_rows
    need = lib.r4d_jaccard_topk_workspace_bytes(nq, np_, k)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty((need,), dtype=torch.uint8, device=q.bits.device)
    check(lib.r4d_jaccard_topk_scatter(_ptr(q.bits), _ptr(q.card), nq, _ptr(p.bits), _ptr(p.card), np_, q.words,
                                       q.pitch_words, k, int(bool(zero_diag)), query_base, pool_base, peer_ptrs, world,
                                       rank, _ptr(workspace), workspace.numel(), _stream()), "r4d_jaccard_topk_scatter")
    _count(2 if nq and np_ else (1 if nq else 0))


@dataclass
class PostingsIndex:
    """Inverted index of a pool (shard): node id -> pool rows holding it, in HBM (r4d_postings_build).  State of the
    pool like its bitsets; built once."""
    blob: torch.Tensor      # uint8, the index blob
    card: torch.Tensor      # int32 [n_rows], |set| per pool row (shared with the BitsetMatrix)
    n_rows: int
    n_bits: int
    nnz: int

    @property
    def device(self):
        return self.blob.device


@_on_input_device
def build_postings(p):
    """BitsetMatrix of the pool (shard) -> PostingsIndex.  r4d_postings_build.  One-time pool set-up: reads the exact
    posting count (sum of the cardinalities) and the build status back, i.e. synchronises."""
    lib = _lib.load()
    dev = p.bits.device
    with torch.cuda.device(dev):
        nnz = int(p.card.sum(dtype=torch.int64).item()) if p.n_rows else 0
        need = lib.r4d_postings_index_bytes(p.n_rows, p.n_bits, nnz)
        if need == 0:
            raise R4DError(f"postings index: unsupported shape (rows={p.n_rows}, n_bits={p.n_bits}, postings={nnz}); "
                           "use the bitset path (jaccard_topk)")
        blob = torch.empty((need,), dtype=torch.uint8, device=dev)
        ws = torch.empty((lib.r4d_postings_build_workspace_bytes(p.n_rows, p.n_bits),), dtype=torch.uint8, device=dev)
        check(lib.r4d_postings_build(_ptr(p.bits), _ptr(p.card), p.n_rows, p.n_bits, p.pitch_words, nnz, _ptr(blob),
                                     blob.numel(), _ptr(ws), ws.numel(), _stream()), "r4d_postings_build")
        status = int(blob[4:8].view(torch.int32).item())
        if status != 0:
            raise R4DError("postings index: build overflowed its capacity (cardinalities do not match the bitsets)")
    return PostingsIndex(blob, p.card, p.n_rows, p.n_bits, nnz)


def _postings_args(q_ids, q_off, index):
    q_ids = _dev_tensor(q_ids, torch.int32, "q_ids")
    q_off = _dev_tensor(q_off, torch.int64, "q_off")
    if q_ids.device != index.device or q_off.device != index.device:
        raise R4DError("jaccard_topk_postings: queries and index live on different devices")
    nq = q_off.numel() - 1
    if nq < 0:
        raise R4DError("q_off must hold nq+1 offsets")
    return q_ids, q_off, nq


@_on_input_device
def jaccard_topk_postings(q_ids, q_off, index, k, zero_diag=False, query_base=0, pool_base=0, workspace=None, out=None,
                          q_nnz=None):
    """Fused Jaccard scorer + top-K over pool postings: queries as CSR id lists (int32 ids, int64 offsets, CUDA),
    (inter, union, idx) int32 [nq, k] in the canonical order.  r4d_jaccard_topk_postings.
    q_nnz: number of ids the nq rows hold, when q_off is a row range of a larger CSR (a sizing hint only)."""
    lib = _lib.load()
    q_ids, q_off, nq = _postings_args(q_ids, q_off, index)
    dev = index.device
    with torch.cuda.device(dev):
        need = lib.r4d_jaccard_topk_postings_workspace_bytes(nq)
        if workspace is None or workspace.numel() < need:
            workspace = torch.empty((need,), dtype=torch.uint8, device=dev)
        if out is None:
            out = tuple(torch.empty((nq, k), dtype=torch.int32, device=dev) for _ in range(3))
        else:
            # device tensors, or PINNED host tensors (device-addressable: the kernel then stores over PCIe directly)
            for o in out:
                if not (o.dtype == torch.int32 and o.is_contiguous() and tuple(o.shape) == (nq, k) and
                        (o.is_cuda or o.is_pinned())):
                    raise R4DError("jaccard_topk_postings: `out` must be contiguous int32 [nq, k] CUDA or pinned host tensors")
        check(lib.r4d_jaccard_topk_postings(_ptr(q_ids), _ptr(q_off), nq, q_ids.numel() if q_nnz is None else int(q_nnz),
                                            _ptr(index.blob), _ptr(index.card), index.n_rows,
                                            index.n_bits, index.nnz, k, int(bool(zero_diag)), query_base, pool_base,
                                            _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(workspace), workspace.numel(),
                                            _stream()), "r4d_jaccard_topk_postings")
    return out


@_on_input_device
def jaccard_topk_postings_packed(q_ids, q_off, index, k, zero_diag=False, query_base=0, pool_base=0, workspace=None, out=None,
                                 q_nnz=None):
    """The same top-K with PACKED results (r4d_jaccard_topk_postings_packed): out = (pair int32 [nq, k] holding
    inter << 16 | |pool set|, idx int32 [nq, k], q_card int32 [nq]); CUDA or pinned host tensors.  8 bytes per entry instead
    of 12 — for results that cross PCIe.  `unpack_topk` recovers (inter, union, idx)."""
    lib = _lib.load()
    q_ids, q_off, nq = _postings_args(q_ids, q_off, index)
    dev = index.device
    with torch.cuda.device(dev):
        need = lib.r4d_jaccard_topk_postings_workspace_bytes(nq)
        if workspace is None or workspace.numel() < need:
            workspace = torch.empty((need,), dtype=torch.uint8, device=dev)
        if out is None:
            out = (torch.empty((nq, k), dtype=torch.int32, device=dev), torch.empty((nq, k), dtype=torch.int32, device=dev),
                   torch.empty((nq,), dtype=torch.int32, device=dev))
        else:
            for o, shape in zip(out, ((nq, k), (nq, k), (nq,))):
                if not (o.dtype == torch.int32 and o.is_contiguous() and tuple(o.shape) == shape and
                        (o.is_cuda or o.is_pinned())):
                    raise R4DError("jaccard_topk_postings_packed: `out` must be contiguous int32 ([nq, k], [nq, k], [nq]) "
                                   "CUDA or pinned host tensors")
        check(lib.r4d_jaccard_topk_postings_packed(_ptr(q_ids), _ptr(q_off), nq, q_ids.numel() if q_nnz is None else int(q_nnz),
                                                   _ptr(index.blob), _ptr(index.card), index.n_rows,
                                                   index.n_bits, index.nnz, k, int(bool(zero_diag)), query_base, pool_base,
                                                   _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(workspace),
                                                   workspace.numel(), _stream()), "r4d_jaccard_topk_postings_packed")
    return out


def unpack_topk(pair, idx, q_card):
    """(pair, idx, q_card) of jaccard_topk_postings_packed -> (inter, union, idx) int32 [nq, k], the planes
    jaccard_topk_postings writes: inter = pair >> 16, union = q_card + (pair & 0xffff) - inter; padding entries
    (idx == 0x7fffffff) are (0, 1)."""
    inter = (pair >> 16) & 0xffff
    union = q_card.unsqueeze(1) + (pair & 0xffff) - inter
    union = torch.where(idx == 0x7FFFFFFF, torch.ones_like(union), union)
    return inter, union, idx


@_on_input_device
def jaccard_topk_postings_scatter(q_ids, q_off, index, k, peer_ptrs, world, rank, zero_diag=False, query_base=0,
                                  pool_base=0, workspace=None):
    """Fused exchange variant: final lists go to slot `rank` of every peer's gather buffer [3][world][nq][k]."""
    lib = _lib.load()
    q_ids, q_off, nq = _postings_args(q_ids, q_off, index)
    dev = index.device
    with torch.cuda.device(dev):
        need = lib.r4d_jaccard_topk_postings_workspace_bytes(nq)
        if workspace is None or workspace.numel() < need:
            workspace = torch.empty((need,), dtype=torch.uint8, device=dev)
        check(lib.r4d_jaccard_topk_postings_scatter(_ptr(q_ids), _ptr(q_off), nq, q_ids.numel(), _ptr(index.blob), _ptr(index.card),
                                                    index.n_rows, index.n_bits, index.nnz, k, int(bool(zero_diag)),
                                                    query_base, pool_base, peer_ptrs, world, rank, _ptr(workspace),
                                                    workspace.numel(), _stream()), "r4d_jaccard_topk_postings_scatter")


@_on_input_device
def jaccard_topk_merge(inter, uni, idx, k_out):
    """Merge candidate lists [n_lists, nq, k_in] -> [nq, k_out].  r4d_jaccard_topk_merge."""
    lib = _lib.load()
    inter = _dev_tensor(inter, torch.int32, "inter")
    uni = _dev_tensor(uni, torch.int32, "uni")
    idx = _dev_tensor(idx, torch.int32, "idx")
    n_lists, nq, k_in = inter.shape
    dev = inter.device
    o_i = torch.empty((nq, k_out), dtype=torch.int32, device=dev)
    o_u = torch.empty((nq, k_out), dtype=torch.int32, device=dev)
    o_x = torch.empty((nq, k_out), dtype=torch.int32, device=dev)
    check(lib.r4d_jaccard_topk_merge(_ptr(inter), _ptr(uni), _ptr(idx), n_lists, nq, k_in, k_out, _ptr(o_i), _ptr(o_u),
                                     _ptr(o_x), _stream()), "r4d_jaccard_topk_merge")
    _count(1 if nq else 0)
    return o_i, o_u, o_x


@_on_input_device
def rank_rows(scores):
    """order[q] = argsort(-scores[q], stable) as int32 [nq, n].  float64 or float32 CUDA matrix."""
    lib = _lib.load()
    if scores.dtype == torch.float64:
        fn, eb = lib.r4d_rank_rows_f64, 8
    elif scores.dtype == torch.float32:
        fn, eb = lib.r4d_rank_rows_f32, 4
    else:
        raise R4DError(f"rank_rows: unsupported dtype {scores.dtype}")
    scores = _dev_tensor(scores, scores.dtype, "scores")
    nq, n = scores.shape
    dev = scores.device
    order = torch.empty((nq, n), dtype=torch.int32, device=dev)
    need = lib.r4d_rank_rows_workspace_bytes(nq, n, eb)
    ws = torch.empty((need,), dtype=torch.uint8, device=dev)
    check(fn(_ptr(scores), nq, n, n, _ptr(order), _ptr(ws), ws.numel(), _stream()), "r4d_rank_rows")
    _count(1 if nq and n else 0)
    return order


@_on_input_device
def topk_rows(scores, k):
    """Top-k of each row of an explicit float64 matrix: (scores f64 [nq,k], idx int32 [nq,k])."""
    lib = _lib.load()
    scores = _dev_tensor(scores, torch.float64, "scores")
    nq, n = scores.shape
    dev = scores.device
    ts = torch.empty((nq, k), dtype=torch.float64, device=dev)
    ti = torch.empty((nq, k), dtype=torch.int32, device=dev)
    check(lib.r4d_topk_rows_f64(_ptr(scores), nq, n, n, k, _ptr(ts), _ptr(ti), _stream()), "r4d_topk_rows_f64")
    _count(1 if nq else 0)
    return ts, ti


@_on_input_device
def triplet_mine(out, inn, thr, neg_num):
    """Device part of save_train_annotation: (n_pos [n], neg [n, neg_num], n_neg [n]) int32."""
    lib = _lib.load()
    out = _dev_tensor(out, torch.float64, "out")
    inn = _dev_tensor(inn, torch.float64, "in")
    n = out.shape[0]
    if out.shape != (n, n) or inn.shape != (n, n):
        raise R4DError("triplet_mine: out/in must be square matrices of the same size")
    dev = out.device
    n_pos = torch.empty((n,), dtype=torch.int32, device=dev)
    neg = torch.empty((n, neg_num), dtype=torch.int32, device=dev)
    n_neg = torch.empty((n,), dtype=torch.int32, device=dev)
    check(lib.r4d_triplet_mine_f64(_ptr(out), _ptr(inn), n, n, float(thr), neg_num, _ptr(n_pos), _ptr(neg), _ptr(n_neg),
                                   _stream()), "r4d_triplet_mine_f64")
    _count(1 if n else 0)
    return n_pos, neg, n_neg


@_on_input_device
def triplet_mine_bits(b_out, b_in, thr, neg_num, zero_diag=True, pos_cap=None):
    """save_train_annotation's device half from the OUT / IN bitsets of the train pool (r4d_triplet_mine; no [n, n] matrix).
    Returns a dict: n_pos [n], pos_row / pos_col [P] (row-major order, like np.where per row), pos_inter / pos_union [P]
    (OUT counts), neg / neg_inter / neg_union [n, neg_num], n_neg [n].  Synchronises once (number of positives)."""
    lib = _lib.load()
    n = b_out.n_rows
    if b_in.n_rows != n:
        raise R4DError("triplet_mine_bits: the OUT and IN bitsets must describe the same rows")
    dev = b_out.bits.device
    i32 = dict(dtype=torch.int32, device=dev)
    n_pos, n_neg = torch.empty((n,), **i32), torch.empty((n,), **i32)
    neg, neg_i, neg_u = (torch.empty((n, neg_num), **i32) for _ in range(3))
    total = torch.zeros((1,), dtype=torch.int64, device=dev)
    cap = int(pos_cap) if pos_cap is not None else max(1 << 16, 16 * n)
    while True:
        key = torch.empty((max(cap, 1),), dtype=torch.int64, device=dev)
        cnt = torch.empty((max(cap, 1),), dtype=torch.int64, device=dev)
        check(lib.r4d_triplet_mine(_ptr(b_out.bits), _ptr(b_out.card), b_out.words, b_out.pitch_words, _ptr(b_in.bits),
                                   _ptr(b_in.card), b_in.words, b_in.pitch_words, n, float(thr), neg_num,
                                   int(bool(zero_diag)), _ptr(n_pos), _ptr(key), _ptr(cnt), cap, _ptr(total), _ptr(neg),
                                   _ptr(neg_i), _ptr(neg_u), _ptr(n_neg), _stream()), "r4d_triplet_mine")
        found = int(total.item())
        if found <= cap:
            break
        cap = found                                   # rare: more positives than the first guess; mine again
    key, order = torch.sort(key[:found])
    cnt = cnt[:found][order]
    return {"n_pos": n_pos, "pos_row": key >> 32, "pos_col": key & 0xFFFFFFFF, "pos_inter": cnt >> 32,
            "pos_union": cnt & 0xFFFFFFFF, "neg": neg, "neg_inter": neg_i, "neg_union": neg_u, "n_neg": n_neg}


@_on_input_device
def triplet_sample(pos_row, row_start, neg, n_neg, seed):
    """Counter-based negative choice per positive pair (r4d_triplet_sample): int32 [n_pairs]."""
    lib = _lib.load()
    pos_row = _dev_tensor(pos_row, torch.int64, "pos_row")
    row_start = _dev_tensor(row_start, torch.int64, "row_start")
    neg = _dev_tensor(neg, torch.int32, "neg")
    n_neg = _dev_tensor(n_neg, torch.int32, "n_neg")
    n_pairs = pos_row.numel()
    choice = torch.empty((n_pairs,), dtype=torch.int32, device=pos_row.device)
    check(lib.r4d_triplet_sample(_ptr(pos_row), _ptr(row_start), n_pairs, _ptr(neg), _ptr(n_neg), neg.shape[1],
                                 int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(choice), _stream()), "r4d_triplet_sample")
    _count(1 if n_pairs else 0)
    return choice


# ---------------------------------------------------------------------------------------------- dense scorer
@dataclass
class DensePlanes:
    """L2-normalised embeddings in the scorer's layout: bf16 hi plane [n, d_pad] (+ lo plane for BF16X3)."""
    hi: torch.Tensor
    lo: torch.Tensor  # None for PREC_BF16
    d: int
    d_pad: int
    prec: int

    @property
    def n_rows(self):
        return self.hi.shape[0]

    def rows(self, start, stop):
        return DensePlanes(self.hi[start:stop], None if self.lo is None else self.lo[start:stop], self.d, self.d_pad,
                           self.prec)


@_on_input_device
def dense_prepare(x, prec=PREC_BF16X3):
    """fp32 [n, d] -> row-normalised bf16 planes.  r4d_dense_prepare."""
    lib = _lib.load()
    x = _dev_tensor(x, torch.float32, "x")
    n, d = x.shape
    d_pad = lib.r4d_dense_dpad(d)
    hi = torch.empty((n, d_pad), dtype=torch.bfloat16, device=x.device)
    lo = torch.empty((n, d_pad), dtype=torch.bfloat16, device=x.device) if prec == PREC_BF16X3 else None
    check(lib.r4d_dense_prepare(_ptr(x), n, d, d, prec, _ptr(hi), _ptr(lo), _stream()), "r4d_dense_prepare")
    _count(1 if n else 0)
    return DensePlanes(hi, lo, d, d_pad, prec)


@_on_input_device
def meanpool_prepare(hidden, prec=PREC_BF16X3, want_mean=False):
    """hidden fp32 [B, L, D] -> (DensePlanes of the L2-normalised mean over L, fp32 means or None).
    r4d_meanpool_prepare: replaces torch.mean(h, dim=1) + per-batch normalisation (train_retriever.py:420,433)."""
    lib = _lib.load()
    hidden = _dev_tensor(hidden, torch.float32, "hidden")
    b, l, d = hidden.shape
    d_pad = lib.r4d_dense_dpad(d)
    dev = hidden.device
    hi = torch.empty((b, d_pad), dtype=torch.bfloat16, device=dev)
    lo = torch.empty((b, d_pad), dtype=torch.bfloat16, device=dev) if prec == PREC_BF16X3 else None
    mean = torch.empty((b, d), dtype=torch.float32, device=dev) if want_mean else None
    ws = torch.empty((lib.r4d_meanpool_workspace_bytes(b, d),), dtype=torch.uint8, device=dev)
    check(lib.r4d_meanpool_prepare(_ptr(hidden), b, l, d, prec, _ptr(mean), _ptr(hi), _ptr(lo), _ptr(ws), ws.numel(),
                                   _stream()), "r4d_meanpool_prepare")
    _count(2 if b else 0)
    return DensePlanes(hi, lo, d, d_pad, prec), mean


def _check_dense(q, p, q_time, p_time, mode):
    if q.d_pad != p.d_pad or q.prec != p.prec:
        raise R4DError("dense: query/pool planes differ in width or precision")
    if mode != DENSE_HALF_COS:
        if q_time is None or p_time is None:
            raise R4DError("dense: decay modes need q_time and p_time")
        _dev_tensor(q_time, torch.float32, "q_time")
        _dev_tensor(p_time, torch.float32, "p_time")
        if q_time.numel() != q.n_rows or p_time.numel() != p.n_rows:
            raise R4DError("dense: time vectors must have one entry per row")


@_on_input_device
def dense_topk(q, p, k, mode=DENSE_HALF_COS, q_time=None, p_time=None, lam=0.0, pool_base=0, workspace=None):
    """Fused tcgen05 contraction + epilogue + top-K: (score f32 [nq,k], idx int32 [nq,k]).  r4d_dense_topk."""
    lib = _lib.load()
    _check_dense(q, p, q_time, p_time, mode)
    nq, np_ = q.n_rows, p.n_rows
    dev = q.hi.device
    need = lib.r4d_dense_topk_workspace_bytes(nq, np_, k)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty((need,), dtype=torch.uint8, device=dev)
    ts = torch.empty((nq, k), dtype=torch.float32, device=dev)
    ti = torch.empty((nq, k), dtype=torch.int32, device=dev)
    check(lib.r4d_dense_topk(_ptr(q.hi), _ptr(q.lo), nq, _ptr(p.hi), _ptr(p.lo), np_, q.d_pad, q.prec, _ptr(q_time),
                             _ptr(p_time), float(lam), mode, k, pool_base, _ptr(ts), _ptr(ti), _ptr(workspace),
                             workspace.numel(), _stream()), "r4d_dense_topk")
    _count(2 if nq and np_ else (1 if nq else 0))
    return ts, ti


@_on_input_device
def dense_topk_scatter(q, p, k, peer_ptrs, world, rank, mode=DENSE_HALF_COS, q_time=None, p_time=None, lam=0.0,
                       pool_base=0, workspace=None):
    """Fused exchange for the dense scorer: peers' buffers are [2][world][nq][k] (float32 scores, int32 indices).
    r4d_dense_topk_scatter."""
    lib = _lib.load()
    _check_dense(q, p, q_time, p_time, mode)
    nq, np_ = q.n_rows, p.n_rows
    need = lib.r4d_dense_topk_workspace_bytes(nq, np_, k)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty((need,), dtype=torch.uint8, device=q.hi.device)
    check(lib.r4d_dense_topk_scatter(_ptr(q.hi), _ptr(q.lo), nq, _ptr(p.hi), _ptr(p.lo), np_, q.d_pad, q.prec,
                                     _ptr(q_time), _ptr(p_time), float(lam), mode, k, pool_base, peer_ptrs, world, rank,
                                     _ptr(workspace), workspace.numel(), _stream()), "r4d_dense_topk_scatter")
    _count(2 if nq and np_ else (1 if nq else 0))


@_on_input_device
def dense_full(q, p, mode=DENSE_HALF_COS, q_time=None, p_time=None, lam=0.0):
    """Full score rows f32 [nq, np].  r4d_dense_full."""
    lib = _lib.load()
    _check_dense(q, p, q_time, p_time, mode)
    nq, np_ = q.n_rows, p.n_rows
    scores = torch.empty((nq, np_), dtype=torch.float32, device=q.hi.device)
    check(lib.r4d_dense_full(_ptr(q.hi), _ptr(q.lo), nq, _ptr(p.hi), _ptr(p.lo), np_, q.d_pad, q.prec, _ptr(q_time),
                             _ptr(p_time), float(lam), mode, _ptr(scores), np_, _stream()), "r4d_dense_full")
    _count(1 if nq and np_ else 0)
    return scores


@_on_input_device
def dense_topk_merge(score, idx, k_out):
    """Merge [n_lists, nq, k_in] dense candidate lists -> [nq, k_out]."""
    lib = _lib.load()
    score = _dev_tensor(score, torch.float32, "score")
    idx = _dev_tensor(idx, torch.int32, "idx")
    n_lists, nq, k_in = score.shape
    o_s = torch.empty((nq, k_out), dtype=torch.float32, device=score.device)
    o_i = torch.empty((nq, k_out), dtype=torch.int32, device=score.device)
    check(lib.r4d_dense_topk_merge(_ptr(score), _ptr(idx), n_lists, nq, k_in, k_out, _ptr(o_s), _ptr(o_i), _stream()),
          "r4d_dense_topk_merge")
    _count(1 if nq else 0)
    return o_s, o_i
