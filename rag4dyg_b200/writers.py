"""Text writers producing byte-identical reference files (retrieval_data_annotation.py:88-103,
train/train_retriever.py:357-368) without a Python-level `str()` per element.

The reference formats every score with `str(np.float64)` (shortest round-trip repr) or f"{x:.4f}".  Jaccard score
matrices hold few distinct values (hard part 6), so each distinct value is formatted once with the
reference's own formatter and rows are assembled by table lookup.
"""
import ctypes

import numpy as np
import pandas as pd

from . import _lib

_CHUNK_BYTES = 64 << 20  # formatting buffer per call into the native formatter


def write_int_rows(path, rows, mode="w"):
    """Each row of the 2-D integer array as space separated decimals (== ' '.join(str(x) for x in row)),
    formatted by the native r4d_format_int_rows."""
    lib = _lib.load()
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    if rows.ndim != 2:
        raise ValueError("write_int_rows expects a 2-D array")
    nq, n = rows.shape
    per_row = max(1, n * 12 + 1)
    step = max(1, _CHUNK_BYTES // per_row)
    with open(path, mode + "b") as f:
        for r0 in range(0, nq, step):
            blk = rows[r0:r0 + step]
            cap = lib.r4d_format_int_rows_bound(blk.shape[0], n)
            buf = ctypes.create_string_buffer(cap)
            got = lib.r4d_format_int_rows(blk.ctypes.data, blk.shape[0], n, n, ctypes.addressof(buf), cap)
            if got < 0:
                raise _lib.R4DError(f"r4d_format_int_rows failed ({got}): {_lib.last_error()}")
            f.write(memoryview(buf)[:got])


def _float_codes(mat, fmt):
    """(codes int [shape of mat], strings object[n_distinct]) with strings[codes] == fmt(mat) elementwise."""
    mat = np.ascontiguousarray(mat)
    if mat.dtype == np.float64:
        raw = mat.view(np.uint64)
    elif mat.dtype == np.float32:
        raw = mat.view(np.uint32)
    else:
        raise TypeError(f"unsupported score dtype {mat.dtype}")
    # hash-based factorisation (O(n)); np.unique would sort 10^7 values to find ~10^3 distinct ones
    inv, uniq = pd.factorize(raw.ravel())
    vals = np.asarray(uniq).view(mat.dtype)
    strings = np.array([fmt(v) for v in vals], dtype=object)
    return inv.reshape(mat.shape), strings


def fmt_str(v):
    """The reference's `str(x)` on a numpy float64 scalar (retrieval_data_annotation.py:93,103, :81 via f-string)."""
    return str(v)


def fmt_4f(v):
    """The retriever's f"{x:.4f}" (train/train_retriever.py:363,368)."""
    return f"{v:.4f}"


def write_float_rows(path, mat, fmt=fmt_str, mode="w"):
    """Rows of floats as text.  Every DISTINCT value is formatted once with `fmt` (the reference's own formatter), the
    rows are assembled by the native r4d_format_lut_rows."""
    lib = _lib.load()
    mat = np.asarray(mat)
    if mat.ndim != 2:
        raise ValueError("write_float_rows expects a 2-D array")
    if mat.shape[1] == 0:                     # empty pool: ' '.join([]) + '\n' per row, like the reference
        with open(path, mode + "b") as f:
            f.write(b"\n" * mat.shape[0])
        return
    codes, strings = _float_codes(mat, fmt)
    write_lut_rows(path, codes, strings.tolist(), mode)


def write_lut_rows(path, codes, strings, mode="w"):
    """Rows of table entries: line r == ' '.join(strings[c] for c in codes[r]).  Assembled by r4d_format_lut_rows."""
    lib = _lib.load()
    codes = np.ascontiguousarray(codes, dtype=np.int32)
    enc = [s.encode("ascii") for s in strings]
    off = np.zeros(len(enc) + 1, dtype=np.int64)
    np.cumsum([len(e) for e in enc], out=off[1:])
    blob = b"".join(enc)
    max_len = max((len(e) for e in enc), default=1)
    nq, n = codes.shape
    per_row = max(1, n * (max_len + 1) + 1)
    step = max(1, _CHUNK_BYTES // per_row)
    with open(path, mode + "b") as f:
        for r0 in range(0, nq, step):
            blk = codes[r0:r0 + step]
            cap = blk.shape[0] * per_row + 1
            buf = ctypes.create_string_buffer(cap)
            got = lib.r4d_format_lut_rows(blk.ctypes.data, blk.shape[0], n, n, blob, off.ctypes.data, len(enc),
                                          ctypes.addressof(buf), cap)
            if got < 0:
                raise _lib.R4DError(f"r4d_format_lut_rows failed ({got}): {_lib.last_error()}")
            f.write(memoryview(buf)[:got])


def write_triplet_scores(path, rows, s_pos, s_neg, fmt=fmt_str):
    """train_score.retrieval: one line f"{i} {out[i, pos]} {out[i, neg]}" per triplet (retrieval_data_annotation.py:81).
    One table holds the decimal row numbers and every distinct score formatted once by `fmt`."""
    rows = np.asarray(rows, dtype=np.int64)
    scores = np.stack([np.asarray(s_pos, dtype=np.float64), np.asarray(s_neg, dtype=np.float64)], axis=1)
    if rows.size == 0:
        open(path, "w").close()
        return
    s_codes, s_strings = _float_codes(scores, fmt)
    n_row_strings = int(rows.max()) + 1
    table = [str(i) for i in range(n_row_strings)] + s_strings.tolist()
    codes = np.concatenate([rows[:, None], s_codes + n_row_strings], axis=1)
    write_lut_rows(path, codes, table)


def jaccard_scores_f64(inter, union):
    """Exact float64 Jaccard from integer counts: identical to Python's len(inter)/len(union) (:14); 0/0 -> 0."""
    inter = np.asarray(inter).astype(np.float64)
    union = np.asarray(union).astype(np.float64)
    out = np.zeros_like(inter)
    np.divide(inter, union, out=out, where=union > 0)
    return out
