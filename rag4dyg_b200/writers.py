"""Text writers producing byte-identical reference files (retrieval_data_annotation.py:88-103,
train/train_retriever.py:357-368) without a Python-level `str()` per element.

The reference formats every score with `str(np.float64)` (shortest round-trip repr) or f"{x:.4f}".  Jaccard score
matrices hold few distinct values (hard part 6), so each distinct value is formatted once with the
reference's own formatter and rows are assembled by table lookup.
"""
import numpy as np

_INT_LUT = np.empty(0, dtype=object)


def _int_strings(n):
    """object array with str(i) for i < n (grown on demand)."""
    global _INT_LUT
    if _INT_LUT.shape[0] < n:
        _INT_LUT = np.array([str(i) for i in range(n)], dtype=object)
    return _INT_LUT


def write_int_rows(path, rows, mode="w"):
    """Each row of the 2-D integer array as space separated decimals (== ' '.join(str(x) for x in row))."""
    rows = np.asarray(rows)
    if rows.size and rows.min() < 0:
        raise ValueError("negative index in an index file")
    lut = _int_strings(int(rows.max()) + 1 if rows.size else 0)
    with open(path, mode) as f:
        for r in rows:
            f.write(" ".join(lut[r].tolist()) + "\n")


def _float_codes(mat, fmt):
    """(codes int [shape of mat], strings object[n_distinct]) with strings[codes] == fmt(mat) elementwise."""
    mat = np.ascontiguousarray(mat)
    if mat.dtype == np.float64:
        raw = mat.view(np.uint64)
    elif mat.dtype == np.float32:
        raw = mat.view(np.uint32)
    else:
        raise TypeError(f"unsupported score dtype {mat.dtype}")
    uniq, inv = np.unique(raw.ravel(), return_inverse=True)
    vals = uniq.view(mat.dtype)
    strings = np.array([fmt(v) for v in vals], dtype=object)
    return inv.reshape(mat.shape), strings


def fmt_str(v):
    """The reference's `str(x)` on a numpy float64 scalar (retrieval_data_annotation.py:93,103, :81 via f-string)."""
    return str(v)


def fmt_4f(v):
    """The retriever's f"{x:.4f}" (train/train_retriever.py:363,368)."""
    return f"{v:.4f}"


def write_float_rows(path, mat, fmt=fmt_str, mode="w"):
    mat = np.asarray(mat)
    codes, strings = _float_codes(mat, fmt)
    with open(path, mode) as f:
        for r in codes:
            f.write(" ".join(strings[r].tolist()) + "\n")


def jaccard_scores_f64(inter, union):
    """Exact float64 Jaccard from integer counts: identical to Python's len(inter)/len(union) (:14); 0/0 -> 0."""
    inter = np.asarray(inter).astype(np.float64)
    union = np.asarray(union).astype(np.float64)
    out = np.zeros_like(inter)
    np.divide(inter, union, out=out, where=union > 0)
    return out
