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


# ---------------------------------------------------------------------------------------------- device-side assembly
_PIN = {}     # two reusable pinned staging buffers per process (allocating pinned memory costs more than the copy)
_PIN_BYTES = 32 << 20


def _drain_to_file(f, text_dev):
    """Device text -> file through two pinned staging buffers: the copy of chunk i + 1 overlaps the write of chunk i."""
    import torch
    if "bufs" not in _PIN:
        _PIN["bufs"] = [torch.empty(_PIN_BYTES, dtype=torch.uint8).pin_memory() for _ in range(2)]
        _PIN["ev"] = [torch.cuda.Event() for _ in range(2)]
    bufs, evs = _PIN["bufs"], _PIN["ev"]
    total = text_dev.numel()
    chunks = [(a, min(a + _PIN_BYTES, total)) for a in range(0, total, _PIN_BYTES)]
    for i, (a, b) in enumerate(chunks[:1]):
        bufs[0][:b - a].copy_(text_dev[a:b], non_blocking=True)
        evs[0].record()
    for i, (a, b) in enumerate(chunks):
        if i + 1 < len(chunks):
            a2, b2 = chunks[i + 1]
            bufs[(i + 1) & 1][:b2 - a2].copy_(text_dev[a2:b2], non_blocking=True)
            evs[(i + 1) & 1].record()
        evs[i & 1].synchronize()
        f.write(memoryview(bufs[i & 1].numpy())[:b - a])


def _format_rows_device(vals, lut_blob=None, lut_off=None):
    """int32 CUDA matrix (+ optional device string table) -> uint8 CUDA tensor holding the file's bytes."""
    import torch
    lib = _lib.load()
    if not (isinstance(vals, torch.Tensor) and vals.is_cuda and vals.dtype == torch.int32 and vals.dim() == 2):
        raise _lib.R4DError("device text assembly expects a 2-D int32 CUDA tensor")
    vals = vals.contiguous()
    nq, n = vals.shape
    dev = vals.device
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream().cuda_stream
        row_off = torch.empty((nq + 1,), dtype=torch.int64, device=dev)
        status = torch.zeros((1,), dtype=torch.int32, device=dev)
        n_codes = 0 if lut_off is None else lut_off.numel() - 1
        p_off = 0 if lut_off is None else lut_off.data_ptr()
        p_blob = 0 if lut_blob is None else lut_blob.data_ptr()
        _lib.check(lib.r4d_format_rows_device_sizes(vals.data_ptr(), nq, n, n, p_off, n_codes, row_off.data_ptr(),
                                                    status.data_ptr(), st), "r4d_format_rows_device_sizes")
        total = int(row_off[-1].item())
        if int(status.item()) != 0:
            raise _lib.R4DError("device text assembly: a code lies outside the string table")
        text = torch.empty((max(total, 1),), dtype=torch.uint8, device=dev)
        _lib.check(lib.r4d_format_rows_device(vals.data_ptr(), nq, n, n, p_blob, p_off, n_codes, row_off.data_ptr(),
                                              text.data_ptr(), st), "r4d_format_rows_device")
    return text[:total]


def write_int_rows_device(path, rows, mode="w"):
    """write_int_rows for an int32 CUDA matrix: the text is assembled on the GPU (r4d_format_rows_device)."""
    text = _format_rows_device(rows)
    with open(path, mode + "b") as f:
        _drain_to_file(f, text)


def write_float_rows_device(path, mat, fmt=fmt_str, mode="w"):
    """write_float_rows for a float64 / float32 CUDA matrix: distinct values are found on the device, each is formatted
    once on the host with `fmt` (the reference's own formatter), the rows are assembled on the GPU."""
    import torch
    if mat.shape[1] == 0:
        with open(path, mode + "b") as f:
            f.write(b"\n" * mat.shape[0])
        return
    raw = mat.contiguous().view(torch.int64 if mat.dtype == torch.float64 else torch.int32)     # bit patterns: -0.0 != 0.0
    uniq, inv = torch.unique(raw, return_inverse=True)
    vals = uniq.cpu().numpy().view(np.float64 if mat.dtype == torch.float64 else np.float32)
    enc = [fmt(v).encode("ascii") for v in vals]
    off = np.zeros(len(enc) + 1, dtype=np.int64)
    np.cumsum([len(e) for e in enc], out=off[1:])
    blob = torch.frombuffer(bytearray(b"".join(enc) or b"\0"), dtype=torch.uint8).to(mat.device)
    text = _format_rows_device(inv.to(torch.int32), blob, torch.from_numpy(off).to(mat.device))
    with open(path, mode + "b") as f:
        _drain_to_file(f, text)


def jaccard_scores_f64(inter, union):
    """Exact float64 Jaccard from integer counts: identical to Python's len(inter)/len(union) (:14); 0/0 -> 0."""
    inter = np.asarray(inter).astype(np.float64)
    union = np.asarray(union).astype(np.float64)
    out = np.zeros_like(inter)
    np.divide(inter, union, out=out, where=union > 0)
    return out
