"""Readers for the index / score files this path writes (consumer half).

The reference's generator dataloader parses them with one Python object per number
(`list(map(int, line.split()))`, `list(map(float, line.split()))`, dataloader/generator.py:32-48; the filter
`len(line) > 0 and not line.isspace()` drops blank lines).  Here the whole file is parsed natively
(r4d_parse_int_rows / r4d_parse_float_rows, multi-threaded strtoll / strtod) into numpy arrays; `as_lists` gives the
reference's list-of-lists back when a caller needs exactly that type.
"""
import ctypes

import numpy as np

from . import _lib


def _parse(path, fn_name, dtype):
    lib = _lib.load()
    with open(path, "rb") as f:
        data = f.read()
    view = np.frombuffer(data, dtype=np.uint8)            # no copy; the parser only reads
    addr = view.ctypes.data if len(data) else 0
    n_fields = ctypes.c_int64(0)
    n_rows = lib.r4d_parse_rows_count(addr, len(data), ctypes.byref(n_fields))
    if n_rows < 0:
        raise _lib.R4DError(f"r4d_parse_rows_count failed ({n_rows}): {_lib.last_error()}")
    row_off = np.zeros(n_rows + 1, dtype=np.int64)
    values = np.empty(max(1, n_fields.value), dtype=dtype)
    got = getattr(lib, fn_name)(addr, len(data), row_off.ctypes.data, values.ctypes.data, n_rows, n_fields.value)
    if got < 0:
        raise _lib.R4DError(f"{fn_name}({path}) failed ({got}): {_lib.last_error()}")
    return values[:n_fields.value], row_off


def _shape(values, row_off):
    """[n_rows, width] when every row has the same number of fields (the files of this path do), else CSR."""
    n_rows = row_off.size - 1
    if n_rows > 0:
        width = int(row_off[1] - row_off[0])
        if width * n_rows == values.size and np.all(np.diff(row_off) == width):
            return values.reshape(n_rows, width)
    return values, row_off


def read_int_rows(path):
    """`*_index.retrieval` / `*_index.gen`: int64 [n_rows, width] (or (values, row_off) for ragged files)."""
    return _shape(*_parse(path, "r4d_parse_int_rows", np.int64))


def read_float_rows(path):
    """`*_score.retrieval` / `*_score.gen`: float64 [n_rows, width], bit-identical to float(text) per field."""
    return _shape(*_parse(path, "r4d_parse_float_rows", np.float64))


def as_lists(rows):
    """The reference's in-memory form: list of per-line lists (dataloader/generator.py:45-47)."""
    if isinstance(rows, tuple):
        values, row_off = rows
        return [values[row_off[i]:row_off[i + 1]].tolist() for i in range(row_off.size - 1)]
    return rows.tolist()
