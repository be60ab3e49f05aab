"""CPU: the native readers (host functions of libr4d.so) return exactly what the reference's dataloader builds with
list(map(int, line.split())) / list(map(float, line.split())) over the non-blank lines (dataloader/generator.py:32-48),
checked on the reference's own golden output files and on ragged / blank-line / CRLF / special-value inputs."""
import lzma
import os

import numpy as np
import pytest

from rag4dyg_b200 import _lib, build, readers, writers

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def setup_module(module):
    build.build_lib()


def _ref_parse(text, conv):
    lines = [line for line in text.splitlines() if (len(line) > 0 and not line.isspace())]   # generator.py:32-35
    return [list(map(conv, line.split())) for line in lines]                                   # generator.py:45-47


def _same_floats(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


@pytest.mark.parametrize("ds", ["UCI_13", "hepth"])
@pytest.mark.parametrize("name", ["test_index.retrieval", "val_index.retrieval", "train_index.gen", "train_index.retrieval"])
def test_golden_index_files(ds, name, tmp_path):
    src = os.path.join(GOLD, ds, name + ".xz")
    if not os.path.exists(src):
        pytest.skip("golden file not stored verbatim for this dataset")
    text = lzma.open(src, "rt").read()
    p = tmp_path / name
    p.write_text(text)
    got = readers.read_int_rows(str(p))
    assert readers.as_lists(got) == _ref_parse(text, int)


@pytest.mark.parametrize("ds", ["UCI_13", "hepth"])
@pytest.mark.parametrize("name", ["test_score.retrieval", "val_score.retrieval", "train_score.gen", "train_score.retrieval"])
def test_golden_score_files(ds, name, tmp_path):
    src = os.path.join(GOLD, ds, name + ".xz")
    if not os.path.exists(src):
        pytest.skip("golden file not stored verbatim for this dataset")
    text = lzma.open(src, "rt").read()
    p = tmp_path / name
    p.write_text(text)
    got = readers.read_float_rows(str(p))
    ref = _ref_parse(text, float)
    lists = readers.as_lists(got)
    assert len(lists) == len(ref) and all(_same_floats(a, b) for a, b in zip(lists, ref))


def test_round_trip_with_the_writers(tmp_path):
    rng = np.random.default_rng(3)
    idx = rng.integers(0, 2**31 - 1, size=(300, 41)).astype(np.int32)
    p = tmp_path / "i.txt"
    writers.write_int_rows(str(p), idx)
    assert np.array_equal(readers.read_int_rows(str(p)), idx.astype(np.int64))
    sc = (rng.integers(0, 9, size=(300, 41)) / rng.integers(1, 9, size=(300, 41))).astype(np.float64)
    writers.write_float_rows(str(p), sc, writers.fmt_str)          # str(np.float64): shortest round-trip repr
    assert _same_floats(readers.read_float_rows(str(p)), sc)
    s32 = rng.random((50, 7)).astype(np.float32)
    writers.write_float_rows(str(p), s32, writers.fmt_4f)
    assert _same_floats(readers.read_float_rows(str(p)), np.array([[float(f"{x:.4f}") for x in r] for r in s32]))


def test_ragged_blank_lines_crlf_and_special_values(tmp_path):
    text = "1 2 3\n\n   \t \n-4\r\n  5   6  \n+7 8\n"
    p = tmp_path / "r.txt"
    p.write_bytes(text.encode())
    got = readers.read_int_rows(str(p))
    assert isinstance(got, tuple) and readers.as_lists(got) == _ref_parse(text, int)
    ftext = "0.0 1.0 0.06666666666666667 1e-05\nnan inf -inf 5E-324 1.7976931348623157e+308\n2.5\n"
    p.write_bytes(ftext.encode())
    lists, ref = readers.as_lists(readers.read_float_rows(str(p))), _ref_parse(ftext, float)
    assert len(lists) == len(ref) and all(_same_floats(a, b) for a, b in zip(lists, ref))
    p.write_bytes(b"")
    assert readers.as_lists(readers.read_int_rows(str(p))) == []
    p.write_bytes(b"7 8 9")                                         # no trailing newline
    assert readers.read_int_rows(str(p)).tolist() == [[7, 8, 9]]


def test_malformed_input_is_an_error(tmp_path):
    p = tmp_path / "bad.txt"
    for payload in (b"1 2 x3\n", b"1_000\n", b"0x10\n"):
        p.write_bytes(payload)
        with pytest.raises(_lib.R4DError):
            readers.read_int_rows(str(p))
    for payload in (b"1.0 abc\n", b"0x1p3\n", b"nan(1)\n", b"1_0.5\n"):
        p.write_bytes(payload)
        with pytest.raises(_lib.R4DError):
            readers.read_float_rows(str(p))


def test_line_structure_follows_python_text_mode_and_splitlines(tmp_path):
    """Bare \\r, \\v, \\f, \\x1c-\\x1e end a line; \\x1f separates fields; non-ASCII bytes are rejected."""
    p = tmp_path / "t.txt"
    raw = b"1 2\r3 4\x0b5\x0c6 7\x1c8\x1d9\x1e10\x1f11 12\r\n\r\n13\n"
    p.write_bytes(raw)
    with open(p, encoding="utf-8") as f:                 # the reference's read: universal newlines, then splitlines()
        text = f.read()
    got = readers.read_int_rows(str(p))
    assert readers.as_lists(got) == _ref_parse(text, int)
    assert readers.as_lists(got) == [[1, 2], [3, 4], [5], [6, 7], [8], [9], [10, 11, 12], [13]]
    for payload in ("1 2\x853\n".encode("utf-8"), "1 2\n".encode("utf-8"), "1\xa02\n".encode("utf-8")):
        p.write_bytes(payload)
        with pytest.raises(_lib.R4DError):
            readers.read_int_rows(str(p))
