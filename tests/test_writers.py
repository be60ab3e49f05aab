"""CPU: the native text formatters (host functions of libr4d.so) write the exact bytes of the reference's
' '.join(str(x) for x in row) loops (retrieval_data_annotation.py:92-93; train/train_retriever.py:362-363)."""
import numpy as np

from rag4dyg_b200 import build, writers


def setup_module(module):
    build.build_lib()


def test_int_rows_bytes(tmp_path):
    rng = np.random.default_rng(0)
    rows = rng.integers(0, 2**31 - 1, size=(37, 513)).astype(np.int32)
    rows[0, :5] = [0, 9, 10, 99, 100]
    p = tmp_path / "i.txt"
    writers.write_int_rows(str(p), rows)
    ref = "".join(" ".join(str(x) for x in r) + "\n" for r in rows)
    assert p.read_text() == ref
    writers.write_int_rows(str(p), rows[:2], mode="a")           # append mode ('a' of train_retriever.py:365)
    assert p.read_text() == ref + "".join(" ".join(str(x) for x in r) + "\n" for r in rows[:2])


def test_int_rows_chunking(tmp_path, monkeypatch):
    monkeypatch.setattr(writers, "_CHUNK_BYTES", 1000)             # force many native calls
    rows = np.arange(200 * 50, dtype=np.int32).reshape(200, 50)
    p = tmp_path / "c.txt"
    writers.write_int_rows(str(p), rows)
    assert p.read_text() == "".join(" ".join(str(x) for x in r) + "\n" for r in rows)


def test_float_rows_str_and_4f(tmp_path):
    vals = np.array([0.0, 1.0, 0.5, 1 / 3, 0.06666666666666667, 2 / 3, 1e-05, 5e-05, 0.1 + 0.2], dtype=np.float64)
    rng = np.random.default_rng(1)
    m = vals[rng.integers(0, len(vals), size=(21, 300))]
    p = tmp_path / "f.txt"
    writers.write_float_rows(str(p), m, writers.fmt_str)
    assert p.read_text() == "".join(" ".join(str(x) for x in r) + "\n" for r in m)
    m32 = rng.random((9, 77)).astype(np.float32)
    writers.write_float_rows(str(p), m32, writers.fmt_4f)
    assert p.read_text() == "".join(" ".join(f"{x:.4f}" for x in r) + "\n" for r in m32)


def test_empty(tmp_path):
    p = tmp_path / "e.txt"
    writers.write_int_rows(str(p), np.zeros((0, 5), dtype=np.int32))
    assert p.read_text() == ""
    writers.write_int_rows(str(p), np.zeros((2, 0), dtype=np.int32))
    assert p.read_text() == "\n\n"


def test_float_rows_with_zero_columns(tmp_path):
    """An empty pool: the reference writes ' '.join([]) + '\\n' per query row; index and score files must agree."""
    p = tmp_path / "z.txt"
    writers.write_float_rows(str(p), np.zeros((3, 0), dtype=np.float64))
    assert p.read_text() == "\n\n\n"
    writers.write_float_rows(str(p), np.zeros((0, 0), dtype=np.float64))
    assert p.read_text() == ""
