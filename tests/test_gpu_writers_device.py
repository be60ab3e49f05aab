"""GPU: the device-side text assembly (r4d_format_rows_device) writes exactly the bytes of the reference's
' '.join(str(x) for x in row) + '\\n' (retrieval_data_annotation.py:92-93) — compared with the host writers, which
tests/test_writers.py pins against the Python expression itself."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from rag4dyg_b200 import writers  # noqa: E402


@pytest.mark.parametrize("nq,n", [(1, 1), (3, 5), (17, 1024), (5, 1025), (9, 4097), (300, 41), (2, 0), (0, 7)])
def test_int_rows_device_equals_python_join(tmp_path, nq, n):
    rng = np.random.default_rng(nq * 31 + n)
    m = rng.integers(0, 2**31 - 1, size=(nq, n)).astype(np.int32)
    if m.size:
        m.flat[0] = 0
        m.flat[-1] = 2**31 - 1
        m.flat[m.size // 2] = -12345          # negative values are formatted like str(int) too
    p = tmp_path / "d.txt"
    writers.write_int_rows_device(str(p), torch.from_numpy(m).cuda())
    assert p.read_text() == "".join(" ".join(str(int(x)) for x in r) + "\n" for r in m)


@pytest.mark.parametrize("dtype,fmt", [(np.float64, writers.fmt_str), (np.float32, writers.fmt_4f)])
def test_float_rows_device_equals_host_writer(tmp_path, dtype, fmt):
    rng = np.random.default_rng(4)
    m = (rng.integers(0, 7, size=(257, 1333)) / rng.integers(1, 9, size=(257, 1333))).astype(dtype)
    m[0, :5] = [0.0, 1.0, 1.0 / 15.0, 1e-5, 123456.789]
    a, b = tmp_path / "dev.txt", tmp_path / "host.txt"
    writers.write_float_rows_device(str(a), torch.from_numpy(m).cuda(), fmt)
    writers.write_float_rows(str(b), m, fmt)
    assert a.read_bytes() == b.read_bytes()
    assert a.read_text().splitlines()[0].split()[:5] == [fmt(v) for v in m[0, :5]]


def test_large_file_goes_through_several_staging_chunks(tmp_path, monkeypatch):
    monkeypatch.setattr(writers, "_PIN_BYTES", 1 << 16)
    monkeypatch.setattr(writers, "_PIN", {})
    m = np.arange(200 * 3000, dtype=np.int32).reshape(200, 3000)
    p = tmp_path / "big.txt"
    writers.write_int_rows_device(str(p), torch.from_numpy(m).cuda())
    assert p.read_text() == "".join(" ".join(map(str, r)) + "\n" for r in m)
    writers.write_float_rows_device(str(p), torch.zeros((3, 0), dtype=torch.float64, device="cuda"))
    assert p.read_text() == "\n\n\n"
