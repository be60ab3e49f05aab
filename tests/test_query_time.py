"""CPU: rag4dyg_b200.query_time reproduces the reference's get_train_query_time.py output bit for bit
(pool_time in tests/golden/dense_UCI13.npz was written by the unmodified reference script)."""
import lzma
import os

import numpy as np

from conftest import GOLD, extract_inputs
from rag4dyg_b200 import query_time


def test_uci_pool_times_bit_exact(tmp_path):
    extract_inputs("UCI_13", str(tmp_path))
    base = tmp_path / "resources" / "UCI_13" / "12"
    with lzma.open(os.path.join(GOLD, "ml_UCI_13.csv.xz")) as f:
        (base / "ml_UCI_13.csv").write_bytes(f.read())
    got = query_time.get_query_time_all("UCI_13", "12", root=str(tmp_path))
    ref = np.load(os.path.join(GOLD, "dense_UCI13.npz"))["pool_time"]
    assert got.dtype.is_floating_point and got.shape == (1708,)
    assert np.array_equal(got.numpy(), ref)
    assert (tmp_path / "resources" / "UCI_13_train_query_time.pt").exists()


import pytest  # noqa: E402


@pytest.mark.parametrize("ds,T,n", [("hepth", "11", 3965), ("dialog", "15", 7464)])
def test_hepth_dialog_pool_times_bit_exact(tmp_path, ds, T, n):
    """Goldens written by the unmodified reference script (oracle/make_query_time_golden.py): 30-day / unit scales
    (get_train_query_time.py:47-54)."""
    extract_inputs(ds, str(tmp_path))
    base = tmp_path / "resources" / ds / T
    with lzma.open(os.path.join(GOLD, f"ml_{ds}.csv.xz")) as f:
        (base / f"ml_{ds}.csv").write_bytes(f.read())
    got = query_time.get_query_time_all(ds, T, root=str(tmp_path))
    ref = np.load(os.path.join(GOLD, f"query_time_{ds}.npy"))
    assert got.shape == (n,) and np.array_equal(got.numpy(), ref)


def test_root_shim_has_the_reference_cli(tmp_path, monkeypatch):
    """`python get_train_query_time.py <dataset> <timestep>` from a directory holding resources/ (reference :45-58)."""
    import runpy
    import sys
    import torch
    extract_inputs("UCI_13", str(tmp_path))
    base = tmp_path / "resources" / "UCI_13" / "12"
    with lzma.open(os.path.join(GOLD, "ml_UCI_13.csv.xz")) as f:
        (base / "ml_UCI_13.csv").write_bytes(f.read())
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(sys, "argv", ["get_train_query_time.py", "UCI_13", "12"])
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    runpy.run_path(os.path.join(root, "get_train_query_time.py"), run_name="__main__")
    saved = torch.load(tmp_path / "resources" / "UCI_13_train_query_time.pt")
    assert np.array_equal(saved.numpy(), np.load(os.path.join(GOLD, "dense_UCI13.npz"))["pool_time"])
