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
