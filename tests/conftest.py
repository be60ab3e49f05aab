import io
import json
import lzma
import os
import sys
import tarfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
DATASETS = {"UCI_13": "12", "hepth": "11", "dialog": "15"}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: long CPU test, enabled with R4D_SLOW=1")


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLD, "manifest.json")) as f:
        return json.load(f)


def extract_inputs(ds, dst):
    """Unpack tests/golden/inputs_<ds>.tar.xz (the shipped *.link_prediction files) under dst/resources/..."""
    with open(os.path.join(GOLD, f"inputs_{ds}.tar.xz"), "rb") as f:
        raw = lzma.decompress(f.read())
    with tarfile.open(fileobj=io.BytesIO(raw)) as tar:
        tar.extractall(dst, filter="data")


@pytest.fixture()
def dataset_dir(tmp_path):
    def make(ds):
        extract_inputs(ds, str(tmp_path))
        return str(tmp_path)
    return make


def random_sets(rng, n, vocab, mean=2.2, max_len=64, p_empty=0.0, dup=False):
    """Ragged id lists: lengths 1+Geometric, ids uniform (optionally with duplicates), some empty."""
    out = []
    for _ in range(n):
        if rng.random() < p_empty:
            out.append([])
            continue
        ln = int(min(max_len, rng.geometric(1.0 / mean)))
        ids = rng.choice(vocab, size=min(ln, vocab), replace=False).tolist()
        if dup and ids:
            ids = ids + [ids[0]] * int(rng.integers(0, 3))
        out.append(ids)
    return out


def to_csr(sets):
    off = np.zeros(len(sets) + 1, dtype=np.int64)
    flat = []
    for i, s in enumerate(sets):
        flat.extend(s)
        off[i + 1] = len(flat)
    return np.asarray(flat, dtype=np.int32), off
