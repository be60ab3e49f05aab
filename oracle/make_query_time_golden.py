"""Golden pool-time vectors for hepth / dialog from the UNMODIFIED reference script get_train_query_time.py
(/root/reference/get_train_query_time.py:17-58, scales 30 days / 1), run here via runpy from a scratch directory.

    python oracle/make_query_time_golden.py

Writes tests/golden/query_time_<ds>.npy (float32 [N], the tensor the reference saves as resources/<ds>_train_query_time.pt)
and tests/golden/ml_<ds>.csv.xz (the four columns of the reference's edge list that the script reads: u, i, ts,
timestamp).  UCI_13's vector already lives in tests/golden/dense_UCI13.npz (oracle/make_dense_golden.py).
TEST INFRASTRUCTURE ONLY."""
import lzma
import os
import runpy
import shutil
import sys
import tempfile

import numpy as np
import pandas as pd
import torch

REF = os.environ.get("R4D_REFERENCE_ROOT", "/root/reference")
GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
DATASETS = {"hepth": "11", "dialog": "15"}


def main():
    for ds, T in DATASETS.items():
        with tempfile.TemporaryDirectory() as d:
            dst = os.path.join(d, "resources", ds, T)
            os.makedirs(dst)
            for name in (f"ml_{ds}.csv", "train.link_prediction"):
                shutil.copy(os.path.join(REF, "resources", ds, T, name), dst)
            cwd, argv = os.getcwd(), sys.argv
            os.chdir(d)
            sys.argv = ["get_train_query_time.py", ds, T]
            try:
                runpy.run_path(os.path.join(REF, "get_train_query_time.py"), run_name="__main__")
            finally:
                os.chdir(cwd)
                sys.argv = argv
            t = torch.load(os.path.join(d, "resources", f"{ds}_train_query_time.pt")).numpy()
            np.save(os.path.join(GOLD, f"query_time_{ds}.npy"), t)
            cols = pd.read_csv(os.path.join(dst, f"ml_{ds}.csv"), usecols=["u", "i", "ts", "timestamp"])
            with lzma.open(os.path.join(GOLD, f"ml_{ds}.csv.xz"), "wt") as f:
                cols.to_csv(f, index=False)
            print(ds, t.shape, t.dtype, float(t.min()), float(t.max()))


if __name__ == "__main__":
    main()
