#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference in this container.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path imports this file.

What it does
------------
For each shipped dataset (UCI_13/12, hepth/11, dialog/15):

1. copies ``/root/reference/resources/<ds>/<T>/*.link_prediction`` into a scratch CWD
   (the reference script reads/writes CWD-relative paths, ``retrieval_data_annotation.py:117-144``,
   and ``/root/reference`` is read-only),
2. runs ``/root/reference/retrieval_data_annotation.py <ds> <T> 0.8`` through ``runpy`` with two
   switches that make its output a function of the data:
     * ``np.argsort`` forced to ``kind='stable'``  -> canonical tie rule (score desc, index asc),
     * ``np.random.seed(0)`` before ``__main__``  -> reproducible ``np.random.choice`` (``:79``),
3. records sha256 + size + line count of the eight output files in ``tests/golden/manifest.json``,
4. stores the *inputs* (xz tarball) and the *small* outputs (xz) under ``tests/golden/`` so the GPU
   box, which has no ``/root/reference``, can run file-level parity tests.

Usage:  python oracle/make_golden.py [UCI_13 hepth dialog]      (dialog takes ~12 min on one core)
"""
import hashlib
import io
import json
import lzma
import os
import runpy
import shutil
import sys
import tarfile
import tempfile
import time

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
DATASETS = {"UCI_13": "12", "hepth": "11", "dialog": "15"}
INPUT_FILES = ["train.link_prediction", "val.link_prediction", "val_gt.link_prediction",
               "test.link_prediction", "test_gt.link_prediction"]
# outputs kept verbatim (xz) when small; all outputs are always hashed
KEEP_MAX_BYTES = 3 << 20


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def run_reference(ds, T, scratch, stable=True, seed=0):
    """Run the reference script unmodified, from `scratch`, with stable argsort + seeded RNG."""
    dst = os.path.join(scratch, "resources", ds, T)
    os.makedirs(dst, exist_ok=True)
    for fn in INPUT_FILES:
        shutil.copy(os.path.join(REF, "resources", ds, T, fn), dst)
    orig_argsort = np.argsort
    if stable:
        def stable_argsort(a, axis=-1, kind=None, order=None, **kw):
            return orig_argsort(a, axis=axis, kind="stable", order=order, **kw)
        np.argsort = stable_argsort
    cwd, argv = os.getcwd(), sys.argv
    try:
        os.chdir(scratch)
        sys.argv = ["retrieval_data_annotation.py", ds, T, "0.8"]
        np.random.seed(seed)
        t0 = time.time()
        runpy.run_path(os.path.join(REF, "retrieval_data_annotation.py"), run_name="__main__")
        dt = time.time() - t0
    finally:
        np.argsort = orig_argsort
        os.chdir(cwd)
        sys.argv = argv
    return dt


def outputs(ds, T):
    r = f"resources/{ds}/{T}/train_retrieval"
    g = f"resources/train_generator/{ds}/{T}/train_gt_topk"
    return [f"{r}/train_index.retrieval", f"{r}/train_score.retrieval",
            f"{r}/test_index.retrieval", f"{r}/test_score.retrieval",
            f"{r}/val_index.retrieval", f"{r}/val_score.retrieval",
            f"{g}/train_index.gen", f"{g}/train_score.gen"]


def main():
    which = sys.argv[1:] or list(DATASETS)
    os.makedirs(GOLD, exist_ok=True)
    man_path = os.path.join(GOLD, "manifest.json")
    manifest = json.load(open(man_path)) if os.path.exists(man_path) else {}
    for ds in which:
        T = DATASETS[ds]
        # inputs -> xz tarball (data fixtures, not reference source)
        buf = io.BytesIO()
        with tarfile.open(fileobj=buf, mode="w") as tar:
            for fn in INPUT_FILES:
                tar.add(os.path.join(REF, "resources", ds, T, fn), arcname=f"resources/{ds}/{T}/{fn}")
        with open(os.path.join(GOLD, f"inputs_{ds}.tar.xz"), "wb") as f:
            f.write(lzma.compress(buf.getvalue(), preset=9))
        scratch = tempfile.mkdtemp(prefix=f"r4d_golden_{ds}_")
        dt = run_reference(ds, T, scratch)
        entry = {"timestep": T, "threshold": 0.8, "seed": 0, "argsort": "stable",
                 "reference_wall_s": round(dt, 2), "numpy": np.__version__, "files": {}}
        for rel in outputs(ds, T):
            p = os.path.join(scratch, rel)
            n_lines = sum(1 for _ in open(p, "rb"))
            entry["files"][rel] = {"sha256": sha256(p), "bytes": os.path.getsize(p), "lines": n_lines}
            if os.path.getsize(p) <= KEEP_MAX_BYTES:
                out = os.path.join(GOLD, ds, os.path.basename(rel) + ".xz")
                os.makedirs(os.path.dirname(out), exist_ok=True)
                with open(p, "rb") as fi, open(out, "wb") as fo:
                    fo.write(lzma.compress(fi.read(), preset=9))
        manifest[ds] = entry
        with open(man_path, "w") as f:
            json.dump(manifest, f, indent=1, sort_keys=True)
        print(f"[{ds}] reference wall {dt:.1f}s; hashed {len(entry['files'])} files", flush=True)
        shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__":
    main()
