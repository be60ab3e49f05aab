#!/usr/bin/env python
"""Wall time per stage of the drop-in CLI on the three shipped datasets.
Writes gpurun_out/cli_wall.json.  The reference takes 29 s / 99-116 s / 727-860 s on one core (SURVEY.md section 6,
tests/golden/manifest.json reference_wall_s)."""
import io, json, lzma, os, sys, tarfile, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from rag4dyg_b200 import retrieval_data_annotation as rda

out = {}
man = json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))
torch.zeros(1, device="cuda"); torch.cuda.synchronize()       # CUDA context creation is not part of the stage
for ds, T in (("UCI_13", "12"), ("hepth", "11"), ("dialog", "15"), ("UCI_13", "12"), ("hepth", "11"), ("dialog", "15")):
    with tempfile.TemporaryDirectory() as d:
        raw = lzma.decompress(open(os.path.join(ROOT, "tests", "golden", f"inputs_{ds}.tar.xz"), "rb").read())
        tarfile.open(fileobj=io.BytesIO(raw)).extractall(d, filter="data")
        os.chdir(d)
        np.random.seed(0)
        stages = {}
        t0 = time.perf_counter()
        rda.annotate(ds, T, 0.8, timing=stages)
        wall = time.perf_counter() - t0
        os.chdir(ROOT)
    key = ds if ds not in out else ds + " (warm)"
    gpu = sum(v for k, v in stages.items() if "(GPU)" in k and "write" not in k and "replay" not in k)
    out[key] = {"wall_s": round(wall, 3), "stages_s": {k: round(v, 3) for k, v in stages.items()},
                "gpu_stage_s": round(gpu, 3), "reference_wall_s": man[ds]["reference_wall_s"]}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "cli_wall.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
