"""One C4 call sequence of the postings path (for ncu captures); prints the hand-over counters of the kernel chain and
CUDA-event times of the call per first-stage kernel (option postings_kernel: 0 head kernel, 2 register kernel)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from rag4dyg_b200 import _lib, engine, set_encoder
dev = torch.device("cuda", 0)
mean = float(os.environ.get("MEAN", 1 / 0.45)); nq = int(os.environ.get("NQ", 100000))
pool_ids, pool_off = bench.synth_sets(1_000_000, bench.SEED_POOL, mean)
q_ids, q_off = bench.synth_sets(nq, bench.SEED_QUERY, mean)
bp = set_encoder.encode_csr(pool_ids, pool_off, bench.V_BITS, dev)
index = engine.build_postings(bp)
dq, do = q_ids.to(dev), q_off.to(dev)
ws = torch.empty((_lib.load().r4d_jaccard_topk_postings_workspace_bytes(nq),), dtype=torch.uint8, device=dev)
modes = [int(x) for x in os.environ.get("MODES", "0,2").split(",")]
for mode in modes:
    _lib.set_option("postings_kernel", mode)
    for _ in range(3):
        engine.jaccard_topk_postings(dq, do, index, 10, workspace=ws)
    torch.cuda.synchronize()
    c = ws[:64 * 4].view(torch.int32)[:8].tolist()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        engine.jaccard_topk_postings(dq, do, index, 10, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    print(f"postings_kernel={mode}: counters[work0, listA, heavy_work, light_work, listB, reg_work, listA2] = {c[:7]}  "
          f"{e0.elapsed_time(e1) / 10:.4f} ms per call (L2 warm)")
_lib.set_option("postings_kernel", 0)
print("ok")
