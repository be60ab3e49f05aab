"""One C4 call sequence of the postings path (for ncu captures)."""
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
for _ in range(4):
    engine.jaccard_topk_postings(dq, do, index, 10)
torch.cuda.synchronize()
print("ok")
