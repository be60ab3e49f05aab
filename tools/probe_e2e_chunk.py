"""Kernel time of a C4 top-K call by destination (HBM / pinned host) and queries per grab (option postings_chunk)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from rag4dyg_b200 import _lib
from rag4dyg_b200.jaccard_pool import JaccardPool
dev = torch.device("cuda", 0)
nq = 100000
pool_ids, pool_off = bench.synth_sets(1_000_000, bench.SEED_POOL, 1 / 0.45)
q_ids, q_off = bench.synth_sets(nq, bench.SEED_QUERY, 1 / 0.45)
pool = JaccardPool.from_csr(pool_ids.pin_memory(), pool_off.pin_memory(), bench.V_BITS, dev)
dq, do = q_ids.to(dev), q_off.to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
mk = lambda pin: tuple((torch.empty(s, dtype=torch.int32).pin_memory() if pin else torch.empty(s, dtype=torch.int32, device=dev))
                       for s in ((nq, 10), (nq, 10), (nq,)))
outs = {"HBM": mk(False), "pinned host": mk(True)}
for chunk in (0, 1, 2, 3, 4, 6, 8):
    _lib.set_option("postings_chunk", chunk)
    line = f"chunk {chunk}:"
    for name, out in outs.items():
        for _ in range(3):
            pool.topk_packed(dq, do, 10, out=out)
        tot = 0.0
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            pool.topk_packed(dq, do, 10, out=out)
            b.record()
            b.synchronize()
            tot += a.elapsed_time(b)
        line += f"  {name} {tot / 10:.4f} ms"
    print(line)
