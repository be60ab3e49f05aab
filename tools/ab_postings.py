"""A/B probe (same box, alternating, L2 flushed before every call) of a postings-path option on the C4 data.
Usage: python tools/ab_postings.py key v0 v1 [nq] [reps]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from rag4dyg_b200 import _lib
from rag4dyg_b200.jaccard_pool import JaccardPool
key, v0, v1 = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
nq = int(sys.argv[4]) if len(sys.argv) > 4 else 100000
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
dev = torch.device("cuda")
pi, po = bench.synth_sets(1_000_000, bench.SEED_POOL, 1 / 0.45); qi, qo = bench.synth_sets(nq, bench.SEED_QUERY, 1 / 0.45)
pool = JaccardPool.from_csr(pi, po, bench.V_BITS, dev)
dq, do = qi.to(dev), qo.to(dev)
out = tuple(torch.empty((nq, 10), dtype=torch.int32, device=dev) for _ in range(3))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(steps):
    tot = 0.0
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
        pool.topk(dq, do, 10, out=out)
        e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / steps
for r in range(reps):
    for v in (v0, v1):
        _lib.set_option(key, v); run(3); ms = run(20)
        print(f"rep {r} nq={nq} {key}={v}: {ms * 1e3:.1f} us/call", flush=True)
