"""Call time vs query count on the C4 pool (L2 flushed before every call): whole call (CUDA events) and the light kernel alone."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from rag4dyg_b200 import _lib
from rag4dyg_b200.jaccard_pool import JaccardPool
dev = torch.device("cuda")
pi, po = bench.synth_sets(1_000_000, bench.SEED_POOL, 1 / 0.45)
pool = JaccardPool.from_csr(pi, po, bench.V_BITS, dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for nq in (1, 100, 1000, 4736, 9472, 12500, 25000, 50000, 100000):
    qi, qo = bench.synth_sets(nq, bench.SEED_QUERY, 1 / 0.45)
    dq, do = qi.to(dev), qo.to(dev)
    out = tuple(torch.empty((nq, 10), dtype=torch.int32, device=dev) for _ in range(3))
    for cold in (1, 0):
        _lib.set_option("kernel_timing", 1); _lib.profile_read("jaccard_postings")
        tot = 0.0
        for it in range(13):
            if cold: flush.zero_()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
            pool.topk(dq, do, 10, out=out)
            e1.record(); torch.cuda.synchronize()
            if it >= 3: tot += e0.elapsed_time(e1)
        ms, n = _lib.profile_read("jaccard_postings"); _lib.set_option("kernel_timing", 0)
        print(f"nq={nq:7d} cold_l2={cold}: call {tot / 10 * 1e3:7.1f} us   light kernel {ms / n * 1e3:7.1f} us", flush=True)
