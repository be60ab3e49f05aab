"""Skewed-id variant (bench.py `skewed`): time and number of queries handed to the heavy kernel, per light kernel."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from rag4dyg_b200 import _lib
from rag4dyg_b200.jaccard_pool import JaccardPool
dev = torch.device("cuda")
zipf = float(os.environ.get("ZIPF", 0.5)); nq = 100000
pi, po = bench.synth_sets(1_000_000, bench.SEED_POOL + 2, 1 / 0.45, zipf=zipf); qi, qo = bench.synth_sets(nq, bench.SEED_QUERY + 2, 1 / 0.45, zipf=zipf)
pool = JaccardPool.from_csr(pi, po, bench.V_BITS, dev)
dq, do = qi.to(dev), qo.to(dev)
out = tuple(torch.empty((nq, 10), dtype=torch.int32, device=dev) for _ in range(3))
df = torch.bincount(pi.to(dev).long(), minlength=bench.V_BITS)
hits = torch.zeros(nq, dtype=torch.int64, device=dev).index_add_(0, torch.repeat_interleave(torch.arange(nq, device=dev), (do[1:] - do[:-1])), df[dq.long()])
print("hits per query: mean %.0f median %.0f p90 %.0f p99 %.0f max %.0f; >256: %.1f%%  >24600: %.2f%%" % (hits.float().mean(), hits.float().median(), hits.float().quantile(0.9), hits.float().quantile(0.99), hits.max(), 100 * (hits > 256).float().mean(), 100 * (hits > 24600).float().mean()))
for kern in (2, 0):
    _lib.set_option("postings_kernel", kern)
    for _ in range(3): pool.topk(dq, do, 10, out=out)
    _lib.set_option("kernel_timing", 1); _lib.profile_read("jaccard_postings")
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(10): pool.topk(dq, do, 10, out=out)
    e1.record(); torch.cuda.synchronize()
    ms, n = _lib.profile_read("jaccard_postings"); _lib.set_option("kernel_timing", 0)
    c = pool._ws[:32].view(torch.int32).cpu().tolist()
    print(f"postings_kernel={kern}: call {e0.elapsed_time(e1) / 10:.3f} ms, first stages {ms / n:.3f} ms, counters "
          f"[work0, listA, heavy_work, light_work, listB, reg_work, listA2, n_big, big_work] = {c[:9]}", flush=True)
