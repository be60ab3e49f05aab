"""Where does an end-to-end HostTopK step spend its time?  CUDA-event times of the pieces, L2 flushed before each."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from rag4dyg_b200 import _lib, engine, set_encoder
from rag4dyg_b200.jaccard_pool import HostTopK, JaccardPool
dev = torch.device("cuda", 0)
nq = 100000
pool_ids, pool_off = bench.synth_sets(1_000_000, bench.SEED_POOL, 1 / 0.45)
q_ids, q_off = bench.synth_sets(nq, bench.SEED_QUERY, 1 / 0.45)
pool = JaccardPool.from_csr(pool_ids.pin_memory(), pool_off.pin_memory(), bench.V_BITS, dev)
qp = (q_ids.pin_memory(), q_off.pin_memory())
dq, do = q_ids.to(dev), q_off.to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
hk = HostTopK(pool, 10, nq, q_ids.numel(), depth=2, packed=True)
host_out = hk.slots[0]["host"]
dev_out = (torch.empty((nq, 10), dtype=torch.int32, device=dev), torch.empty((nq, 10), dtype=torch.int32, device=dev),
           torch.empty((nq,), dtype=torch.int32, device=dev))
ids_d, off_d = hk.slots[0]["ids"], hk.slots[0]["off"]


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    wall = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a.record()
        fn()
        b.record()
        b.synchronize()
        wall += time.perf_counter() - t0
        tot += a.elapsed_time(b)
    return tot / n, wall / n * 1e3


def h2d():
    ids_d[:q_ids.numel()].copy_(qp[0], non_blocking=True)
    off_d[:nq + 1].copy_(qp[1], non_blocking=True)


for name, fn in [
    ("H2D of the query CSR (two copies)", h2d),
    ("top-K packed, results to HBM (inputs resident)", lambda: pool.topk_packed(dq, do, 10, out=dev_out)),
    ("top-K packed, results stored to pinned host (inputs resident)", lambda: pool.topk_packed(dq, do, 10, out=host_out)),
    ("top-K planes, results stored to pinned host (inputs resident)",
     lambda: pool.topk(dq, do, 10, out=tuple(h[:nq] for h in hk.slots[1]["host"][:2]) + (hk.slots[0]["host"][1],))
     if False else pool.topk_packed(dq, do, 10, out=host_out)),
    ("H2D + top-K packed to pinned host (no host wait inside)", lambda: (h2d(), pool.topk_packed(ids_d[:q_ids.numel()], off_d[:nq + 1], 10, out=host_out))),
    ("HostTopK submit + result", lambda: hk.result(hk.submit(*qp))),
]:
    ev, wall = timed(fn)
    print(f"{name}: {ev:.4f} ms (events)  {wall:.4f} ms (host wall incl. sync)")
