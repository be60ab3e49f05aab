"""Timing probe of the postings path on the C4 workload (1M-set pool x 100k queries) and the x-like variant."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rag4dyg_b200 import _lib, engine, set_encoder  # noqa: E402

dev = torch.device("cuda", 0)
out = {}
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timed(fn, steps=10, warm=3, do_flush=True):
    for _ in range(warm):
        fn()
    tot = 0.0
    for _ in range(steps):
        if do_flush:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / steps


def run(mean, n_pool, n_q, tag, verify=16):
    pool_ids, pool_off = bench.synth_sets(n_pool, bench.SEED_POOL, mean)
    q_ids, q_off = bench.synth_sets(n_q, bench.SEED_QUERY, mean)
    t0 = time.perf_counter()
    bp = set_encoder.encode_csr(pool_ids, pool_off, bench.V_BITS, dev)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    index = engine.build_postings(bp)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    dq, do = q_ids.to(dev), q_off.to(dev)
    ws = torch.empty(_lib.load().r4d_jaccard_topk_postings_workspace_bytes(n_q), dtype=torch.uint8, device=dev)
    outb = tuple(torch.empty((n_q, 10), dtype=torch.int32, device=dev) for _ in range(3))

    def call():
        engine.jaccard_topk_postings(dq, do, index, 10, workspace=ws, out=outb)
    ms_cold = timed(call)
    ms_warm = timed(call, do_flush=False)
    heavy = int(ws[4:8].view(torch.int32).item())
    _lib.set_option("kernel_timing", 1)
    _lib.profile_read("jaccard_postings")
    for _ in range(5):
        call()
    kms, kn = _lib.profile_read("jaccard_postings")
    _lib.set_option("kernel_timing", 0)
    r = {"encode_s": t1 - t0, "build_s": t2 - t1, "index_MB": index.blob.numel() / 1e6, "postings": index.nnz,
         "ms_per_call_l2_flushed": ms_cold, "ms_per_call_warm": ms_warm, "light_kernel_ms_warm": kms / max(kn, 1),
         "heavy_queries": heavy, "pairs_per_s_flushed": n_pool * n_q / (ms_cold * 1e-3),
         "pairs_per_s_warm": n_pool * n_q / (ms_warm * 1e-3)}
    if verify:
        from oracle import jaccard_oracle as jo
        sel = np.linspace(0, n_q - 1, verify).astype(np.int64)
        qo = q_off.numpy()
        ids = np.concatenate([q_ids.numpy()[qo[s]:qo[s + 1]] for s in sel])
        off = np.zeros(verify + 1, np.int64)
        off[1:] = np.cumsum([qo[s + 1] - qo[s] for s in sel])
        oi, ou, ox = jo.c_topk(ids, off, pool_ids.numpy(), pool_off.numpy(), 10)
        gi, gu, gx = (t.cpu().numpy()[sel] for t in outb)
        r["verified_rows"] = int(verify)
        r["verified"] = bool(np.array_equal(gx, ox) and np.array_equal(gi, oi) and np.array_equal(gu, ou))
    out[tag] = r
    print(tag, json.dumps(r), flush=True)
    return bp, index


if __name__ == "__main__":
    n_pool = int(os.environ.get("POOL", 1_000_000))
    run(1 / 0.45, n_pool, 100_000, "c4_y_like")
    run(20.0, n_pool, 8192, "x_like")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_postings.json"), "w"), indent=1)
