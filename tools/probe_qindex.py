#!/usr/bin/env python
"""Stage breakdown of the query-index Jaccard kernel at the bench configuration: times r4d_jaccard_topk with the
stage bypasses of r4d_set_option("jaccard_debug", n) (results are WRONG for n != 0; this is a measurement tool).
Usage: probe_qindex.py [pool] [queries_per_step] [levels...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from rag4dyg_b200 import _lib, engine, set_encoder
    n_pool = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    qs = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    levels = [int(x) for x in sys.argv[3:]] or [0, 1, 2, 3, 0]
    dev = torch.device("cuda", 0)
    pool_ids, pool_off = bench.synth_sets(n_pool, bench.SEED_POOL, 1.0 / 0.45)
    q_ids, q_off = bench.synth_sets(4 * qs, bench.SEED_QUERY, 1.0 / 0.45)
    bp = set_encoder.encode_csr(pool_ids, pool_off, bench.V_BITS, dev)
    bq_all = set_encoder.encode_csr(q_ids, q_off, bench.V_BITS, dev)
    ws = torch.empty(_lib.load().r4d_jaccard_topk_workspace_bytes(qs, n_pool, bench.TOPK), dtype=torch.uint8, device=dev)
    out = {}
    for lv in levels:
        _lib.set_option("jaccard_debug", lv)
        for i in range(3):
            engine.jaccard_topk(bq_all.rows((i % 4) * qs, (i % 4 + 1) * qs), bp, bench.TOPK, workspace=ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for i in range(n):
            engine.jaccard_topk(bq_all.rows((i % 4) * qs, (i % 4 + 1) * qs), bp, bench.TOPK, workspace=ws)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        out.setdefault(str(lv), []).append(ms)
        print(f"debug={lv} ms_per_step={ms:.4f}", flush=True)
    _lib.set_option("jaccard_debug", 0)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
