#!/usr/bin/env python
"""bench.py — query-pool pairs scored + top-K'd per second (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scorers jaccard,dense]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads (BASELINE.json configs[3] and [4]), synthetic data, fixed TOTAL size (strong
scaling: the pool is sharded row-wise over the N ranks, queries replicated, local top-K lists exchanged over NVLink
and merged):
  jaccard : 1,000,000-set pool x 100,000 queries, 20,000-node vocab (W = 625 words), K = 10.
            A step = ONE r4d_jaccard_topk call of 32,768 queries against the WHOLE 1M pool + top-K (+ exchange + merge);
            the library serves it as four 8,192-query launch sequences, each streaming the pool's bitsets once.
  dense   : 10,000,000 x 768 pool embeddings x 100,000 queries, K = 10, exp(-lambda|dt|) epilogue, bf16 tensor cores.
            A step = one 8,192-query batch against the whole 10M pool.
One JSON line: the Jaccard scorer is the headline (`value`), the dense scorer is reported under "dense".
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

V_BITS = 20000
POOL_N = 1_000_000
QUERY_N = 100_000
TOPK = 10
Q_STEP = 8192
JQ_STEP = 32768          # Jaccard step: four 8,192-query launch sequences inside ONE C-ABI call
JQ_LAUNCH = 8192         # query rows the library serves per pool stream (SQ_QB in csrc/jaccard_common.cuh)
DENSE_POOL_N = 10_000_000
DENSE_D = 768
DENSE_LAMBDA = 1e-4
SEED_POOL, SEED_QUERY = 1234, 5678
SEED_DPOOL, SEED_DQUERY = 4321, 8765


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scorers", default="jaccard,dense")
    ap.add_argument("--pool", type=int, default=POOL_N)
    ap.add_argument("--dense-pool", type=int, default=DENSE_POOL_N)
    ap.add_argument("--queries-per-step", type=int, default=Q_STEP, help="dense scorer: queries per step")
    ap.add_argument("--jaccard-queries-per-step", type=int, default=JQ_STEP,
                    help="Jaccard scorer: queries per step (one r4d_jaccard_topk call = ceil(q / 8192) launch sequences)")
    ap.add_argument("--mean-set", type=float, default=1.0 / 0.45, help="mean set size (y-like 2.2; x-like 20)")
    ap.add_argument("--dense-d", type=int, default=DENSE_D, help="embedding width (experiments; the metric uses 768)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: fused NVLink peer-store exchange (symmetric memory) or NCCL all-gathers")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ synthetic data
def synth_sets(n, seed, mean, vocab=V_BITS, max_len=64):
    """CSR id lists: |set| = min(64, Geometric(1/mean)) (support 1.., as the shipped label sets), ids uniform.
    (Drawn with replacement; the rare duplicate collapses in the set encoder exactly like Python's set().)"""
    import torch
    g = torch.Generator().manual_seed(seed)
    p = 1.0 / mean
    u = torch.rand(n, generator=g, dtype=torch.float64).clamp_(min=1e-300)
    lens = torch.floor(torch.log(u) / math.log(1.0 - p)).to(torch.int64) + 1
    lens.clamp_(min=1, max=max_len)
    off = torch.zeros(n + 1, dtype=torch.int64)
    torch.cumsum(lens, 0, out=off[1:])
    ids = torch.randint(0, vocab, (int(off[-1]),), generator=g, dtype=torch.int32)
    return ids, off


def csr_rows(ids, off, a, b):
    o = off[a:b + 1]
    return ids[int(o[0]):int(o[-1])].contiguous(), (o - o[0]).contiguous()


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass
        sm, smax, reasons = [], 0.0, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = max(smax, float(f[2]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU baselines
def _token_lists(ids, off, a, b):
    ids = ids.numpy() if hasattr(ids, "numpy") else ids
    off = off.numpy() if hasattr(off, "numpy") else off
    return [[str(t) for t in ids[off[r]:off[r + 1]]] for r in range(a, b)]


def _cpu_port_job(args):
    """One worker: the reference's algorithm on its query slice (Python sets, retrieval_data_annotation.py:36-41,
    then np.argsort(-row)[:k], :101)."""
    q_lists, p_lists, k = args
    from oracle import jaccard_oracle as jo
    t0 = time.perf_counter()
    m = jo.occurrence_matrix(q_lists, p_lists)
    jo.topk_stable(m, k)
    return time.perf_counter() - t0


def cpu_port_sample(pool_ids, pool_off, q_ids, q_off, n_q, n_p, procs):
    """pairs/s of the oracle port on `procs` host processes over disjoint query slices (bounded sample)."""
    p_lists = _token_lists(pool_ids, pool_off, 0, n_p)
    per = max(1, n_q // procs)
    jobs = [(_token_lists(q_ids, q_off, w * per, (w + 1) * per), p_lists, TOPK) for w in range(procs)]
    t0 = time.perf_counter()
    if procs == 1:
        _cpu_port_job(jobs[0])
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            pool.map(_cpu_port_job, jobs)
    dt = time.perf_counter() - t0
    return per * procs * n_p / dt, dt, per * procs


def cpu_c_port_sample(pool_ids, pool_off, q_ids, q_off, n_q, n_p):
    from oracle import jaccard_oracle as jo
    qi, qo = csr_rows(q_ids, q_off, 0, n_q)
    pi, po = csr_rows(pool_ids, pool_off, 0, n_p)
    t0 = time.perf_counter()
    jo.c_topk(qi.numpy(), qo.numpy(), pi.numpy(), po.numpy(), TOPK)
    dt = time.perf_counter() - t0
    return n_q * n_p / dt, dt


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is pure Python
    (nothing to compile into oracle/_ref), so this times the oracle port — the same Python-set double loop +
    argsort — on all host cores.  Rank 0 only."""
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    n_p_sample, q_per_worker = 100_000, 4
    pool_ids, pool_off = synth_sets(n_p_sample, SEED_POOL, args.mean_set)
    q_ids, q_off = synth_sets(JQ_LAUNCH, SEED_QUERY, args.mean_set)
    times = []
    for step in range(args.warmup + args.steps):
        rate, dt, nq = cpu_port_sample(pool_ids, pool_off, q_ids, q_off, q_per_worker * procs, n_p_sample, procs)
        if step >= args.warmup:
            times.append(dt)
    pairs = q_per_worker * procs * n_p_sample
    total = sum(times)
    value = pairs * args.steps / total
    sample = (f"{q_per_worker * procs} queries x {n_p_sample} pool sets per step (same distribution/seeds as the GPU "
              f"workload), Python-set Jaccard + stable argsort top-{TOPK}, {procs} processes on disjoint query slices")
    emit({
        "impl": "reference", "metric": "query-pool pairs scored+top-K/sec (Jaccard)", "value": value,
        "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int (Python set algebra) / float64 divide", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def workload_config(args, world):
    return {"workload": f"synthetic Jaccard top-K: {args.pool:,}-set pool x {QUERY_N:,} queries, vocab {V_BITS:,} "
                        f"(W=625 uint32 words), K={TOPK}; step = one r4d_jaccard_topk call: {args.jaccard_queries_per_step:,} queries "
                        f"vs the whole pool (the library serves it as {JQ_LAUNCH:,}-query launch sequences)",
            "pool": args.pool, "queries_total": QUERY_N, "queries_per_step": args.jaccard_queries_per_step, "vocab": V_BITS,
            "k": TOPK, "mean_set_size": round(args.mean_set, 3), "parallelism": f"pool-sharded x{world}",
            "l2": "inputs larger than L2 (pool bitsets 2.56 GB / n_gpus; a different query batch every step)"}


# ------------------------------------------------------------------------------------------ our arm
_REAL_STDOUT = None


def emit(obj):
    """The ONE JSON line of the contract, written to the process' original stdout."""
    line = (json.dumps(obj) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, line)


def main():
    global _REAL_STDOUT
    args = parse_args()
    # everything else that reaches fd 1 (library banners such as "NCCL version ...", stray prints) goes to stderr
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from rag4dyg_b200 import _lib, engine, set_encoder, sharded

    _lib.require_device()  # fail loudly: no CPU fallback
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # the version banner goes to stdout and would precede the JSON line
        dist.init_process_group("nccl", device_id=dev)
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    scorers = args.scorers.split(",")
    K, W = args.steps, args.warmup
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp))

    def dram_traffic(key, **cfg):
        # ncu-measured DRAM bytes per launch; only valid for the exact configuration it was captured on
        t = traffic.get(key)
        if t and all(t.get(k) == v for k, v in cfg.items()):
            return t["dram_bytes"]
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(step_fn, sampler=None, steps=None, sampler_on=False):
        """W warm-up steps, then `steps` (default K) timed steps between barrier+synchronize; device time by CUDA
        events on the launch stream, max over ranks.  Returns (ms per step * steps, launches, clocks)."""
        n = K if steps is None else steps
        for i in range(W):
            step_fn(i)
        barrier()
        sampler_on = sampler_on or sampler is not None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        engine.reset_launch_count()
        e0.record()
        for i in range(n):
            step_fn(W + i)
        e1.record()
        barrier()
        launches = engine.launch_count()
        ms = max_over_ranks(e0.elapsed_time(e1))
        clocks = None
        if sampler_on:
            # nvidia-smi samples every 100 ms: when the timed region is shorter than ~1.2 s the SAME step keeps running
            # (untimed) until the sampler has seen that much load; the step count is derived from the max-over-ranks
            # time, so every rank runs the same number (the step contains the exchange at N > 1)
            extra = 0
            if ms < 1200.0:
                extra = min(20000, int(math.ceil((1200.0 - ms) / max(ms / n, 1e-3))))
                for i in range(extra):
                    step_fn(W + n + i)
                barrier()
            if sampler:
                clocks = sampler.stop()
                clocks["sampled_over"] = (f"the {n} timed steps" if extra == 0 else
                                          f"the {n} timed steps + {extra} further identical steps run right after them "
                                          f"(the timed region alone gives the 100 ms sampler too few samples)")
        return ms, launches, clocks

    K_AUX = min(K, 5)   # auxiliary measurements (dense case, x-like variant) never run more than 5 timed steps

    out = {}
    # =============================================================== Jaccard
    if "jaccard" in scorers:
        n_pool, qs = args.pool, args.jaccard_queries_per_step
        ql = min(qs, JQ_LAUNCH)                                 # query rows per launch sequence
        n_launch = (qs + JQ_LAUNCH - 1) // JQ_LAUNCH            # pool streams per step
        lo, hi = rank * n_pool // world, (rank + 1) * n_pool // world
        pool_ids, pool_off = synth_sets(n_pool, SEED_POOL, args.mean_set)
        q_ids, q_off = synth_sets(QUERY_N, SEED_QUERY, args.mean_set)
        sh_ids, sh_off = csr_rows(pool_ids, pool_off, lo, hi)
        sh_ids_pin, sh_off_pin = sh_ids.pin_memory(), sh_off.pin_memory()
        bp = set_encoder.encode_csr(sh_ids_pin, sh_off_pin, V_BITS, dev)      # pool shard resident in HBM
        bq_all = set_encoder.encode_csr(q_ids, q_off, V_BITS, dev)            # all 100k queries resident (250 MB)
        n_batches = max(1, QUERY_N // qs)
        ws = torch.empty(_lib.load().r4d_jaccard_topk_workspace_bytes(qs, hi - lo, TOPK), dtype=torch.uint8, device=dev)
        result = {}
        jex, exchange_used = None, "none (single GPU)"
        if world > 1:
            exchange_used = "nccl all-gather"
            if args.exchange == "p2p":
                try:
                    jex = sharded.P2PExchange(qs, TOPK, 3)
                    exchange_used = "fused NVLink peer stores from the merge kernel + 1 symmetric-memory barrier"
                except Exception as e:          # symmetric memory not available: the NCCL collective is used instead
                    exchange_used = f"nccl all-gather (p2p unavailable: {type(e).__name__})"

        def step_resident(i):
            b = i % n_batches
            bq = bq_all.rows(b * qs, (b + 1) * qs)
            # local fused top-K on the shard, then ONE exchange of [Q, K] candidates + merge
            result["last"] = sharded.jaccard_topk_sharded(bq, bp, TOPK, pool_base=lo, workspace=ws, exchange=jex)

        sampler = ClockSampler(local_rank) if rank == 0 else None
        ms, launches, clocks = timed(step_resident, sampler, sampler_on=True)
        pairs_per_step = qs * n_pool
        value = pairs_per_step * K / (ms * 1e-3)

        # the C-ABI call alone (r4d_jaccard_topk on this rank's shard: index build + main kernel + stripe merge), then the
        # dominant kernel alone: the library brackets it with CUDA events on its launch stream ("kernel_timing")
        def kernel_only(i):
            b = i % n_batches
            engine.jaccard_topk(bq_all.rows(b * qs, (b + 1) * qs), bp, TOPK, pool_base=lo, workspace=ws)
        k_ms, _, _ = timed(kernel_only)
        _lib.set_option("kernel_timing", 1)
        _lib.profile_read("jaccard_qindex")
        for i in range(K):
            kernel_only(W + i)
        main_ms, main_n = _lib.profile_read("jaccard_qindex")
        _lib.set_option("kernel_timing", 0)
        main_s = max_over_ranks(main_ms / max(main_n, 1)) * 1e-3
        # dense case: the bitset-streaming kernel with zero-span skipping disabled executes every algorithmic word-op
        # (one 8,192-query launch per step: it is ~1000x slower than the index path)
        def kernel_dense_case(i):
            engine.jaccard_topk(bq_all.rows((i % 4) * ql, (i % 4 + 1) * ql), bp, TOPK, pool_base=lo, workspace=ws)
        _lib.set_option("jaccard_skip_zero", 0)
        kd_ms, _, _ = timed(kernel_dense_case, steps=K_AUX)
        _lib.set_option("jaccard_skip_zero", 1)
        kd_s = kd_ms * 1e-3 / K_AUX
        words = 625
        word_ops = ql * (hi - lo) * words                      # algorithmic AND+POPC word-ops per launch (W per pair)
        sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz", 1965.0)
        popc_peak = 148 * 16 * sm_max * 1e6                    # 16 POPC lanes/clk/SM (measured 15.8, tools/microbench.cu)
        two_pipe = 148 * (64 / 2.125) * sm_max * 1e6           # CSA kernel: 17 ALU ops (64 lanes/clk/SM) + 4 POPC per 8 words
        pool_bytes = (hi - lo) * words * 4                     # the main kernel streams every pool row once per launch
        call_bytes = (n_launch * (hi - lo) + qs) * words * 4 + qs * TOPK * 12   # whole call: one pool stream per launch
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        roofline = {"bound": "hbm", "achieved": pool_bytes / main_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": pool_bytes / main_s / 1e9 / hbm_peak,
                    "traffic": dram_traffic("jaccard_topk", queries=ql, pool=hi - lo) if world == 1 else None,
                    "kernel": "r4d::jaccard_qindex_kernel<ROW1> (query-side bit index in smem, pool bitsets streamed once "
                              "per 8,192-query batch by per-warp TMA bulk copies)",
                    "kernel_ms": main_s * 1e3, "launches_timed": int(main_n),
                    "algorithmic_bytes": pool_bytes,
                    "algorithmic_bytes_what": "4*W B per pool row of the shard, each row read once per launch (W = 625 words)",
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy bandwidth)" if peaks else "fallback 6650 GB/s",
                    "call": {"what": f"whole r4d_jaccard_topk call = {n_launch} x (qindex_kernel + memset + main kernel + merge kernel)",
                             "ms": k_ms / K, "algorithmic_bytes": call_bytes, "achieved_GBps": call_bytes / (k_ms * 1e-3 / K) / 1e9,
                             "frac": call_bytes / (k_ms * 1e-3 / K) / 1e9 / hbm_peak},
                    "int_pipe_equivalent": {"what": "the same launch expressed in the algorithmic W word-ops per pair of SURVEY "
                                                    "8(d) against the naive 1-POPC-per-word roof (the index never executes them)",
                                            "achieved_Twordops": word_ops / main_s / 1e12, "popc_roof_Twordops": popc_peak / 1e12,
                                            "ratio": word_ops / main_s / popc_peak},
                    "dense_case": {"what": "the bitset-streaming kernel (jaccard_kernel<TOPK,16,noskip>) on the same data: "
                                           "every word-op executed; INT-pipe bound",
                                   "pairs_per_s": ql * (hi - lo) / kd_s * world, "kernel_ms": kd_ms / K_AUX,
                                   "achieved_Twordops": word_ops / kd_s / 1e12, "frac_popc_roof": word_ops / kd_s / popc_peak,
                                   "two_pipe_roof": two_pipe / 1e12, "frac_two_pipe_roof": word_ops / kd_s / two_pipe}}

        # SURVEY.md C4 "x-like" variant: history-like sets (mean 20 ids) — denser bitsets, less zero-span skipping
        x_like = None
        if not args.no_e2e and abs(args.mean_set - 1.0 / 0.45) < 1e-6:
            xp_ids, xp_off = synth_sets(n_pool, SEED_POOL + 1, 20.0)
            xq_ids, xq_off = synth_sets(ql, SEED_QUERY + 1, 20.0)
            xi, xo = csr_rows(xp_ids, xp_off, lo, hi)
            bxp = set_encoder.encode_csr(xi, xo, V_BITS, dev)
            bxq = set_encoder.encode_csr(xq_ids, xq_off, V_BITS, dev)
            # (NCCL exchange here: the fused-exchange buffers are sized for the headline step)
            x_ms, _, _ = timed(lambda i: sharded.jaccard_topk_sharded(bxq, bxp, TOPK, pool_base=lo, workspace=ws), steps=K_AUX)
            x_like = {"value": ql * n_pool * K_AUX / (x_ms * 1e-3), "unit": "pairs/s", "ms_per_step": x_ms / K_AUX,
                      "mean_set_size": 20.0, "queries_per_step": ql}
            del bxp, bxq

        e2e = None
        if not args.no_e2e:
            q_pins = []
            for b in range(n_batches):
                qi, qo = csr_rows(q_ids, q_off, b * qs, (b + 1) * qs)
                q_pins.append((qi.pin_memory(), qo.pin_memory()))
            host_bufs = [torch.empty((qs, TOPK), dtype=torch.int32).pin_memory() for _ in range(3)]

            def fetch(r):                                                             # D2H of the step's result
                for t, h in zip(r, host_bufs):
                    h.copy_(t, non_blocking=True)
                torch.cuda.current_stream().synchronize()

            def step_e2e(i):                      # the step's input (query batch) comes from the host, the pool is state
                qi, qo = q_pins[i % n_batches]
                bq = set_encoder.encode_csr(qi, qo, V_BITS, dev)                     # H2D + encode (queries)
                fetch(sharded.jaccard_topk_sharded(bq, bp, TOPK, pool_base=lo, workspace=ws, exchange=jex))
            e_ms, _, _ = timed(step_e2e)
            h2d_q = q_pins[0][0].numel() * 4 + q_pins[0][1].numel() * 8
            h2d_p = sh_ids.numel() * 4 + sh_off.numel() * 8

            def step_e2e_cold(i):                 # variant: the pool shard's id lists are uploaded and encoded every step too
                qi, qo = q_pins[i % n_batches]
                bq = set_encoder.encode_csr(qi, qo, V_BITS, dev)
                bpool = set_encoder.encode_csr(sh_ids_pin, sh_off_pin, V_BITS, dev)   # H2D + encode (pool shard)
                fetch(sharded.jaccard_topk_sharded(bq, bpool, TOPK, pool_base=lo, workspace=ws, exchange=jex))
            ec_ms, _, _ = timed(step_e2e_cold, steps=K_AUX)
            e2e = {"value": pairs_per_step * K / (e_ms * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": int(h2d_q) * world,
                   "d2h_bytes_per_step": qs * TOPK * 12, "ms_per_step": e_ms / K,
                   "what": "r4d C-ABI through the Python host API: host CSR id lists (pinned) of the step's query batch -> "
                           "H2D -> set encoder -> fused Jaccard top-K (-> exchange + merge) -> D2H of [Q,K] (inter, union, idx) "
                           "into pinned host buffers; the pool shard's bitsets are state resident in HBM, like the pool "
                           "embeddings of the dense scorer",
                   "pool_upload_every_step": {
                       "value": pairs_per_step * K_AUX / (ec_ms * 1e-3), "unit": "pairs/s", "ms_per_step": ec_ms / K_AUX,
                       "h2d_bytes_per_step": int(h2d_q + h2d_p) * world, "d2h_bytes_per_step": qs * TOPK * 12,
                       "what": "same, but the pool shard's CSR id lists are ALSO copied from the host and re-encoded "
                               "(2.56 GB of bitsets / n_gpus rebuilt) inside every step"}}

        cpu_baseline = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            n_p_s, n_q_s = 100_000, 256            # ~14 s of single-core CPU work
            rate, dt, nq_s = cpu_port_sample(pool_ids, pool_off, q_ids, q_off, n_q_s, n_p_s, 1)
            c_rate, c_dt = cpu_c_port_sample(pool_ids, pool_off, q_ids, q_off, 512, n_p_s)
            cpu_baseline = {"value": rate, "unit": "pairs/s", "cores": 1, "kind": "port",
                            "sample": f"{nq_s} queries x {n_p_s} pool sets of the same workload ({dt:.1f} s): Python-set "
                                      f"Jaccard double loop + stable argsort top-{TOPK} (the reference is single-threaded)",
                            "c_port": {"value": c_rate, "unit": "pairs/s", "cores": 1,
                                       "sample": f"512 x {n_p_s} (sorted-list merge in C, {c_dt:.1f} s)"}}
        out = {"metric": "query-pool pairs scored+top-K/sec (Jaccard)", "value": value, "unit": "pairs/s",
               "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "u32 (bitset AND+POPC, exact integer counts)",
               "data": "synthetic", "config": dict(workload_config(args, world), exchange=exchange_used), "roofline": roofline,
               "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches * world, "clocks": clocks,
               "x_like": x_like}
        del bp, bq_all, ws
        torch.cuda.empty_cache()

    # =============================================================== dense
    if "dense" in scorers:
        n_pool, qs = args.dense_pool, args.queries_per_step
        DENSE_D = args.dense_d
        lo, hi = rank * n_pool // world, (rank + 1) * n_pool // world
        g = torch.Generator(device=dev).manual_seed(SEED_DPOOL + rank)
        hi_plane = torch.empty((hi - lo, DENSE_D), dtype=torch.bfloat16, device=dev)
        chunk = 500_000
        for a in range(0, hi - lo, chunk):
            b = min(a + chunk, hi - lo)
            x = torch.randn((b - a, DENSE_D), generator=g, device=dev)
            hi_plane[a:b] = engine.dense_prepare(x, engine.PREC_BF16).hi
            del x
        pool = engine.DensePlanes(hi_plane, None, DENSE_D, DENSE_D, engine.PREC_BF16)
        p_time = torch.rand(hi - lo, generator=g, device=dev) * 110.0
        gq = torch.Generator().manual_seed(SEED_DQUERY)
        n_batches = 4                                           # distinct query batches cycled through
        q_host = [torch.randn((qs, DENSE_D), generator=gq).pin_memory() for _ in range(n_batches)]
        qt_host = [(torch.rand(qs, generator=gq) * 110.0).pin_memory() for _ in range(n_batches)]
        q_planes = [engine.dense_prepare(q.to(dev), engine.PREC_BF16) for q in q_host]
        q_times = [t.to(dev) for t in qt_host]
        ws = torch.empty(_lib.load().r4d_dense_topk_workspace_bytes(qs, hi - lo, TOPK), dtype=torch.uint8, device=dev)
        mode = engine.DENSE_COS_DECAY
        dex = None
        if world > 1 and args.exchange == "p2p":
            try:
                dex = sharded.P2PExchange(qs, TOPK, 2)
            except Exception:
                dex = None

        def dstep(i):
            b = i % n_batches
            sharded.dense_topk_sharded(q_planes[b], pool, TOPK, pool_base=lo, mode=mode, q_time=q_times[b], p_time=p_time,
                                       lam=DENSE_LAMBDA, workspace=ws, exchange=dex)
        sampler = ClockSampler(local_rank) if rank == 0 else None
        d_ms, d_launch, d_clocks = timed(dstep, sampler, sampler_on=True)
        pairs = qs * n_pool
        d_value = pairs * K / (d_ms * 1e-3)

        def dkernel(i):
            b = i % n_batches
            engine.dense_topk(q_planes[b], pool, TOPK, mode, q_times[b], p_time, DENSE_LAMBDA, pool_base=lo, workspace=ws)
        dk_ms, _, _ = timed(dkernel)
        dk_s = dk_ms * 1e-3 / K
        flops = 2.0 * DENSE_D * qs * (hi - lo)
        peak_tf = peaks.get("bf16_tflops", 1590.0)
        d_roof = {"bound": "tensor", "achieved": flops / dk_s / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                  "frac": flops / dk_s / 1e12 / peak_tf,
                  "traffic": dram_traffic("dense_topk", queries=qs, pool=hi - lo, d=DENSE_D) if world == 1 else None,
                  "kernel": "r4d::dense2_kernel<256, streaming> (CTA pair, tcgen05.mma.cta_group::2)",
                  "kernel_ms": dk_ms / K,
                  "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst)" if peaks else "fallback 1.59 PFLOP/s"}
        d_e2e = None
        if not args.no_e2e:
            host = {}

            def dstep_e2e(i):
                b = i % n_batches
                qd = q_host[b].to(dev, non_blocking=True)                     # H2D fp32 query embeddings + times
                qt = qt_host[b].to(dev, non_blocking=True)
                r = sharded.dense_topk_sharded(engine.dense_prepare(qd, engine.PREC_BF16), pool, TOPK, pool_base=lo, mode=mode,
                                               q_time=qt, p_time=p_time, lam=DENSE_LAMBDA, workspace=ws, exchange=dex)
                host["r"] = [t.cpu() for t in r]
            de_ms, _, _ = timed(dstep_e2e)
            d_e2e = {"value": pairs * K / (de_ms * 1e-3), "unit": "pairs/s",
                     "h2d_bytes_per_step": (qs * DENSE_D * 4 + qs * 4) * world, "d2h_bytes_per_step": qs * TOPK * 8,
                     "ms_per_step": de_ms / K,
                     "what": "fp32 query embeddings + times from pinned host memory -> H2D -> normalise/bf16 -> tcgen05 "
                             "top-K -> D2H; pool embeddings stay resident in HBM (the reference keeps train_embeddings "
                             "on the GPU too, train/train_retriever.py:423,435)"}
        d_cpu = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            from oracle import dense_oracle
            torch.manual_seed(0)
            nqc, npc = 256, 200_000
            qc, pc = torch.randn(nqc, DENSE_D), torch.randn(npc, DENSE_D)
            tqc, tpc = torch.rand(nqc) * 110, torch.rand(npc) * 110
            t0 = time.perf_counter()
            sc = dense_oracle.scores(qc, pc, 1, tqc, tpc, DENSE_LAMBDA).numpy()
            dense_oracle.rank_stable(sc)
            dt = time.perf_counter() - t0
            d_cpu = {"value": nqc * npc / dt, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
                     "sample": f"{nqc} x {npc} x {DENSE_D} fp32 torch-CPU restatement of train_retriever.py:433-438 + "
                               f"decay + full stable argsort ({dt:.1f} s)"}
        dense = {"metric": "query-pool pairs scored+top-K/sec (dense)", "value": d_value, "unit": "pairs/s",
                 "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": d_ms / K, "higher_is_better": True,
                 "scaling": "strong", "dtype": "bf16 operands, fp32 accumulate (tcgen05 kind::f16)", "data": "synthetic",
                 "config": {"workload": f"synthetic dense top-K: {n_pool:,} x {DENSE_D} pool x {QUERY_N:,} queries, K={TOPK}, "
                                        f"cos*exp(-{DENSE_LAMBDA}|dt|) epilogue; step = {qs:,}-query batch vs the whole pool",
                            "parallelism": f"pool-sharded x{world}", "l2": "inputs larger than L2 (pool 15.4 GB / n_gpus)",
                            "exchange": ("fused NVLink peer stores" if dex is not None else
                                         ("nccl all-gather" if world > 1 else "none (single GPU)"))},
                 "roofline": d_roof, "cpu_baseline": d_cpu, "e2e": d_e2e, "gpu_launches": d_launch * world,
                 "clocks": d_clocks}
        if out:
            out["dense"] = dense
        else:
            out = dense

    if rank == 0:
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
