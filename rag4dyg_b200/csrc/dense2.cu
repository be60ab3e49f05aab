// dense2.cu — dense scorer, throughput kernel: CTA pairs (cta_group::2, UMMA M = 256), query tile resident in smem.
//
// Why: dense.cu streams both operands per k-block (96 B/clk/SM of L2->smem traffic at full MMA rate), which is what
// caps it near half of the tensor peak.  Here each CTA of a 2-CTA cluster keeps its 128 query rows x D resident in
// shared memory for a whole work item and streams only ITS HALF of each pool tile; the pair's single
// tcgen05.mma.cta_group::2 consumes A = 256 query rows (128 per CTA) and B = N pool rows (N/2 per CTA), so the
// streamed traffic drops to 32 B/clk/SM.
//
// Roles per CTA (320 threads): warp 0 TMA producer (both CTAs; cp.async.bulk.tensor ... .cta_group::2 signals the
// LEADER's full barrier), warp 1 MMA issuer (leader CTA only; commits multicast to both CTAs' barriers), warps 2..9
// epilogue on the CTA's own 128 TMEM lanes: per 32-column chunk the raw cosines give an upper bound on every score
// (decay <= 1), so chunks that cannot beat the row's k-th best skip the MUFU/decay work entirely; survivors go into
// per-thread sorted smem lists; the two column halves of a row share their thresholds.
//
// At D = 768 the resident tile (192 KB) would leave only 32 KB of pool stages in flight per CTA, which is TMA-latency
// bound (measured: no gain over dense.cu), so wide D uses QRES = false: the pair streams both k-blocks (64 B/clk/SM,
// 7-stage ring) and still halves the pool traffic per MMA.
//
// Supports: PREC_BF16 and PREC_BF16X3 (three products: hi.hi + hi.lo + lo.hi, streamed operands), top-K with k <= 32.
// Full score rows (r4d_dense_full) and shapes whose lists leave fewer than 3 stages go through dense.cu.
#include <cmath>
#include <cstdlib>

#include "dense_common.cuh"

namespace r4d {

constexpr int D2_THREADS = 320;
constexpr int D2_EPI_WARPS = 8;
constexpr int D2_EPI_THREADS = 256;
constexpr int D2_KMAX = R4D_TOPK_MAX;  // every top-K width of the ABI: the per-thread lists of k = 32 (64 KB) still leave 5 stages

struct Dense2Params {
    int64_t nq, np;
    int32_t n_kblocks;
    int32_t* progress;   // [n_items] tiles issued per work item (walker throttle, see the producer); nullptr = off
    int32_t window;      // tiles a walker may lead the slowest concurrent walker of its pool stripe
    int32_t n_segs;   // 1: bf16 single pass; 3: BF16X3 = q_hi.p_hi + q_hi.p_lo + q_lo.p_hi per k-block, one TMEM tile
    int32_t mode;
    int32_t k;
    int32_t n_stages;
    float neg_lambda_log2e;
    int64_t pool_base;
    const float* q_time;
    const float* p_time;
    int32_t n_qpairs, n_ptiles, n_stripes, ptiles_per_stripe;
    int32_t interleave;  // 1: stripe s owns tiles s, s+S, ...; 0: tiles [s*T, (s+1)*T)
    float* part_score;  // [2*n_stripes][nq][k]
    int32_t* part_idx;
};

// ---- cluster / 2-SM PTX wrappers
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// NOTE: deliberately without .release.cluster — that form compiles to MEMBAR.ALL.GPU + ERRBAR per call and
// serialises the TMA pipeline; the data itself is published by the TMA's complete_tx, not by this arrive.
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA tile load into this CTA's smem whose completion is signalled on a barrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst_smem, const CUtensorMap* map, uint32_t mbar_cluster_addr,
                                                int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
        "[%2];" ::"r"(dst_smem),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// commit all prior MMAs of this thread; arrive on the barrier at the same smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_2sm_mc(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}

// DPN : pool rows per pair tile (UMMA N), 256 or 128; each CTA streams DPN/2 rows per k-block.
// QRES: true  = this CTA's query tile (128 x D) stays resident in smem for the whole work item (32 B/clk/SM streamed);
//              needs D small enough to leave >= 4 pool stages (D <= 512).
//       false = the query k-block is streamed next to the pool k-block (64 B/clk/SM, deep ring) — wide D.
__device__ __forceinline__ void st_relaxed_i32(int32_t* p, int32_t v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int32_t ld_relaxed_i32(const int32_t* p) {
    int32_t v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// X3C: split precision with COMBINED stages — one stage holds the hi AND lo planes of the pool half-tile and of the query
// tile for one k-block (4 x 16 KB), and feeds all three products; the separate-stage form loads p_hi and q_hi twice
// (L2 -> SM traffic 96 KB instead of 64 KB per k-block), which costs power the tensor pipe could use.
template <int DPN, bool QRES, bool X3C>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(D2_THREADS, 1)
dense2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_p,
              const __grid_constant__ CUtensorMap tm_qlo, const __grid_constant__ CUtensorMap tm_plo,
              const Dense2Params prm) {
    constexpr int P_TILE_BYTES_ = (DPN / 2) * DKB * 2;
    static_assert(!(QRES && X3C), "combined split-precision stages stream both operands");
    // streamed Q k-block sits after the P tile; X3C: [p_hi | q_hi | p_lo | q_lo]
    constexpr int P_STAGE_BYTES = X3C ? 2 * (P_TILE_BYTES_ + Q_TILE_BYTES) : P_TILE_BYTES_ + (QRES ? 0 : Q_TILE_BYTES);
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* q_res = smem;                                              // [n_kblocks][128 x 64 bf16]  resident (QRES)
    uint8_t* stages = smem + (QRES ? (size_t)prm.n_kblocks * Q_TILE_BYTES : 0);  // [n_stages][stage]
    uint8_t* after = stages + (size_t)prm.n_stages * P_STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(after);            // [8]  used in the LEADER (count 2)
    uint64_t* empty_bar = full_bar + 8;                                 // [8]  per CTA (multicast commit)
    uint64_t* tfull_bar = empty_bar + 8;                                // [2]  per CTA (multicast commit)
    uint64_t* tempty_bar = tfull_bar + 2;                               // [2]  LEADER (count 16: both CTAs' epilogue warps)
    uint64_t* qfull_bar = tempty_bar + 2;                               // [1]  LEADER (count 2)
    uint64_t* qempty_bar = qfull_bar + 1;                               // [1]  per CTA (multicast commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(qempty_bar + 1);
    float* ptime_s = reinterpret_cast<float*>(tmem_slot + 4);           // [2][DPN]
    float* thr_s = ptime_s + 2 * DPN;                                   // [256] k-th best per (row, column half)
    float* list_s = thr_s + D2_EPI_THREADS;                             // [k][256] per-thread sorted lists
    int32_t* list_i = reinterpret_cast<int32_t*>(list_s + (size_t)prm.k * D2_EPI_THREADS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B tiles need a 1024-byte aligned base

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_p);
        if (prm.n_segs > 1) {
            tma_prefetch_desc(&tm_qlo);
            tma_prefetch_desc(&tm_plo);
        }
        for (int s = 0; s < 8; ++s) {
            mbar_init(&full_bar[s], 2);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull_bar[b], 1);
            mbar_init(&tempty_bar[b], 2 * D2_EPI_WARPS);
        }
        mbar_init(qfull_bar, 2);
        mbar_init(qempty_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc_2sm(tmem_slot, 512);
        tmem_relinquish_2sm();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_items = prm.n_qpairs * prm.n_stripes;
    const int n_clusters = gridDim.x >> 1;
    const int cluster_id = blockIdx.x >> 1;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs; lane 0 issues)
        // Walker throttle: the n_qpairs work items of one pool stripe read the same pool tiles.  Left alone they drift
        // apart (a walker that falls out of the L2 window sees DRAM latency and falls further back) and every tile is
        // fetched from HBM ~12 times (ncu, round 1/2).  Every 2 tiles a walker publishes its tile count and waits while
        // it leads the slowest walker of the same stripe and round by more than `window` tiles, so a stripe's
        // walkers share one L2-resident window and the pool crosses HBM about once.  Deadlock-free: the slowest walker
        // of the earliest unfinished round never waits, and that round's items are all running (a cluster takes its
        // items in order); the wait is bounded anyway.
        int stage = 0;
        uint32_t phase = 0, item_seq = 0;
        const uint32_t qfull_leader = mapa_shared(smem_u32(qfull_bar), 0);
        for (int item = cluster_id; item < n_items; item += n_clusters, ++item_seq) {
            const int stripe = item / prm.n_qpairs;
            const int pt_step = prm.interleave ? prm.n_stripes : 1;
            const int pt0 = prm.interleave ? stripe : stripe * prm.ptiles_per_stripe;
            const int pt_lim = prm.interleave ? prm.n_ptiles : min(pt0 + prm.ptiles_per_stripe, prm.n_ptiles);
            const int qtile = (item - stripe * prm.n_qpairs) * 2 + (int)rank;
            if (QRES && lane == 0) {
                // resident query tile: wait until the previous item's MMAs have released it
                mbar_wait(qempty_bar, (item_seq & 1u) ^ 1u);
                mbar_arrive_expect_tx_cluster(qfull_leader, (uint32_t)(prm.n_kblocks * Q_TILE_BYTES));
                for (int kb = 0; kb < prm.n_kblocks; ++kb)
                    tma_load_2d_2sm(smem_u32(q_res + (size_t)kb * Q_TILE_BYTES), &tm_q, qfull_leader, kb * DKB,
                                    qtile * DQ);
            }
            const int round_first = (item / n_clusters) * n_clusters;      // items running concurrently with this one
            const int peer0 = stripe * prm.n_qpairs;
            int tiles_done = 0;
            for (int pt = pt0; pt < pt_lim; pt += pt_step, ++tiles_done) {
                if (prm.progress != nullptr && (tiles_done & 1) == 0) {
                    if (lane == 0 && leader) st_relaxed_i32(prm.progress + item, tiles_done);
                    for (int spin = 0; spin < 20000; ++spin) {
                        int slowest = 0x7fffffff;
                        for (int pl = lane; pl < prm.n_qpairs; pl += 32) {
                            const int it = peer0 + pl;
                            if (it >= round_first && it < round_first + n_clusters && it < n_items)
                                slowest = min(slowest, ld_relaxed_i32(prm.progress + it));
                        }
                        slowest = __reduce_min_sync(0xffffffffu, slowest);
                        if (tiles_done - slowest <= prm.window) break;
                        __nanosleep(256);
                    }
                }
                if (lane == 0 && X3C) {
                    for (int kb = 0; kb < prm.n_kblocks; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[stage]), 0);
                        const uint32_t sdst = smem_u32(stages + (size_t)stage * P_STAGE_BYTES);
                        mbar_arrive_expect_tx_cluster(full_leader, (uint32_t)P_STAGE_BYTES);
                        const int prow = pt * DPN + (int)rank * (DPN / 2);
                        tma_load_2d_2sm(sdst, &tm_p, full_leader, kb * DKB, prow);
                        tma_load_2d_2sm(sdst + P_TILE_BYTES_, &tm_q, full_leader, kb * DKB, qtile * DQ);
                        tma_load_2d_2sm(sdst + P_TILE_BYTES_ + Q_TILE_BYTES, &tm_plo, full_leader, kb * DKB, prow);
                        tma_load_2d_2sm(sdst + 2 * P_TILE_BYTES_ + Q_TILE_BYTES, &tm_qlo, full_leader, kb * DKB, qtile * DQ);
                        if (++stage == prm.n_stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
                if (lane == 0 && !X3C) {
                    // split precision: per k-block the products q_hi.p_hi, q_hi.p_lo, q_lo.p_hi in this order — the
                    // same sequence of K = 16 MMA steps as dense.cu issues, so both kernels accumulate identically
                    for (int kb = 0; kb < prm.n_kblocks; ++kb) {
                        for (int seg = 0; seg < prm.n_segs; ++seg) {
                            const CUtensorMap* mp = seg == 1 ? &tm_plo : &tm_p;
                            const CUtensorMap* mq = seg == 2 ? &tm_qlo : &tm_q;
                            mbar_wait(&empty_bar[stage], phase ^ 1);
                            const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[stage]), 0);
                            const uint32_t sdst = smem_u32(stages + (size_t)stage * P_STAGE_BYTES);
                            mbar_arrive_expect_tx_cluster(full_leader, (uint32_t)P_STAGE_BYTES);
                            tma_load_2d_2sm(sdst, mp, full_leader, kb * DKB, pt * DPN + (int)rank * (DPN / 2));
                            if (!QRES) tma_load_2d_2sm(sdst + P_TILE_BYTES_, mq, full_leader, kb * DKB, qtile * DQ);
                            if (++stage == prm.n_stages) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                    }
                }
                __syncwarp();
            }
            if (prm.progress != nullptr && lane == 0 && leader) st_relaxed_i32(prm.progress + item, 0x7fffffff);   // done
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA, one lane)
        if (leader && lane == 0) {
            // D=f32, A=B=bf16, K-major; N>>3 at bit 17; M = 256 (pair) -> M>>4 at bit 24
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(DPN >> 3) << 17) |
                                   ((uint32_t)(256 >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0, tile_seq = 0, item_seq = 0;
            for (int item = cluster_id; item < n_items; item += n_clusters, ++item_seq) {
                const int stripe = item / prm.n_qpairs;
                const int pt_step = prm.interleave ? prm.n_stripes : 1;
                const int pt0 = prm.interleave ? stripe : stripe * prm.ptiles_per_stripe;
                const int pt_lim = prm.interleave ? prm.n_ptiles : min(pt0 + prm.ptiles_per_stripe, prm.n_ptiles);
                if (QRES) {
                    mbar_wait(qfull_bar, item_seq & 1u);  // both CTAs' query tiles have landed
                    tc_fence_after();
                }
                const uint32_t qbase = smem_u32(q_res);
                for (int pt = pt0; pt < pt_lim; pt += pt_step, ++tile_seq) {
                    const uint32_t buf = tile_seq & 1u;
                    mbar_wait(&tempty_bar[buf], ((tile_seq >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + buf * DPN;
                    const int n_vk = X3C ? prm.n_kblocks : prm.n_kblocks * prm.n_segs;   // all segments accumulate into one tile
                    for (int kb = 0; kb < n_vk; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t sbase = smem_u32(stages + (size_t)stage * P_STAGE_BYTES);
                        const uint64_t qd = make_smem_desc(QRES ? qbase + (uint32_t)kb * Q_TILE_BYTES : sbase + P_TILE_BYTES_);
                        const uint64_t pd = make_smem_desc(sbase);
#pragma unroll
                        for (int ks = 0; ks < DKB / 16; ++ks)
                            umma_bf16_2sm(tmem_d, qd + (uint64_t)(2 * ks), pd + (uint64_t)(2 * ks), idesc,
                                          (kb | ks) != 0 ? 1u : 0u);
                        if (X3C) {   // + q_hi.p_lo + q_lo.p_hi from the same stage (same order as dense.cu)
                            const uint64_t pl = make_smem_desc(sbase + P_TILE_BYTES_ + Q_TILE_BYTES);
                            const uint64_t ql = make_smem_desc(sbase + 2 * P_TILE_BYTES_ + Q_TILE_BYTES);
#pragma unroll
                            for (int ks = 0; ks < DKB / 16; ++ks)
                                umma_bf16_2sm(tmem_d, qd + (uint64_t)(2 * ks), pl + (uint64_t)(2 * ks), idesc, 1u);
#pragma unroll
                            for (int ks = 0; ks < DKB / 16; ++ks)
                                umma_bf16_2sm(tmem_d, ql + (uint64_t)(2 * ks), pd + (uint64_t)(2 * ks), idesc, 1u);
                        }
                        tc_commit_2sm_mc(&empty_bar[stage]);  // pool stage free in both CTAs
                        if (++stage == prm.n_stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    tc_commit_2sm_mc(&tfull_bar[buf]);  // accumulator complete in both CTAs
                }
                if (QRES) tc_commit_2sm_mc(qempty_bar);  // resident query tiles may be overwritten
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue warps 2..9 (own 128 TMEM lanes)
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int half = ew >> 2;
        const int etid = threadIdx.x - 64;
        const int row_in_tile = quarter * 32 + lane;
        constexpr int COLS_PER_THREAD = DPN / 2;
        uint32_t tile_seq = 0;
        uint32_t tempty_leader[2];
        tempty_leader[0] = mapa_shared(smem_u32(&tempty_bar[0]), 0);
        tempty_leader[1] = mapa_shared(smem_u32(&tempty_bar[1]), 0);

        for (int item = cluster_id; item < n_items; item += n_clusters) {
            const int stripe = item / prm.n_qpairs;
                const int pt_step = prm.interleave ? prm.n_stripes : 1;
                const int pt0 = prm.interleave ? stripe : stripe * prm.ptiles_per_stripe;
                const int pt_lim = prm.interleave ? prm.n_ptiles : min(pt0 + prm.ptiles_per_stripe, prm.n_ptiles);
            const int qtile = (item - stripe * prm.n_qpairs) * 2 + (int)rank;
            const int64_t gq = (int64_t)qtile * DQ + row_in_tile;
            const bool q_ok = gq < prm.nq;
            const float tq = (prm.mode != R4D_DENSE_HALF_COS && q_ok) ? prm.q_time[gq] : 0.f;
            float* ls = list_s + (half * 128 + row_in_tile);
            int32_t* li = list_i + (half * 128 + row_in_tile);
            for (int t = 0; t < prm.k; ++t) {
                ls[t * D2_EPI_THREADS] = -INFINITY;
                li[t * D2_EPI_THREADS] = R4D_IDX_NONE;
            }
            float thr = -INFINITY;  // this thread's k-th best
            // the other column half of the same query row shares its threshold (stored one ulp low so that the strict
            // test below keeps equal scores, whose index order across halves is not monotone)
            thr_s[half * 128 + row_in_tile] = -INFINITY;
            named_bar_sync(3, D2_EPI_THREADS);
            const float* thr_partner = thr_s + ((half ^ 1) * 128 + row_in_tile);
            // decay in (0, 1] lets the raw cosine bound every score of a chunk from above (lambda >= 0 only)
            const bool can_bound = prm.neg_lambda_log2e <= 0.f;
            for (int pt = pt0; pt < pt_lim; pt += pt_step, ++tile_seq) {
                const uint32_t buf = tile_seq & 1u;
                if (prm.mode != R4D_DENSE_HALF_COS) {
                    if (etid < DPN) {
                        const int64_t gp = (int64_t)pt * DPN + etid;
                        ptime_s[buf * DPN + etid] = gp < prm.np ? prm.p_time[gp] : 0.f;
                    }
                    named_bar_sync(2, D2_EPI_THREADS);
                }
                float cut = fmaxf(thr, *thr_partner);  // candidates must beat both halves' k-th best
                mbar_wait(&tfull_bar[buf], (tile_seq >> 1) & 1u);
                tc_fence_after();
#pragma unroll 1
                for (int ch = 0; ch < COLS_PER_THREAD / 32; ++ch) {
                    const int col0 = half * COLS_PER_THREAD + ch * 32;
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * DPN + col0, v);
                    tmem_ld_wait();
                    const int64_t gp0 = (int64_t)pt * DPN + col0;
                    // fast reject: upper bound of the chunk's scores from the raw cosines (no MUFU, no pool times)
                    float xm = __uint_as_float(v[0]);
#pragma unroll
                    for (int j = 1; j < 32; ++j) xm = fmaxf(xm, __uint_as_float(v[j]));
                    const float ub = prm.mode == R4D_DENSE_COS_DECAY ? fmaxf(xm, 0.f) : fmaxf((xm + 1.0f) * 0.5f, 0.f);
                    if (!q_ok || (can_bound && !(ub > cut))) continue;
                    float s[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float tp = prm.mode == R4D_DENSE_HALF_COS ? 0.f : ptime_s[buf * DPN + col0 + j];
                        s[j] = dense_score(__uint_as_float(v[j]), prm.mode, tq, tp, prm.neg_lambda_log2e);
                    }
                    if (gp0 + 32 > prm.np) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (gp0 + j >= prm.np) s[j] = -INFINITY;
                    }
                    float m = s[0];
#pragma unroll
                    for (int j = 1; j < 32; ++j) m = fmaxf(m, s[j]);
                    if (m > cut) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (s[j] > cut) {
                                thr = list_insert(ls, li, prm.k, s[j], (int32_t)(prm.pool_base + gp0 + j));
                                cut = fmaxf(cut, thr);
                            }
                    }
                }
                thr_s[half * 128 + row_in_tile] = nextafterf(thr, -INFINITY);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tempty_leader[buf]);
            }
            if (q_ok) {
                const int64_t o = ((int64_t)(stripe * 2 + half) * prm.nq + gq) * prm.k;
                for (int t = 0; t < prm.k; ++t) {
                    prm.part_score[o + t] = ls[t * D2_EPI_THREADS];
                    prm.part_idx[o + t] = li[t * D2_EPI_THREADS];
                }
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();  // no CTA may exit while its peer can still signal its barriers / read its smem
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------- host side
struct Dense2Plan {
    int32_t n_qpairs, n_ptiles, n_stripes, ptiles_per_stripe, dpn, n_stages;
    bool qres;
    bool x3c;   // split precision with combined 64 KB stages
    bool ok;
    size_t smem;
};

static Dense2Plan dense2_plan(int64_t nq, int64_t np, int32_t d_pad, int32_t k, bool x3 = false) {
    Dense2Plan pl{};
    pl.ok = false;
    if (k > D2_KMAX || nq <= 0 || np <= 0) return pl;
    const int n_kblocks = d_pad / DKB;
    const long total = 227 * 1024;
    const long fixed = 22 * 8 + 16 + 2 * 256 * 4 + 64 + 256 * 4 + (long)k * 256 * 8;  // barriers, times, thr, lists
    pl.dpn = 256;
    const long q_bytes = (long)n_kblocks * Q_TILE_BYTES;
    const int force = options().dense_pair_qres;  // -1 auto, 0 never resident
    const long avail_res = total - fixed - q_bytes;
    pl.qres = avail_res >= 4 * 16384 && force != 0 && !x3;   // split precision streams hi and lo planes of both operands
    long stage_bytes = 16384 + (pl.qres ? 0 : Q_TILE_BYTES);
    pl.x3c = x3 && options().dense_x3_combined && (total - fixed) / (2 * (16384 + Q_TILE_BYTES)) >= 3;
    if (pl.x3c) stage_bytes = 2 * (16384 + Q_TILE_BYTES);
    pl.n_stages = (int)((pl.qres ? avail_res : total - fixed) / stage_bytes);
    if (pl.n_stages > 8) pl.n_stages = 8;
    if (pl.n_stages < 3) return pl;
    pl.smem = (size_t)fixed + (size_t)pl.n_stages * (size_t)stage_bytes + (pl.qres ? (size_t)q_bytes : 0);
    pl.n_qpairs = (int32_t)((nq + 2 * DQ - 1) / (2 * DQ));
    pl.n_ptiles = (int32_t)((np + pl.dpn - 1) / pl.dpn);
    // Stripe count by a small cost model: rounds(items / clusters) x (tile time + top-K warm-up per stripe).  A stripe
    // of L columns costs ~24 cycles per column (MMA bound) plus ~32*k*(1 + ln(L / (32*k))) list insertions per warp.
    const int n_clusters = num_sms() / 2;
    const int64_t max_stripes = pl.n_ptiles < 4096 ? pl.n_ptiles : 4096;
    double best_cost = 0;
    int64_t best = 1;
    for (int64_t st = 1; st <= max_stripes; ++st) {
        const int64_t tiles = (pl.n_ptiles + st - 1) / st;
        const int64_t real_st = (pl.n_ptiles + tiles - 1) / tiles;
        if (real_st != st) continue;
        const double L = (double)tiles * pl.dpn;
        const double ev = 32.0 * k * (1.0 + (L > 32.0 * k ? log(L / (32.0 * k)) : 0.0));
        const double per_item = 24.0 * L + 150.0 * ev + 3000.0;
        const int64_t items = (int64_t)pl.n_qpairs * st;
        const int64_t rounds = (items + n_clusters - 1) / n_clusters;
        const double cost = (double)rounds * per_item;
        if (st == 1 || cost < best_cost) {
            best_cost = cost;
            best = st;
        }
    }
    if (options().dense_stripes > 0) best = options().dense_stripes < pl.n_ptiles ? options().dense_stripes : pl.n_ptiles;
    pl.ptiles_per_stripe = (int32_t)((pl.n_ptiles + best - 1) / best);
    pl.n_stripes = (pl.n_ptiles + pl.ptiles_per_stripe - 1) / pl.ptiles_per_stripe;
    pl.ok = true;
    return pl;
}

// exported to dense.cu
bool dense2_supported(int64_t nq, int64_t np, int32_t d_pad, int32_t prec, int32_t k) {
    if (!options().dense_pair_kernel || (prec != R4D_PREC_BF16 && prec != R4D_PREC_BF16X3)) return false;
    return dense2_plan(nq, np, d_pad, k, prec == R4D_PREC_BF16X3).ok;
}

// workspace: part_score [n_lists][nq][k] f32 | part_idx [n_lists][nq][k] i32 | progress [n_items] i32 (256-byte aligned)
static size_t dense2_lists_bytes(const Dense2Plan& pl, int64_t nq, int32_t k) {
    return (((size_t)pl.n_stripes * 2 * (size_t)nq * (size_t)k * 4) + 255) / 256 * 256;
}
size_t dense2_workspace_bytes(int64_t nq, int64_t np, int32_t d_pad, int32_t k, bool x3) {
    const Dense2Plan pl = dense2_plan(nq, np, d_pad, k, x3);
    return 2 * dense2_lists_bytes(pl, nq, k) + ((size_t)pl.n_qpairs * pl.n_stripes * 4 + 255) / 256 * 256 + 256;
}

int dense2_topk(const void* q_hi, const void* q_lo, int64_t nq, const void* p_hi, const void* p_lo, int64_t np, int32_t d_pad,
                const float* q_time, const float* p_time, float lambda, int32_t mode, int32_t k, int64_t pool_base,
                void* workspace, float** part_score_out, int32_t** part_idx_out, int32_t* n_lists_out, cudaStream_t st) {
    const bool x3 = q_lo != nullptr && p_lo != nullptr;
    const Dense2Plan pl = dense2_plan(nq, np, d_pad, k, x3);
    if (!pl.ok) {
        set_error("dense2: unsupported shape");
        return R4D_E_ARG;
    }
    const size_t lists = dense2_lists_bytes(pl, nq, k);
    float* part_score = reinterpret_cast<float*>(workspace);
    int32_t* part_idx = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(workspace) + lists);
    int32_t* progress = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(workspace) + 2 * lists);
    *part_score_out = part_score;
    *part_idx_out = part_idx;
    CUtensorMap tm_q, tm_p;
    int rc;
    const uint64_t rs = (uint64_t)d_pad * 2;
    if ((rc = make_tmap_2d(&tm_q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, q_hi, d_pad, nq, rs, DKB, DQ,
                           CU_TENSOR_MAP_SWIZZLE_128B)))
        return rc;
    if ((rc = make_tmap_2d(&tm_p, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p_hi, d_pad, np, rs, DKB, pl.dpn / 2,
                           CU_TENSOR_MAP_SWIZZLE_128B)))
        return rc;
    CUtensorMap tm_qlo = tm_q, tm_plo = tm_p;
    if (x3) {
        if ((rc = make_tmap_2d(&tm_qlo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, q_lo, d_pad, nq, rs, DKB, DQ,
                               CU_TENSOR_MAP_SWIZZLE_128B)))
            return rc;
        if ((rc = make_tmap_2d(&tm_plo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p_lo, d_pad, np, rs, DKB, pl.dpn / 2,
                               CU_TENSOR_MAP_SWIZZLE_128B)))
            return rc;
    }
    Dense2Params prm{};
    prm.nq = nq;
    prm.np = np;
    prm.n_kblocks = d_pad / DKB;
    prm.n_segs = x3 ? 3 : 1;
    prm.mode = mode;
    prm.k = k;
    prm.n_stages = pl.n_stages;
    prm.neg_lambda_log2e = -lambda * 1.4426950408889634f;
    prm.pool_base = pool_base;
    prm.q_time = q_time;
    prm.p_time = p_time;
    prm.n_qpairs = pl.n_qpairs;
    prm.n_ptiles = pl.n_ptiles;
    prm.n_stripes = pl.n_stripes;
    prm.ptiles_per_stripe = pl.ptiles_per_stripe;
    prm.interleave = options().stripe_interleave ? 1 : 0;
    prm.part_score = part_score;
    prm.part_idx = part_idx;
    const size_t smem = pl.smem;
    const int64_t n_items = (int64_t)pl.n_qpairs * pl.n_stripes;
    // walker throttle (option "dense_walker_window": tiles; 0 = off): only useful when several walkers share a stripe
    prm.window = options().dense_walker_window;
    prm.progress = (prm.window > 0 && pl.n_qpairs > 1) ? progress : nullptr;
    if (prm.progress) R4D_CUDA(cudaMemsetAsync(progress, 0, (size_t)n_items * 4, st));
    int n_clusters = num_sms() / 2;
    if (n_items < n_clusters) n_clusters = (int)n_items;
    if (pl.qres) {
        R4D_CUDA(cudaFuncSetAttribute(dense2_kernel<256, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prof_begin(PROF_DENSE_PAIR, st);
        dense2_kernel<256, true, false><<<2 * n_clusters, D2_THREADS, smem, st>>>(tm_q, tm_p, tm_qlo, tm_plo, prm); note_launch();
    } else if (pl.x3c) {
        R4D_CUDA(cudaFuncSetAttribute(dense2_kernel<256, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prof_begin(PROF_DENSE_PAIR, st);
        dense2_kernel<256, false, true><<<2 * n_clusters, D2_THREADS, smem, st>>>(tm_q, tm_p, tm_qlo, tm_plo, prm); note_launch();
    } else {
        R4D_CUDA(cudaFuncSetAttribute(dense2_kernel<256, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prof_begin(PROF_DENSE_PAIR, st);
        dense2_kernel<256, false, false><<<2 * n_clusters, D2_THREADS, smem, st>>>(tm_q, tm_p, tm_qlo, tm_plo, prm); note_launch();
    }
    prof_end(PROF_DENSE_PAIR, st);
    R4D_CUDA(cudaGetLastError());
    *n_lists_out = pl.n_stripes * 2;
    return R4D_OK;
}

}  // namespace r4d
