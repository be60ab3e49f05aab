// jaccard.cu — subsystems 2 + 4: all-pairs Jaccard on bitsets, full-matrix mode and fused top-K mode.
//
// Replaces occurrence_matrix / co_occurrence_ratio (retrieval_data_annotation.py:36-41, :5-15) and, in
// top-K mode, also np.argsort(-row)[:k] (retrieval_data_annotation.py:97-103).
//
// Design (B200, sm_100a)
//  * CTA tile 128 queries x 128 pool rows, persistent CTAs (one per SM) walking a static list of
//    (pool stripe, query tile) work items.
//  * TMA: per 32-word chunk two cp.async.bulk.tensor.2d loads (128 rows x 128 B, SWIZZLE_128B) land in a 3-stage
//    smem ring guarded by full/empty mbarriers; one elected consumer lane issues them two stages ahead.
//  * 8 or 16 consumer warps: each thread owns an 8x8 (or 4x8) micro-tile of intersection counters; per pair of
//    16-byte groups it issues conflict-free LDS.128 (the 128B swizzle spreads the 8 rows a quarter-warp touches
//    over all banks), ANDs 8 words, compresses 7 of them with a LOP3 carry-save tree and issues 4 POPC
//    (the POPC/XU pipe, 16 lanes/clk/SM, is the scarce unit; LOP3 runs at 64 lanes/clk/SM — DESIGN.md).
//  * counts are exact integers; union = |q| + |p| - inter.  Scores are compared as rationals by 64-bit
//    cross-multiplication (JEntry::better), so ranking is exact and ties break by ascending pool index.
//  * top-K mode: the 128x128 count tile goes through smem once, each consumer warp scans 16 query rows,
//    ballots the entries that beat the row's current k-th candidate and inserts them into a one-entry-per-lane
//    sorted list (WarpTopK).  Per-stripe lists are merged by jaccard_merge_kernel.
#include <cstdlib>

#include "jaccard_common.cuh"

namespace r4d {

constexpr int TQ = 128;
constexpr int TP = 128;
constexpr int CHUNK_WORDS = 32;                   // 128 B of each row per pipeline stage
constexpr int OPER_BYTES = TQ * CHUNK_WORDS * 4;  // 16 KB per operand per stage
constexpr int STAGE_BYTES = 2 * OPER_BYTES;
constexpr int NSTAGES = 3;
// consumer warps per CTA: 8 (8x8 counters per thread) or 16 (4x8 counters per thread, 4 warps per scheduler)
constexpr int TILE_LD = TP + 8;  // padded pitch (words) of the count tile: conflict-free register->smem spill
constexpr int LIST_LD = 32;      // one list slot per lane

constexpr int MODE_TOPK = 0;
constexpr int MODE_FULL = 1;

struct JaccardParams {
    const uint32_t* qcard;
    const uint32_t* pcard;
    int64_t nq, np;
    int32_t n_chunks;     // ceil(words / 32)
    int32_t last_groups;  // 16-byte groups holding real words in the last chunk (1..8)
    int32_t k;
    int32_t zero_diag;
    const uint32_t* tile_filter;  // non-null: only query tiles with tile_filter[qtile] != 0 (the rest ran sparse)
    const uint32_t* any_filtered; // non-null: *any_filtered == 0 means no tile passes the filter (the launch is a no-op)
    int64_t query_base, pool_base;
    int32_t n_qtiles, n_ptiles, n_stripes, ptiles_per_stripe;
    // top-K partial lists [n_stripes][nq][k]
    uint4* part;   // {inter, union, idx, 0}
    // full-matrix outputs
    uint32_t* inter;
    int64_t ld_inter;
    double* score;
    int64_t ld_score;
};


__device__ __forceinline__ void lds128(uint4& v, uint32_t addr) {
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
}
// carry-save adder on 32 independent bit columns: sum = a^b^c, carry = majority(a,b,c) — one LOP3 each
__device__ __forceinline__ uint32_t csa_sum(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t csa_carry(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xe8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
// |q & p| over 8 words with 4 POPC instead of 8: the POPC unit (16 lanes/clk/SM) is the saturated pipe, LOP3 runs
// at 64 lanes/clk/SM, so 7 of the 8 AND words are compressed by a carry-save tree (ones, twos, fours) first.
//   8 AND + 8 LOP3(CSA) + 4 POPC + 1 IADD3 + 2 IMAD   vs   8 AND + 8 POPC + 4 IADD3
__device__ __forceinline__ uint32_t popc8_csa(const uint4& qa, const uint4& qb, const uint4& pa, const uint4& pb) {
    const uint32_t a0 = qa.x & pa.x, a1 = qa.y & pa.y, a2 = qa.z & pa.z, a3 = qa.w & pa.w;
    const uint32_t a4 = qb.x & pb.x, a5 = qb.y & pb.y, a6 = qb.z & pb.z, a7 = qb.w & pb.w;
    const uint32_t s1 = csa_sum(a0, a1, a2), c1 = csa_carry(a0, a1, a2);
    const uint32_t s2 = csa_sum(s1, a3, a4), c2 = csa_carry(s1, a3, a4);
    const uint32_t s3 = csa_sum(s2, a5, a6), c3 = csa_carry(s2, a5, a6);
    const uint32_t t = csa_sum(c1, c2, c3), f = csa_carry(c1, c2, c3);
    return __popc(s3) + __popc(a7) + 2u * __popc(t) + 4u * __popc(f);
}

constexpr size_t smem_bytes_for(int mode) {
    size_t b = 1024 /*alignment slack*/ + (size_t)NSTAGES * STAGE_BYTES + 2 * NSTAGES * sizeof(uint64_t);
    if (mode == MODE_TOPK) b += (size_t)TQ * TILE_LD * 4 + 3 * (size_t)TQ * LIST_LD * 4;
    return b;
}

// No dedicated producer warp: registers are split per SM sub-partition (16 K each), so a 9th / 17th warp would put a
// third / fifth warp on one scheduler and cut the per-thread budget to 168 / 96.  With exactly 8 or 16 warps the
// budget is 255 / 128; lane 0 of warp 0 issues the TMA refills inline, two stages ahead of consumption.
// SKIP: bitsets of real node-id sets are very sparse (2.2 of 20 000 bits set in the headline workload).  An 8-word
// span whose query words are zero in every lane of the warp (one REDUX.OR), or whose pool words are zero in every lane
// (one VOTE per pool row), contributes nothing, so its LDS / LOP3 / POPC work is skipped — results are identical.
template <int MODE, int NCW, bool SKIP>
__global__ void __launch_bounds__(32 * NCW, 1)
jaccard_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_p,
               const JaccardParams prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stages = smem;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)NSTAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + NSTAGES;
    uint32_t* tilebuf = reinterpret_cast<uint32_t*>(empty_bar + NSTAGES);  // [TQ][TILE_LD]      (top-K mode)
    uint32_t* l_inter = tilebuf + TQ * TILE_LD;                           // [TQ][LIST_LD] x 3  (top-K mode)
    uint32_t* l_union = l_inter + TQ * LIST_LD;
    int32_t* l_idx = reinterpret_cast<int32_t*>(l_union + TQ * LIST_LD);

    constexpr int N_CONSUMER_WARPS = NCW;
    constexpr int N_CONSUMERS = 32 * NCW;
    constexpr int RQ = 64 / NCW;          // query rows per thread: 8 or 4
    constexpr int SLAB = 4 * RQ;          // query rows per warp slab: 32 or 16
    constexpr int SCAN_ROWS = TQ / NCW;   // query rows each warp scans in the top-K epilogue: 16 or 8
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (prm.any_filtered != nullptr && *prm.any_filtered == 0u) return;  // every tile was served by the query-index kernel

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_p);
        for (int s = 0; s < NSTAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], N_CONSUMER_WARPS);
        }
        fence_barrier_init();
    }
    __syncthreads();

    const int n_items = prm.n_qtiles * prm.n_stripes;

    // ---------------------------------------------------------------- inline TMA producer (thread 0 only)
    // The loads of this CTA form one linear stream of chunks (item -> pool tile -> 32-word chunk); the cursor below
    // walks it NSTAGES-1 chunks ahead of the consumers.
    const bool is_producer = threadIdx.x == 0;
    int p_item = blockIdx.x, p_qtile = 0, p_pt = 0, p_pt_end = 0, p_c = 0, p_stage = 0;
    uint32_t p_phase = 0;
    auto item_mine = [&](int it) { return prm.tile_filter == nullptr || prm.tile_filter[it % prm.n_qtiles] != 0u; };
    auto p_open_item = [&]() {
        while (p_item < n_items && !item_mine(p_item)) p_item += gridDim.x;
        if (p_item < n_items) {
            const int stripe = p_item / prm.n_qtiles;
            p_qtile = p_item - stripe * prm.n_qtiles;
            p_pt = stripe * prm.ptiles_per_stripe;
            p_pt_end = min(p_pt + prm.ptiles_per_stripe, prm.n_ptiles);
            p_c = 0;
        }
    };
    auto p_issue = [&]() {
        if (p_item >= n_items) return;
        mbar_wait(&empty_bar[p_stage], p_phase ^ 1);
        uint8_t* dst = stages + (size_t)p_stage * STAGE_BYTES;
        mbar_arrive_expect_tx(&full_bar[p_stage], STAGE_BYTES);
        tma_load_2d(dst, &tm_q, &full_bar[p_stage], p_c * CHUNK_WORDS, p_qtile * TQ);
        tma_load_2d(dst + OPER_BYTES, &tm_p, &full_bar[p_stage], p_c * CHUNK_WORDS, p_pt * TP);
        if (++p_stage == NSTAGES) {
            p_stage = 0;
            p_phase ^= 1;
        }
        if (++p_c == prm.n_chunks) {
            p_c = 0;
            if (++p_pt == p_pt_end) {
                p_item += gridDim.x;
                p_open_item();
            }
        }
    };
    if (is_producer) {
        p_open_item();
        for (int i = 0; i < NSTAGES - 1; ++i) p_issue();
    }

    // ---------------------------------------------------------------- consumers (all NCW warps)
    const int cw = warp;        // 0..NCW-1
    const int wq = cw >> 1;     // query slab of SLAB rows
    const int wp = cw & 1;      // 0..1 : 64-pool-row slab
    const int lq = lane >> 3;   // 0..3
    const int lp = lane & 7;    // 0..7
    // rows owned by this thread: q(i) = wq*SLAB + i*4 + lq (i < RQ), p(j) = wp*64 + j*8 + lp (j < 8)
    const uint32_t q_off = (uint32_t)(wq * SLAB + lq) * 128u;           // + i*512
    const uint32_t p_off = OPER_BYTES + (uint32_t)(wp * 64 + lp) * 128u;  // + j*1024
    const uint32_t q_xor_even = (uint32_t)lq << 4;                      // (row & 7) << 4 for even i
    const uint32_t q_xor_odd = (uint32_t)(lq + 4) << 4;                 //                    odd i
    const uint32_t p_xor = (uint32_t)lp << 4;
    const uint32_t stages_u32 = smem_u32(stages);

    int stage = 0;
    uint32_t phase = 0;

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        if (!item_mine(item)) continue;  // block-uniform: this query tile was handled by the sparse-query kernel
        const int stripe = item / prm.n_qtiles;
        const int qtile = item - stripe * prm.n_qtiles;
        const int pt_beg = stripe * prm.ptiles_per_stripe;
        const int pt_end = min(pt_beg + prm.ptiles_per_stripe, prm.n_ptiles);

        if (MODE == MODE_TOPK) {
            // reset the lists this warp owns
            for (int qq = 0; qq < SCAN_ROWS; ++qq) {
                const int q = cw * SCAN_ROWS + qq;
                l_inter[q * LIST_LD + lane] = 0u;
                l_union[q * LIST_LD + lane] = 1u;
                l_idx[q * LIST_LD + lane] = R4D_IDX_NONE;
            }
            __syncwarp();
        }

        for (int pt = pt_beg; pt < pt_end; ++pt) {
            uint32_t acc[RQ][8];
#pragma unroll
            for (int i = 0; i < RQ; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = 0u;

            for (int c = 0; c < prm.n_chunks; ++c) {
                if (is_producer) p_issue();  // refill the stage released one chunk ago
                mbar_wait(&full_bar[stage], phase);
                const uint32_t sbase = stages_u32 + (uint32_t)stage * STAGE_BYTES;
                // 16-byte groups are consumed in pairs (8 words per counter update).  Words past W are zero in
                // both operands (encoder padding / TMA zero fill), so rounding the group count up to even is exact.
                const int npairs = ((c == prm.n_chunks - 1) ? prm.last_groups + 1 : 8) >> 1;
#pragma unroll 1
                for (int g2 = 0; g2 < npairs; ++g2) {
                    const uint32_t g = (uint32_t)g2 << 5;  // byte offset of the first group of the pair (2 x 16 B)
                    const uint32_t gq_e0 = sbase + q_off + (g ^ q_xor_even);
                    const uint32_t gq_e1 = sbase + q_off + ((g + 16u) ^ q_xor_even);
                    const uint32_t gq_o0 = sbase + q_off + (g ^ q_xor_odd);
                    const uint32_t gq_o1 = sbase + q_off + ((g + 16u) ^ q_xor_odd);
                    const uint32_t gp0 = sbase + p_off + (g ^ p_xor);
                    const uint32_t gp1 = sbase + p_off + ((g + 16u) ^ p_xor);
                    uint4 qa[RQ], qb[RQ];
#pragma unroll
                    for (int i = 0; i < RQ; ++i) {
                        lds128(qa[i], ((i & 1) ? gq_o0 : gq_e0) + (uint32_t)i * 512u);
                        lds128(qb[i], ((i & 1) ? gq_o1 : gq_e1) + (uint32_t)i * 512u);
                    }
                    uint32_t qmask = (1u << RQ) - 1u;  // bit i: some lane holds a non-zero query word for row slot i
                    if (SKIP) {
                        uint32_t mine = 0;
#pragma unroll
                        for (int i = 0; i < RQ; ++i) {
                            const uint32_t nz = (qa[i].x | qa[i].y | qa[i].z) | (qa[i].w | qb[i].x | qb[i].y) | (qb[i].z | qb[i].w);
                            mine |= (nz != 0u ? 1u : 0u) << i;
                        }
                        qmask = __reduce_or_sync(0xffffffffu, mine);
                        if (qmask == 0u) continue;  // warp-uniform: nothing to intersect in this span
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        uint4 pa, pb;
                        lds128(pa, gp0 + (uint32_t)j * 1024u);
                        lds128(pb, gp1 + (uint32_t)j * 1024u);
                        if (SKIP) {
                            const uint32_t pnz = (pa.x | pa.y | pa.z) | (pa.w | pb.x | pb.y) | (pb.z | pb.w);
                            if (!__any_sync(0xffffffffu, pnz != 0u)) continue;
#pragma unroll
                            for (int i = 0; i < RQ; ++i)
                                if (qmask & (1u << i)) acc[i][j] += popc8_csa(qa[i], qb[i], pa, pb);
                        } else {
#pragma unroll
                            for (int i = 0; i < RQ; ++i) acc[i][j] += popc8_csa(qa[i], qb[i], pa, pb);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[stage]);
                if (++stage == NSTAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }

            // -------------------------------------------------------- epilogue
            if (MODE == MODE_FULL) {
#pragma unroll
                for (int i = 0; i < RQ; ++i) {
                    const int64_t gq = (int64_t)qtile * TQ + wq * SLAB + i * 4 + lq;
                    if (gq >= prm.nq) continue;
                    const uint32_t cq = prm.qcard[gq];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int64_t gp_row = (int64_t)pt * TP + wp * 64 + j * 8 + lp;
                        if (gp_row >= prm.np) continue;
                        uint32_t in = acc[i][j];
                        if (prm.zero_diag && (prm.query_base + gq == prm.pool_base + gp_row)) in = 0u;
                        prm.inter[gq * prm.ld_inter + gp_row] = in;
                        if (prm.score) {
                            const uint32_t un = cq + prm.pcard[gp_row] - in;
                            prm.score[gq * prm.ld_score + gp_row] = in ? (double)in / (double)un : 0.0;
                        }
                    }
                }
            } else {
                named_bar_sync(1, N_CONSUMERS);  // previous tile's scan is finished
#pragma unroll
                for (int i = 0; i < RQ; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        tilebuf[(wq * SLAB + i * 4 + lq) * TILE_LD + wp * 64 + j * 8 + lp] = acc[i][j];
                named_bar_sync(1, N_CONSUMERS);  // count tile complete

                // pool-side constants of the 4 columns this lane scans
                uint32_t cp[4];
                int64_t gpi[4];
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    gpi[s] = (int64_t)pt * TP + s * 32 + lane;
                    cp[s] = gpi[s] < prm.np ? prm.pcard[gpi[s]] : 0u;
                }
                for (int qq = 0; qq < SCAN_ROWS; ++qq) {
                    const int q = cw * SCAN_ROWS + qq;
                    const int64_t gq = (int64_t)qtile * TQ + q;
                    if (gq >= prm.nq) break;  // warp-uniform
                    const uint32_t cq = prm.qcard[gq];
                    WarpTopK<JEntry> tk;
                    tk.k = prm.k;
                    tk.mine = JEntry{l_inter[q * LIST_LD + lane], l_union[q * LIST_LD + lane], l_idx[q * LIST_LD + lane]};
                    tk.refresh_kth();
                    bool changed = false;
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        uint32_t in = tilebuf[q * TILE_LD + s * 32 + lane];
                        const bool valid = gpi[s] < prm.np;
                        if (prm.zero_diag && (prm.query_base + gq == prm.pool_base + gpi[s])) in = 0u;
                        // both sets empty: union 0 -> score 0, represented as 0/1 so cross-multiplication stays valid
                        JEntry c{in, max(cq + cp[s] - in, 1u), (int32_t)(prm.pool_base + gpi[s])};
                        uint32_t m = __ballot_sync(0xffffffffu, valid && JEntry::better(c, tk.kth));
                        while (m) {
                            const int src = __ffs(m) - 1;
                            m &= m - 1;
                            tk.insert(c.shfl(src));
                            changed = true;
                        }
                    }
                    if (changed) {
                        l_inter[q * LIST_LD + lane] = tk.mine.inter;
                        l_union[q * LIST_LD + lane] = tk.mine.uni;
                        l_idx[q * LIST_LD + lane] = tk.mine.idx;
                    }
                }
            }
        }

        if (MODE == MODE_TOPK) {
            __syncwarp();
            // flush this warp's lists to the stripe's partial slot
            for (int qq = 0; qq < SCAN_ROWS; ++qq) {
                const int q = cw * SCAN_ROWS + qq;
                const int64_t gq = (int64_t)qtile * TQ + q;
                if (gq >= prm.nq) break;
                if (lane < prm.k) {
                    const int64_t o = ((int64_t)gq * prm.n_stripes + stripe) * prm.k + lane;   // query-major [nq][n_stripes][k]
                    prm.part[o] = make_uint4(l_inter[q * LIST_LD + lane], l_union[q * LIST_LD + lane],
                                             (uint32_t)l_idx[q * LIST_LD + lane], 0u);
                }
            }
            __syncwarp();
        }
    }
}

// Merge n_lists lists of k_in candidates per query into the best k_out.  One warp per query.
// The kernels' own partial lists (`part`) are query-major, [nq][n_lists][k_in], so that one query's lists and counts
// are contiguous for this kernel; caller planes (all-gather buffers) are list-major, [n_lists][nq][k_in].
// Optional (query-index path): `cnt[q * n_lists + l]` = number of valid, UNSORTED entries of list l (lists of query tiles
// flagged in `tile_dense` are full sorted lists written by the dense kernel); `n_fill` > 0 appends the zero-score
// fillers the query-index kernel never produces: pool rows 0 .. n_fill-1 of this shard unless already listed.
struct MergeExtra {
    const uint32_t* gcount;   // query-index path: candidates offered to the query's contiguous array glist[q][SQ_GC]
    const uint4* glist;
    const uint8_t* cnt;
    const uint32_t* tile_dense;
    const uint32_t* qcard;
    const uint32_t* pcard;
    int64_t pool_base;
    int32_t n_fill;
    int64_t q_off, nq_total;  // fused exchange with query batches: row offset / rows of the whole call
};

// IDX: launched behind the query-index kernel (ex.cnt set); that instance keeps 6 blocks per SM resident.
template <bool IDX>
__global__ void __launch_bounds__(256, IDX ? 6 : 4)
jaccard_merge_kernel(const uint4* __restrict__ part, const uint32_t* __restrict__ inter, const uint32_t* __restrict__ uni,
                     const int32_t* __restrict__ idx, int32_t n_lists, int64_t nq, int32_t k_in, int32_t k_out,
                     uint32_t* __restrict__ out_inter, uint32_t* __restrict__ out_union, int32_t* __restrict__ out_idx,
                     const PeerOut peers, const MergeExtra ex) {
    const int lane = threadIdx.x & 31;
    const int64_t wpg = (int64_t)gridDim.x * (blockDim.x >> 5);
    // candidates come either from the kernels' own partial lists (16-byte entries) or from three caller planes
    auto load = [&](int64_t i) {
        if (part != nullptr) {
            const uint4 e = part[i];
            return JEntry{e.x, e.y, (int32_t)e.z};
        }
        return JEntry{inter[i], uni[i], idx[i]};
    };
    for (int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < nq; q += wpg) {
        WarpTopK<JEntry> tk;
        tk.init(k_out);
        if (IDX && !(ex.tile_dense && ex.tile_dense[q >> 7])) {
            // candidates of the query-index kernel, entries {inter, |pool set|, idx}: union = |q| + |p| - inter.  The
            // first SQ_GC of a query sit in its own contiguous array (coalesced loads); only a query that overflowed
            // it also has entries in the short unsorted per-stripe lists.
            //   pass 1: every lane finds the best of its own candidates; a bitonic sort ranks the 32 lane-bests, the
            //           first k_out of them seed the list (the others cannot be in the top k_out);
            //   pass 2: the remaining candidates (L1/L2 hits now) are inserted only if they beat the current k-th —
            //           a handful per query instead of k (1 + ln(n / k)) serial insertions.
            // unions on this path are below 2^16 (vocabularies of at most SQ_MAX_WORDS * 32 bits): the exact
            // cross-multiplication fits 32 bits
            auto bet = [](const JEntry& a, const JEntry& b) {
                const uint32_t l = a.inter * b.uni, r = b.inter * a.uni;
                return (l > r) || (l == r && a.idx < b.idx);
            };
            const uint32_t cq = ex.qcard[q];
            const uint32_t g_tot = ex.gcount[q];
            const int ng = (int)min(g_tot, (uint32_t)SQ_GC);
            const bool overflow = g_tot > (uint32_t)SQ_GC;   // warp-uniform: later candidates sit in the per-stripe lists
            const uint4* gl = ex.glist + q * SQ_GC;
            JEntry lb = JEntry::worst();
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                auto consume = [&](JEntry(&c)[4]) {
                    if (pass == 0) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (bet(c[j], lb)) lb = c[j];
                    } else {
                        bool any = false;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (c[j].idx == lb.idx) c[j] = JEntry::worst();   // the lane-best was ranked in pass 1
                            any |= bet(c[j], tk.kth);
                        }
                        if (!__ballot_sync(0xffffffffu, any)) return;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint32_t m = __ballot_sync(0xffffffffu, bet(c[j], tk.kth));
                            while (m) {
                                const int src = __ffs(m) - 1;
                                m &= m - 1;
                                tk.insert(c[j].shfl(src));
                            }
                        }
                    }
                };
                // the query's own array: contiguous, four coalesced 16-byte loads in flight per lane
#pragma unroll 1
                for (int e0 = 0; e0 < ng; e0 += 4 * 32) {
                    JEntry c[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int e = e0 + j * 32 + lane;
                        c[j] = JEntry::worst();
                        if (e < ng) {
                            const uint4 x = gl[e];
                            c[j] = JEntry{x.x, cq + x.y - x.x, (int32_t)x.z};
                        }
                    }
                    consume(c);
                }
                if (overflow) {   // lane L walks the stripe lists L, L + 32, ..., four at a time
#pragma unroll 1
                    for (int l0 = 0; l0 < n_lists; l0 += 4 * 32) {
                        int n[4], n_max = 0;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int l = l0 + j * 32 + lane;
                            n[j] = l < n_lists ? (int)ex.cnt[q * n_lists + l] : 0;
                            n_max = max(n_max, n[j]);
                        }
                        n_max = __reduce_max_sync(0xffffffffu, n_max);
                        const uint4* p0 = part + (q * n_lists + l0 + lane) * k_in;   // list j of this lane: p0 + j * 32 * k_in
                        for (int e = 0; e < n_max; ++e) {
                            JEntry c[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                c[j] = JEntry::worst();
                                if (e < n[j]) {
                                    const uint4 x = p0[(int64_t)j * 32 * k_in + e];
                                    c[j] = JEntry{x.x, cq + x.y - x.x, (int32_t)x.z};
                                }
                            }
                            consume(c);
                        }
                    }
                }
                if (pass == 0) {
                    JEntry v = lb;   // bitonic sort of the lane-bests, best first
#pragma unroll
                    for (int k2 = 2; k2 <= 32; k2 <<= 1)
#pragma unroll
                        for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
                            const JEntry o{__shfl_xor_sync(0xffffffffu, v.inter, j2), __shfl_xor_sync(0xffffffffu, v.uni, j2),
                                           __shfl_xor_sync(0xffffffffu, v.idx, j2)};
                            const bool want_better = ((lane & j2) == 0) == ((lane & k2) == 0);
                            if (bet(o, v) == want_better && o.idx != v.idx) v = o;
                        }
                    if (lane < k_out) tk.mine = v;
                    tk.refresh_kth();
                }
            }
        } else if (!IDX && n_lists * k_in <= 6 * 32) {
            // few candidates (the exchange merge: `world` lists of k): all of them are loaded at once, six per lane;
            // the lane-bests, ranked by a bitonic sort, seed the list and only the rest goes through insertions
            const int total = n_lists * k_in;
            JEntry c[6];
            JEntry lb = JEntry::worst();
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const int ci = lane + 32 * i;
                c[i] = JEntry::worst();
                if (ci < total) {
                    const int l = ci / k_in, e = ci - l * k_in;
                    c[i] = load((part != nullptr ? (q * n_lists + l) * k_in : ((int64_t)l * nq + q) * k_in) + e);
                    if (c[i].idx == R4D_IDX_NONE) c[i] = JEntry::worst();
                }
            }
#pragma unroll
            for (int i = 0; i < 6; ++i)
                if (JEntry::better(c[i], lb)) lb = c[i];
            JEntry v = lb;
#pragma unroll
            for (int k2 = 2; k2 <= 32; k2 <<= 1)
#pragma unroll
                for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
                    const JEntry o{__shfl_xor_sync(0xffffffffu, v.inter, j2), __shfl_xor_sync(0xffffffffu, v.uni, j2),
                                   __shfl_xor_sync(0xffffffffu, v.idx, j2)};
                    const bool want_better = ((lane & j2) == 0) == ((lane & k2) == 0);
                    if (JEntry::better(o, v) == want_better && o.idx != v.idx) v = o;
                }
            if (lane < k_out) tk.mine = v;
            tk.refresh_kth();
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                if (32 * i >= total) break;
                uint32_t m = __ballot_sync(0xffffffffu, c[i].idx != lb.idx && JEntry::better(c[i], tk.kth));
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    tk.insert(c[i].shfl(src));
                }
            }
        } else {
            for (int l = 0; l < n_lists; ++l) {
                const int64_t base = part != nullptr ? (q * n_lists + l) * k_in : ((int64_t)l * nq + q) * k_in;
                for (int e0 = 0; e0 < k_in; e0 += 32) {
                    const int e = e0 + lane;
                    JEntry c = JEntry::worst();
                    if (e < k_in) c = load(base + e);
                    uint32_t m = __ballot_sync(0xffffffffu, e < k_in && c.idx != R4D_IDX_NONE && JEntry::better(c, tk.kth));
                    while (m) {
                        const int src = __ffs(m) - 1;
                        m &= m - 1;
                        tk.insert(c.shfl(src));
                    }
                }
            }
        }
        if (ex.n_fill > 0 && tk.kth.inter == 0u) {   // fillers score 0: they only matter while the k-th entry does too
            const uint32_t cq = ex.qcard[q];
            const uint32_t cp = lane < ex.n_fill ? ex.pcard[lane] : 0u;
            for (int i = 0; i < ex.n_fill; ++i) {
                const JEntry c{0u, max(cq + __shfl_sync(0xffffffffu, cp, i), 1u), (int32_t)(ex.pool_base + i)};
                if (__ballot_sync(0xffffffffu, lane < k_out && tk.mine.idx == c.idx)) continue;  // listed already
                tk.insert(c);
            }
        }
        if (lane < k_out) {
            if (peers.world == 0) {
                out_inter[q * k_out + lane] = tk.mine.inter;
                out_union[q * k_out + lane] = tk.mine.uni;
                out_idx[q * k_out + lane] = tk.mine.idx;
            } else {
                // fused exchange: store into slot `rank` of every peer's gather buffer [3][world][nq][k] (NVLink P2P)
                const int64_t nq_all = ex.nq_total > 0 ? ex.nq_total : nq;
                const int64_t plane = (int64_t)peers.world * nq_all * k_out;
                const int64_t at = ((int64_t)peers.rank * nq_all + ex.q_off + q) * k_out + lane;
                for (int r = 0; r < peers.world; ++r) {
                    uint32_t* dst = reinterpret_cast<uint32_t*>(peers.base[r]);
                    dst[at] = tk.mine.inter;
                    dst[plane + at] = tk.mine.uni;
                    dst[2 * plane + at] = (uint32_t)tk.mine.idx;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------- host side
struct JaccardPlan {
    int32_t n_qtiles, n_ptiles, n_stripes, ptiles_per_stripe;
};

static JaccardPlan plan_topk(int64_t nq, int64_t np) {
    JaccardPlan pl;
    pl.n_qtiles = (int32_t)((nq + TQ - 1) / TQ);
    pl.n_ptiles = (int32_t)((np + TP - 1) / TP);
    if (pl.n_qtiles < 1) pl.n_qtiles = 1;
    if (pl.n_ptiles < 1) pl.n_ptiles = 1;
    // ~64 (stripe, query-tile) items per SM: keeps the static round-robin tail below ~2 % and the stripes short
    // enough that the pool windows of the CTAs walking concurrently stay inside L2 (measured DRAM reads at
    // 8 192 x 1 M: 12.9 GB with 32 items/SM, 7.5 GB with 64, 5.4 GB with 128 at +1.6 % time; algorithmic 2.5 GB)
    const int64_t target = (int64_t)num_sms() * 64;
    int64_t stripes = (target + pl.n_qtiles - 1) / pl.n_qtiles;
    if (options().jaccard_stripes > 0) stripes = options().jaccard_stripes;
    if (stripes > pl.n_ptiles) stripes = pl.n_ptiles;
    if (stripes < 1) stripes = 1;
    pl.ptiles_per_stripe = (int32_t)((pl.n_ptiles + stripes - 1) / stripes);
    pl.n_stripes = (pl.n_ptiles + pl.ptiles_per_stripe - 1) / pl.ptiles_per_stripe;
    return pl;
}

static int check_common(const uint32_t* qbits, const uint32_t* qcard, int64_t nq, const uint32_t* pbits,
                        const uint32_t* pcard, int64_t np, int32_t words, int32_t pitch_words) {
    R4D_REQUIRE(nq >= 0 && np >= 0, "jaccard: negative size");
    R4D_REQUIRE(words > 0 && pitch_words >= words && pitch_words % 4 == 0,
                "jaccard: words=%d pitch_words=%d (pitch must be >= words and a multiple of 4)", words, pitch_words);
    R4D_REQUIRE(nq < (int64_t)1 << 31 && np < (int64_t)1 << 31, "jaccard: sizes must be < 2^31");
    if (nq > 0 && np > 0) R4D_REQUIRE(qbits && qcard && pbits && pcard, "jaccard: null pointer");
    return R4D_OK;
}

static int consumer_warps() { return options().jaccard_warps == 8 ? 8 : 16; }

static bool skip_zero_spans() { return options().jaccard_skip_zero != 0; }

template <int MODE, int NCW, bool SKIP>
static int launch_ncw(const CUtensorMap& tm_q, const CUtensorMap& tm_p, const JaccardParams& prm, cudaStream_t st) {
    const size_t smem = smem_bytes_for(MODE);
    static SmemOptIn opt_in;  // one per template instantiation, keyed by device inside
    if (int rc = ensure_dyn_smem(jaccard_kernel<MODE, NCW, SKIP>, smem, opt_in)) return rc;
    const int64_t n_items = (int64_t)prm.n_qtiles * prm.n_stripes;
    int grid = num_sms();
    if (n_items < grid) grid = (int)n_items;
    jaccard_kernel<MODE, NCW, SKIP><<<grid, 32 * NCW, smem, st>>>(tm_q, tm_p, prm); note_launch();
    R4D_CUDA(cudaGetLastError());
    return R4D_OK;
}

template <int MODE>
static int launch(const uint32_t* qbits, int64_t nq, const uint32_t* pbits, int64_t np, int32_t words,
                  int32_t pitch_words, JaccardParams& prm, cudaStream_t st) {
    CUtensorMap tm_q, tm_p;
    int rc = make_tmap_2d(&tm_q, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, qbits, (uint64_t)pitch_words, (uint64_t)nq,
                          (uint64_t)pitch_words * 4, CHUNK_WORDS, TQ, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tmap_2d(&tm_p, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, pbits, (uint64_t)pitch_words, (uint64_t)np,
                      (uint64_t)pitch_words * 4, CHUNK_WORDS, TP, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    prm.n_chunks = (words + CHUNK_WORDS - 1) / CHUNK_WORDS;
    const int rem = words - (prm.n_chunks - 1) * CHUNK_WORDS;  // 1..32 real words in the last chunk
    prm.last_groups = (rem + 3) / 4;
    if (skip_zero_spans())
        return consumer_warps() == 8 ? launch_ncw<MODE, 8, true>(tm_q, tm_p, prm, st)
                                     : launch_ncw<MODE, 16, true>(tm_q, tm_p, prm, st);
    return consumer_warps() == 8 ? launch_ncw<MODE, 8, false>(tm_q, tm_p, prm, st)
                                 : launch_ncw<MODE, 16, false>(tm_q, tm_p, prm, st);
}

}  // namespace r4d

extern "C" {

int r4d_jaccard_full(const uint32_t* qbits, const uint32_t* qcard, int64_t nq, const uint32_t* pbits,
                     const uint32_t* pcard, int64_t np, int32_t words, int32_t pitch_words, int32_t zero_diag,
                     int64_t query_base, int64_t pool_base, uint32_t* inter, int64_t ld_inter, double* score,
                     int64_t ld_score, r4d_stream_t stream) {
    using namespace r4d;
    int rc = check_common(qbits, qcard, nq, pbits, pcard, np, words, pitch_words);
    if (rc) return rc;
    if (nq == 0 || np == 0) return R4D_OK;
    R4D_REQUIRE(inter && ld_inter >= np, "jaccard_full: inter null or ld_inter < np");
    R4D_REQUIRE(!score || ld_score >= np, "jaccard_full: ld_score < np");
    JaccardParams prm{};
    prm.qcard = qcard;
    prm.pcard = pcard;
    prm.nq = nq;
    prm.np = np;
    prm.k = 0;
    prm.zero_diag = zero_diag;
    prm.query_base = query_base;
    prm.pool_base = pool_base;
    prm.n_qtiles = (int32_t)((nq + TQ - 1) / TQ);
    prm.n_ptiles = (int32_t)((np + TP - 1) / TP);
    prm.n_stripes = prm.n_ptiles;  // one pool tile per work item
    prm.ptiles_per_stripe = 1;
    prm.inter = inter;
    prm.ld_inter = ld_inter;
    prm.score = score;
    prm.ld_score = ld_score;
    return launch<MODE_FULL>(qbits, nq, pbits, np, words, pitch_words, prm, as_stream(stream));
}

size_t r4d_jaccard_topk_workspace_bytes(int64_t nq, int64_t np, int32_t k) {
    using namespace r4d;
    if (nq <= 0 || np <= 0 || k <= 0) return 256;
    const JaccardPlan pl = plan_topk(nq, np);
    size_t need = (size_t)pl.n_stripes * (size_t)nq * (size_t)k * SQ_PART_BYTES + 256;
    // query-index path: one index for the whole call + a per-batch region that the <= SQ_QB-row batches reuse
    size_t batch = 0;
    const int64_t sizes[2] = {nq < SQ_QB ? nq : (int64_t)SQ_QB, nq % SQ_QB};
    for (int64_t nb : sizes) {
        if (nb <= 0) continue;
        const JaccardPlan pb = plan_topk(nb, np);
        const size_t b = (size_t)pb.n_stripes * (size_t)nb * (size_t)k * SQ_PART_BYTES + 256 + sparseq_batch_bytes(nb, pb.n_stripes);
        if (b > batch) batch = b;
    }
    batch += sparseq_index_bytes(nq) + 256;
    return need > batch ? need : batch;
}

static int merge_launch(const uint4* part, const uint32_t* inter, const uint32_t* uni, const int32_t* idx, int32_t n_lists, int64_t nq,
                        int32_t k_in, int32_t k_out, uint32_t* out_inter, uint32_t* out_union, int32_t* out_idx,
                        const r4d::PeerOut& peers, const r4d::MergeExtra& ex, r4d_stream_t stream) {
    using namespace r4d;
    int64_t blocks = (nq + 7) / 8;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (ex.cnt != nullptr)
        jaccard_merge_kernel<true><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(part, inter, uni, idx, n_lists, nq, k_in, k_out,
                                                                                   out_inter, out_union, out_idx, peers, ex);
    else
        jaccard_merge_kernel<false><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(part, inter, uni, idx, n_lists, nq, k_in, k_out,
                                                                                    out_inter, out_union, out_idx, peers, ex);
    note_launch();
    R4D_CUDA(cudaGetLastError());
    return R4D_OK;
}

static int make_peers(r4d::PeerOut& po, void* const* peer_base, int32_t world, int32_t rank) {
    using namespace r4d;
    R4D_REQUIRE(peer_base && world >= 1 && world <= R4D_MAX_PEERS && rank >= 0 && rank < world,
                "fused exchange: world=%d rank=%d (max %d peers)", world, rank, R4D_MAX_PEERS);
    po.world = world;
    po.rank = rank;
    for (int r = 0; r < world; ++r) {
        R4D_REQUIRE(peer_base[r] != nullptr, "fused exchange: null peer pointer %d", r);
        po.base[r] = peer_base[r];
    }
    return R4D_OK;
}

static int jaccard_topk_impl(const uint32_t* qbits, const uint32_t* qcard, int64_t nq, const uint32_t* pbits,
                             const uint32_t* pcard, int64_t np, int32_t words, int32_t pitch_words, int32_t k,
                             int32_t zero_diag, int64_t query_base, int64_t pool_base, uint32_t* top_inter,
                             uint32_t* top_union, int32_t* top_idx, const r4d::PeerOut& peers, void* workspace,
                             size_t workspace_bytes, r4d_stream_t stream) {
    using namespace r4d;
    int rc = check_common(qbits, qcard, nq, pbits, pcard, np, words, pitch_words);
    if (rc) return rc;
    R4D_REQUIRE(k >= 1 && k <= R4D_TOPK_MAX, "jaccard_topk: k=%d out of range [1, %d]", k, R4D_TOPK_MAX);
    R4D_REQUIRE(pool_base >= 0 && pool_base + np < (int64_t)R4D_IDX_NONE, "jaccard_topk: pool_base+np exceeds int32");
    if (nq == 0) return R4D_OK;
    R4D_REQUIRE(peers.world > 0 || (top_inter && top_union && top_idx), "jaccard_topk: null output");
    cudaStream_t st = as_stream(stream);
    const MergeExtra no_extra{};
    if (np == 0)  // no pool rows: every list is padding; the merge of zero lists writes it, no workspace needed
        return merge_launch(nullptr, nullptr, nullptr, nullptr, 0, nq, k, k, top_inter, top_union, top_idx, peers, no_extra, stream);
    const bool use_index = sparseq_supported(words, k);
    // query-index path: batches of <= SQ_QB query rows, each a launch sequence on the same per-batch region; the by-row
    // index of ALL batches is built up front by one launch
    const int64_t q_batch = use_index ? (int64_t)SQ_QB : nq;
    const size_t idx_bytes = use_index ? (sparseq_index_bytes(nq) + 255) / 256 * 256 : 0;
    uint8_t* const batch_base = reinterpret_cast<uint8_t*>(workspace) + idx_bytes;
    QIndex q_all{};
    if (use_index) {
        if (workspace_bytes < idx_bytes || !workspace) {
            set_error("jaccard_topk: workspace %zu B < required %zu B", workspace_bytes, idx_bytes);
            return R4D_E_WORKSPACE;
        }
        R4D_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "jaccard_topk: workspace must be 16-byte aligned");
        q_all = sparseq_carve_index(workspace, nq);
        rc = sparseq_build_all(qbits, nq, words, pitch_words, q_all, st);
        if (rc) return rc;
    }
    for (int64_t q0 = 0; q0 < nq; q0 += q_batch) {
        const int64_t nb = nq - q0 < q_batch ? nq - q0 : q_batch;
        const JaccardPlan pl = plan_topk(nb, np);
        const size_t per = (size_t)pl.n_stripes * (size_t)nb * (size_t)k;
        const size_t need = idx_bytes + per * SQ_PART_BYTES + (use_index ? 256 + sparseq_batch_bytes(nb, pl.n_stripes) : 0);
        if (workspace_bytes < need || !workspace) {
            set_error("jaccard_topk: workspace %zu B < required %zu B", workspace_bytes, need);
            return R4D_E_WORKSPACE;
        }
        R4D_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "jaccard_topk: workspace must be 16-byte aligned");
        const uint32_t* qb = qbits + q0 * pitch_words;
        JaccardParams prm{};
        prm.qcard = qcard + q0;
        prm.pcard = pcard;
        prm.nq = nb;
        prm.np = np;
        prm.k = k;
        prm.zero_diag = zero_diag;
        prm.query_base = query_base + q0;
        prm.pool_base = pool_base;
        prm.n_qtiles = pl.n_qtiles;
        prm.n_ptiles = pl.n_ptiles;
        prm.n_stripes = pl.n_stripes;
        prm.ptiles_per_stripe = pl.ptiles_per_stripe;
        prm.part = reinterpret_cast<uint4*>(batch_base);
        MergeExtra ex{};
        ex.q_off = q0;
        ex.nq_total = nq;
        if (use_index) {
            // sparse query tiles -> jaccard_qindex_kernel; tiles flagged dense -> the kernel below (same partial slots)
            const QIndex qi = sparseq_batch_view(q_all, q0, batch_base + per * SQ_PART_BYTES, nb, pl.n_stripes);
            rc = sparseq_clear_batch(qi, nb, pl.n_stripes, st);
            if (rc) return rc;
            rc = sparseq_topk_launch(pbits, prm.qcard, pcard, nb, np, words, pitch_words, k, zero_diag, prm.query_base,
                                     pool_base, pl.n_qtiles, pl.n_ptiles, pl.n_stripes, pl.ptiles_per_stripe,
                                     prm.part, qi, st);
            if (rc) return rc;
            prm.tile_filter = qi.tile_dense;
            prm.any_filtered = qi.any_dense;
            ex.cnt = qi.cnt;
            ex.gcount = qi.gcount;
            ex.glist = qi.glist;
            ex.tile_dense = qi.tile_dense;
            ex.qcard = prm.qcard;
            ex.pcard = pcard;
            ex.pool_base = pool_base;
            ex.n_fill = (int32_t)(np < k ? np : k);
        }
        rc = launch<MODE_TOPK>(qb, nb, pbits, np, words, pitch_words, prm, st);
        if (rc) return rc;
        rc = merge_launch(prm.part, nullptr, nullptr, nullptr, pl.n_stripes, nb, k, k,
                          top_inter ? top_inter + q0 * k : nullptr, top_union ? top_union + q0 * k : nullptr,
                          top_idx ? top_idx + q0 * k : nullptr, peers, ex, stream);
        if (rc) return rc;
    }
    return R4D_OK;
}

int r4d_jaccard_topk(const uint32_t* qbits, const uint32_t* qcard, int64_t nq, const uint32_t* pbits,
                     const uint32_t* pcard, int64_t np, int32_t words, int32_t pitch_words, int32_t k,
                     int32_t zero_diag, int64_t query_base, int64_t pool_base, uint32_t* top_inter,
                     uint32_t* top_union, int32_t* top_idx, void* workspace, size_t workspace_bytes,
                     r4d_stream_t stream) {
    r4d::PeerOut none{};
    return jaccard_topk_impl(qbits, qcard, nq, pbits, pcard, np, words, pitch_words, k, zero_diag, query_base, pool_base,
                             top_inter, top_union, top_idx, none, workspace, workspace_bytes, stream);
}

int r4d_jaccard_topk_scatter(const uint32_t* qbits, const uint32_t* qcard, int64_t nq, const uint32_t* pbits,
                             const uint32_t* pcard, int64_t np, int32_t words, int32_t pitch_words, int32_t k,
                             int32_t zero_diag, int64_t query_base, int64_t pool_base, void* const* peer_base,
                             int32_t world, int32_t rank, void* workspace, size_t workspace_bytes, r4d_stream_t stream) {
    r4d::PeerOut po{};
    int rc = make_peers(po, peer_base, world, rank);
    if (rc) return rc;
    return jaccard_topk_impl(qbits, qcard, nq, pbits, pcard, np, words, pitch_words, k, zero_diag, query_base, pool_base,
                             nullptr, nullptr, nullptr, po, workspace, workspace_bytes, stream);
}

int r4d_jaccard_topk_merge(const uint32_t* inter, const uint32_t* uni, const int32_t* idx, int32_t n_lists,
                           int64_t nq, int32_t k_in, int32_t k_out, uint32_t* out_inter, uint32_t* out_union,
                           int32_t* out_idx, r4d_stream_t stream) {
    using namespace r4d;
    R4D_REQUIRE(n_lists >= 0 && nq >= 0 && k_in >= 1 && k_out >= 1 && k_out <= R4D_TOPK_MAX,
                "jaccard_topk_merge: n_lists=%d k_in=%d k_out=%d", n_lists, k_in, k_out);
    if (nq == 0) return R4D_OK;
    R4D_REQUIRE(out_inter && out_union && out_idx, "jaccard_topk_merge: null output");
    R4D_REQUIRE(n_lists == 0 || (inter && uni && idx), "jaccard_topk_merge: null input");
    r4d::PeerOut none{};
    return merge_launch(nullptr, inter, uni, idx, n_lists, nq, k_in, k_out, out_inter, out_union, out_idx, none, MergeExtra{}, stream);
}

}  // extern "C"
