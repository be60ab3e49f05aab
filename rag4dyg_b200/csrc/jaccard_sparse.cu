// jaccard_sparse.cu — fused Jaccard top-K with a QUERY-SIDE WORD INDEX (pool side unchanged: bitset tiles streamed by
// TMA, AND + POPC, exact rational ranking).
//
// Why: node-id sets are tiny next to the vocabulary (2.2 of 20 000 bits), so a query row has ~2 non-zero words out of
// 625 and only ~2.4e-4 of all (query, pool) pairs intersect at all.  The dense kernel (jaccard.cu) re-streams the pool
// for every 128-query tile and walks every 8-word span of every pair to find that out.  Here
//   1. qindex_kernel turns a query batch (<= 8 192 rows) ONCE into the list of its non-zero words, by row;
//   2. jaccard_qindex_kernel: every CTA owns a pool stripe.  It loads the index of a whole group of query tiles into
//      shared memory re-sorted BY WORD (counting sort), then streams its pool bitsets through a TMA ring exactly once
//      per group.  Each warp owns 8 rows of the 128-row pool tile: two conflict-free LDS.128 per lane per 32-word
//      chunk bring the warp's slice into registers (the ring stage is released right away), a ballot finds the
//      non-zero pool words, and each of them looks up the query entries with the same word id: AND + POPC.
//   3. A hit (query r, pool row p, word w) is completed on the spot: the warp re-reads row r's entries (by-row
//      index, L2) against pool row p, sums the POPCs (full intersection) and keeps the hit only if w is the FIRST
//      intersecting word of the pair, so every intersecting pair is emitted exactly once without any accumulator tile.
//   4. Emitted candidates go to the (stripe, query) partial list in global memory (append while short, replace-the-
//      worst when full) under a per-query lock bit in shared memory.
// Zero-score candidates are never produced here: they only matter as the lowest-index filler of a short list, which
// the merge kernel adds (jaccard.cu, jaccard_merge_kernel `n_fill`).
//
// Query tiles with more than SQ_T1 non-zero words (dense data, e.g. history sets) are flagged by qindex_kernel and
// handled by the dense kernel in the same launch sequence; both write the same per-stripe partial lists.
// Results are bit-identical to the dense kernel (tests/test_gpu_jaccard.py).
#include "jaccard_common.cuh"

namespace r4d {

// ---------------------------------------------------------------------------- by-row index of a query tile
// grid = n_qtiles CTAs of 1024 threads (32 warps, 4 rows each).
constexpr int QI_THREADS = 1024;
__global__ void __launch_bounds__(QI_THREADS)
qindex_kernel(const uint32_t* __restrict__ qbits, int64_t nq, int32_t words, int32_t pitch_words, QIndex qi) {
    __shared__ uint32_t rowcnt[SQ_TQ];
    __shared__ uint32_t rowstart[SQ_TQ + 1];
    const int t = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = warp; i < SQ_TQ; i += QI_THREADS / 32) {
        const int64_t gq = (int64_t)t * SQ_TQ + i;
        uint32_t c = 0;
        if (gq < nq) {
            const uint32_t* row = qbits + gq * pitch_words;
            for (int w0 = 0; w0 < words; w0 += 32) {
                const int w = w0 + lane;
                const uint32_t v = w < words ? row[w] : 0u;
                c += __popc(__ballot_sync(0xffffffffu, v != 0u));
            }
        }
        if (lane == 0) rowcnt[i] = c;
    }
    __syncthreads();
    if (warp == 0) {  // exclusive scan of 128 row counts, 4 per lane
        uint32_t v[4], sum = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j] = rowcnt[lane * 4 + j];
            sum += v[j];
        }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += x;
        }
        uint32_t run = incl - sum;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            rowstart[lane * 4 + j] = run;
            run += v[j];
        }
        if (lane == 31) rowstart[SQ_TQ] = incl;
    }
    __syncthreads();
    const uint32_t total = rowstart[SQ_TQ];
    if (total > (uint32_t)SQ_T1) {  // too dense for the index path: the dense kernel takes this tile
        if (threadIdx.x == 0) {
            qi.tile_dense[t] = 1u;
            qi.tile_cnt[t] = 0u;
        }
        return;
    }
    if (threadIdx.x == 0) {
        qi.tile_dense[t] = 0u;
        qi.tile_cnt[t] = total;
    }
    uint16_t* rowoff = qi.rowoff + (size_t)t * SQ_ROWOFF_LD;
    for (int i = threadIdx.x; i <= SQ_TQ; i += QI_THREADS) rowoff[i] = (uint16_t)rowstart[i];
    uint16_t* ew = qi.ent_word + (size_t)t * SQ_T1;
    uint32_t* ev = qi.ent_val + (size_t)t * SQ_T1;
    for (int i = warp; i < SQ_TQ; i += QI_THREADS / 32) {
        const int64_t gq = (int64_t)t * SQ_TQ + i;
        if (gq >= nq) continue;
        const uint32_t* row = qbits + gq * pitch_words;
        uint32_t base = rowstart[i];
        for (int w0 = 0; w0 < words; w0 += 32) {
            const int w = w0 + lane;
            const uint32_t v = w < words ? row[w] : 0u;
            const uint32_t b = __ballot_sync(0xffffffffu, v != 0u);
            if (v != 0u) {
                const uint32_t pos = base + __popc(b & ((1u << lane) - 1u));
                ew[pos] = (uint16_t)w;
                ev[pos] = v;
            }
            base += __popc(b);
        }
    }
}

// ---------------------------------------------------------------------------- main kernel
struct SparseParams {
    const uint32_t* pbits;
    const uint32_t* qcard;
    const uint32_t* pcard;
    int64_t nq, np;
    int32_t pitch_words, n_words, n_chunks, k, zero_diag;
    int64_t query_base, pool_base;
    int32_t n_qtiles, n_ptiles, n_stripes, ptiles_per_stripe;
    uint32_t* part_inter;
    uint32_t* part_union;
    int32_t* part_idx;
    QIndex qi;
    int32_t debug;  // experiments: 1 = scan only (no lookups), 2 = lookups but no hit completion, 3 = no list update
};

__device__ __forceinline__ uint32_t atoms_or(uint32_t addr, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ uint32_t atoms_add(uint32_t addr, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void atoms_and(uint32_t addr, uint32_t v) {
    asm volatile("red.shared.and.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) {
    asm volatile("st.volatile.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v;
    asm volatile("{\n\t.reg .u16 t;\n\tld.shared.u16 t, [%1];\n\tcvt.u32.u16 %0, t;\n\t}" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void lds128s(uint4& v, uint32_t addr) {
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
}

// Append the candidates of the lanes with `keep` set, one by one, to their (stripe, query) partial list in global
// memory: unsorted while the list is short, replace-the-worst once it holds k entries; a per-query lock bit in shared
// memory serialises the warps of the CTA that hit the same query.
__device__ __noinline__ void append_candidates(const SparseParams& prm, bool keep, uint32_t rr, uint32_t inter, uint32_t uni,
                                               int32_t idx, int t0, int stripe, uint32_t lock_u32, uint32_t count_u32) {
    const int lane = threadIdx.x & 31;
    const int K = prm.k;
    uint32_t pb = __ballot_sync(0xffffffffu, keep);
    while (pb) {
        const int src = __ffs(pb) - 1;
        pb &= pb - 1;
        const uint32_t r = __shfl_sync(0xffffffffu, rr, src);
        const JEntry cand{__shfl_sync(0xffffffffu, inter, src), __shfl_sync(0xffffffffu, uni, src),
                          __shfl_sync(0xffffffffu, idx, src)};
        const int64_t gq = (int64_t)t0 * SQ_TQ + r;
        const int64_t base = ((int64_t)stripe * prm.nq + gq) * K;
        const uint32_t bit = 1u << (r & 31);
        const uint32_t lock_a = lock_u32 + (r >> 5) * 4u, count_a = count_u32 + r;
        if (lane == 0) {
            uint32_t polls = 0;
            while (atoms_or(lock_a, bit) & bit)
                if (++polls > (1u << 28)) __trap();  // a lost unlock must surface as a launch failure, not a hang
        }
        __syncwarp();
        __threadfence_block();
        // read by lane 0 only: after __syncwarp the lanes may still run as separate groups, and lane 0's own update
        // below must not be seen by lanes that read later (n has to be warp-uniform)
        int n = 0;
        if (lane == 0) n = (int)lds_u8(count_a);
        n = __shfl_sync(0xffffffffu, n, 0);
        if (prm.debug == 4) {
        } else if (n < K) {
            if (lane == 0) {
                prm.part_inter[base + n] = cand.inter;
                prm.part_union[base + n] = cand.uni;
                prm.part_idx[base + n] = cand.idx;
                sts_u8(count_a, (uint32_t)(n + 1));
            }
        } else if (prm.debug == 5) {
        } else {  // full: the candidate replaces the worst entry if it ranks before it
            JEntry wv = lane < K ? JEntry{__ldcg(prm.part_inter + base + lane), __ldcg(prm.part_union + base + lane),
                                          __ldcg(prm.part_idx + base + lane)}
                                 : JEntry{0xffffffffu, 1u, -1};  // ranks before every real entry
            int wl = lane;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                const JEntry ov{__shfl_xor_sync(0xffffffffu, wv.inter, o), __shfl_xor_sync(0xffffffffu, wv.uni, o),
                                __shfl_xor_sync(0xffffffffu, wv.idx, o)};
                const int ol = __shfl_xor_sync(0xffffffffu, wl, o);
                if (JEntry::better(wv, ov)) {
                    wv = ov;
                    wl = ol;
                }
            }
            if (lane == 0 && JEntry::better(cand, wv)) {
                prm.part_inter[base + wl] = cand.inter;
                prm.part_union[base + wl] = cand.uni;
                prm.part_idx[base + wl] = cand.idx;
            }
        }
        // entries must be in L2 before another warp can find the list full and read them back (ld.cg)
        if (n + 1 >= K) __threadfence();
        else __threadfence_block();
        __syncwarp();
        if (lane == 0) atoms_and(lock_a, ~bit);
    }
}

// Fallback completion (pool rows too dense for the per-tile lists): up to 32 queued hits, lane i holds hit i
// (a = query row | word << 13, b = pool row relative to the stripe).  Per lane: the row's entries (by-row index, L2)
// against the pool row give the full intersection and the pair's first intersecting word; only the hit AT that word
// is kept, so a pair is emitted exactly once whatever the order the hits arrive in.
__device__ __noinline__ void flush_hits(const SparseParams& prm, uint32_t hit_a, uint32_t hit_b, int qn, int t0, int stripe,
                                        int pt_beg, uint32_t lock_u32, uint32_t count_u32) {
    const int lane = threadIdx.x & 31;
    bool primary = false;
    uint32_t rr = 0, inter = 0, uni = 0;
    int32_t idx = 0;
    if (lane < qn && prm.debug != 2) {
        rr = hit_a & 0x1fffu;
        const uint32_t w = hit_a >> 13;
        const int64_t gp = (int64_t)pt_beg * SQ_TP + hit_b;
        const int t = t0 + (int)(rr >> 7), i = (int)(rr & (SQ_TQ - 1));
        const int64_t gq = (int64_t)t0 * SQ_TQ + rr;
        if (!(prm.zero_diag && prm.query_base + gq == prm.pool_base + gp)) {  // the diagonal is a forced zero: a filler
            const uint16_t* ro = prm.qi.rowoff + (size_t)t * SQ_ROWOFF_LD + i;
            const int rb = ro[0], re = ro[1];
            const uint32_t cq = prm.qcard[gq], cp = prm.pcard[gp];
            const uint16_t* ew = prm.qi.ent_word + (size_t)t * SQ_T1;
            const uint32_t* ev = prm.qi.ent_val + (size_t)t * SQ_T1;
            const uint32_t* prow = prm.pbits + gp * prm.pitch_words;
            uint32_t first = 0xffffffffu;
            for (int e = rb; e < re; ++e) {
                const uint32_t ww = ew[e];
                const uint32_t c = __popc(ev[e] & __ldg(prow + ww));
                inter += c;
                if (c) first = min(first, ww);
            }
            primary = first == w && prm.debug != 3;
            uni = cq + cp - inter;
            idx = (int32_t)(prm.pool_base + gp);
        }
    }
    append_candidates(prm, primary, rr, inter, uni, idx, t0, stripe, lock_u32, count_u32);
}

// The warp's share of one pool tile (8 rows), after all its chunks were scanned: `pw_n` non-zero pool words
// (value, word id | row-in-warp << 11) wait in shared memory.  Each looks up the query entries with its word id
// (AND + POPC); hits (query row, pool row, count) go to the warp's hit buffer; hits of the same pair are summed --
// every word of both rows has been seen, so the sum IS the intersection -- and the pair is emitted once.
// Returns 1 (nothing emitted) when the hit buffer overflows: the caller replays the tile through flush_hits.
constexpr int SQ_PW_CAP = 64;   // non-zero pool words per warp per tile kept for the lookup phase
constexpr int SQ_HB_CAP = 64;   // hits per warp per tile
__device__ __noinline__ int tile_hits(const SparseParams& prm, uint32_t pw_a, int pw_n, uint32_t hb_a, uint32_t off_a,
                                      uint32_t val_a, uint32_t row_a, int t0, int stripe, int64_t gp0, uint32_t lock_u32,
                                      uint32_t count_u32) {
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    int hn = 0;
    for (int i = 0; i < pw_n; ++i) {
        const uint32_t pv = lds_u32(pw_a + i * 8), meta = lds_u32(pw_a + i * 8 + 4);
        const uint32_t w = meta & 0x7ffu, pl = meta >> 11;
        const int beg = (int)lds_u32(off_a + w * 4), end = (int)lds_u32(off_a + w * 4 + 4);
        for (int e0 = beg; e0 < end; e0 += 32) {
            const int e = e0 + lane;
            uint32_t x = 0;
            if (e < end) x = lds_u32(val_a + e * 4) & pv;
            const uint32_t mb = __ballot_sync(0xffffffffu, x != 0u);
            if (!mb) continue;
            const int add = __popc(mb);
            if (hn + add > SQ_HB_CAP) return 1;
            if (x) sts_u32(hb_a + (hn + __popc(mb & lt)) * 4, lds_u16(row_a + e * 2) | ((uint32_t)__popc(x) << 13) | (pl << 19));
            hn += add;
        }
    }
    if (hn == 0 || prm.debug == 2) return 0;
    __syncwarp();
    constexpr uint32_t KEY = 0x1fffu | (7u << 19);
    for (int b = 0; b < hn; b += 32) {
        const int i = b + lane;
        const uint32_t mine = i < hn ? lds_u32(hb_a + i * 4) : 0u;
        uint32_t sum = 0;
        bool leader = i < hn;
        for (int j = 0; j < hn; ++j) {
            const uint32_t hj = lds_u32(hb_a + j * 4);
            if (((hj ^ mine) & KEY) == 0u) {
                sum += (hj >> 13) & 63u;
                if (j < i) leader = false;
            }
        }
        const uint32_t rr = mine & 0x1fffu;
        const int64_t gq = (int64_t)t0 * SQ_TQ + rr, gp = gp0 + (mine >> 19);
        if (prm.zero_diag && prm.query_base + gq == prm.pool_base + gp) leader = false;  // forced zero: a filler
        if (prm.debug == 3) leader = false;
        uint32_t uni = 0;
        if (leader) uni = prm.qcard[gq] + prm.pcard[gp] - sum;
        append_candidates(prm, leader, rr, sum, uni, (int32_t)(prm.pool_base + gp), t0, stripe, lock_u32, count_u32);
    }
    return 0;
}

// shared-memory layout (offsets from the 1024-aligned base)
constexpr int SQ_OFF_WORDS = SQ_MAX_WORDS + 5;  // off[0 .. n_words], padded
constexpr size_t SQ_SM_STAGES = 0;
constexpr size_t SQ_SM_BARS = SQ_SM_STAGES + (size_t)SQ_STAGES * SQ_STAGE_BYTES;
constexpr size_t SQ_SM_OFF = SQ_SM_BARS + 128;
constexpr size_t SQ_SM_VAL = SQ_SM_OFF + (size_t)SQ_OFF_WORDS * 4;
constexpr size_t SQ_SM_ROW = SQ_SM_VAL + (size_t)SQ_E_CAP * 4;
constexpr size_t SQ_SM_COUNT = SQ_SM_ROW + (size_t)SQ_E_CAP * 2;
constexpr size_t SQ_SM_LOCK = SQ_SM_COUNT + SQ_QB;
constexpr size_t SQ_SM_GROUPS = SQ_SM_LOCK + SQ_QB / 8;
constexpr size_t SQ_SM_SCAN = SQ_SM_GROUPS + (size_t)(SQ_MAX_TILES + 2) * 8;
constexpr size_t SQ_SM_PW = SQ_SM_SCAN + 32 * 4;                                  // [16 warps][SQ_PW_CAP] x 8 B
constexpr size_t SQ_SM_HB = SQ_SM_PW + (size_t)SQ_WARPS * SQ_PW_CAP * 8;          // [16 warps][SQ_HB_CAP] x 4 B
constexpr size_t SQ_SM_TOTAL = SQ_SM_HB + (size_t)SQ_WARPS * SQ_HB_CAP * 4 + 1024;  // + alignment slack
static_assert(SQ_SM_TOTAL <= 227 * 1024, "query-index kernel: shared memory budget");

__global__ void __launch_bounds__(SQ_THREADS + 32, 1)
jaccard_qindex_kernel(const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ SparseParams prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stages = smem + SQ_SM_STAGES;                                   // [STAGES][128 rows x 128 B], swizzled
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SQ_SM_BARS);
    uint64_t* empty_bar = full_bar + 8;
    uint32_t* off = reinterpret_cast<uint32_t*>(smem + SQ_SM_OFF);           // entries of word w: [off[w], off[w+1])
    uint32_t* val_s = reinterpret_cast<uint32_t*>(smem + SQ_SM_VAL);         // [E] word value, sorted by word id
    uint16_t* row_s = reinterpret_cast<uint16_t*>(smem + SQ_SM_ROW);         // [E] group-relative query row
    volatile uint8_t* count = reinterpret_cast<volatile uint8_t*>(smem + SQ_SM_COUNT);  // [rows] candidates stored
    uint32_t* lock = reinterpret_cast<uint32_t*>(smem + SQ_SM_LOCK);         // one bit per group row
    int32_t* g_first = reinterpret_cast<int32_t*>(smem + SQ_SM_GROUPS);      // [n_groups + 1] first tile of a group
    int32_t* g_ent = g_first + SQ_MAX_TILES + 2;                             // [n_groups] entries of the group
    uint32_t* scan_s = reinterpret_cast<uint32_t*>(smem + SQ_SM_SCAN);       // [17] warp totals, [31] = n_groups

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_p);
        for (int s = 0; s < SQ_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], SQ_WARPS);
        }
        fence_barrier_init();
        // greedy packing of consecutive query tiles into groups of <= SQ_E_CAP entries (every CTA computes the same)
        int g = 0, sum = 0;
        g_first[0] = 0;
        for (int t = 0; t < prm.n_qtiles; ++t) {
            const int c = (int)prm.qi.tile_cnt[t];
            if (sum + c > SQ_E_CAP) {
                g_ent[g] = sum;
                g_first[++g] = t;
                sum = 0;
            }
            sum += c;
        }
        g_ent[g] = sum;
        g_first[g + 1] = prm.n_qtiles;
        scan_s[31] = (uint32_t)(g + 1);
    }
    __syncthreads();
    const int n_groups = (int)scan_s[31];
    const int n_items = n_groups * prm.n_stripes;

    // ---- dedicated TMA producer (warp 16)
    if (warp == SQ_WARPS) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int g = item / prm.n_stripes, stripe = item - g * prm.n_stripes;
                if (g_ent[g] == 0) continue;
                const int pt_beg = stripe * prm.ptiles_per_stripe;
                const int pt_end = min(pt_beg + prm.ptiles_per_stripe, prm.n_ptiles);
                for (int pt = pt_beg; pt < pt_end; ++pt)
                    for (int c = 0; c < prm.n_chunks; ++c) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&full_bar[stage], SQ_STAGE_BYTES);
                        tma_load_2d(stages + (size_t)stage * SQ_STAGE_BYTES, &tm_p, &full_bar[stage], c * SQ_CHUNK_WORDS,
                                    pt * SQ_TP);
                        if (++stage == SQ_STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
            }
        }
        return;
    }

    const uint32_t stages_u32 = smem_u32(stages);
    const uint32_t lock_u32 = smem_u32(lock), count_u32 = smem_u32(smem + SQ_SM_COUNT), off_u32 = smem_u32(off);
    int stage = 0;
    uint32_t phase = 0;
    int cur_g = -1;
    uint32_t hit_a = 0, hit_b = 0;  // fallback path: register-resident hit queue, lane i holds hit i
    int qn = 0;
    const uint32_t pw_a = smem_u32(smem + SQ_SM_PW) + (uint32_t)warp * SQ_PW_CAP * 8u;  // this warp's pool-word list
    const uint32_t hb_a = smem_u32(smem + SQ_SM_HB) + (uint32_t)warp * SQ_HB_CAP * 4u;  // this warp's hit buffer
    int pw_n = 0;

    // this lane's two 16-byte units of every chunk: rows warp*8 + (lane>>3) + {0, 4}, unit lane&7 (128B swizzle)
    const int prow0 = warp * 8 + (lane >> 3);
    const uint32_t unit = (uint32_t)(lane & 7);
    const uint32_t a0 = (uint32_t)prow0 * 128u + ((unit ^ (uint32_t)(prow0 & 7)) << 4);
    const uint32_t a1 = (uint32_t)(prow0 + 4) * 128u + ((unit ^ (uint32_t)((prow0 + 4) & 7)) << 4);

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int g = item / prm.n_stripes, stripe = item - g * prm.n_stripes;
        if (g_ent[g] == 0) continue;  // block-uniform: nothing but empty / dense-flagged tiles
        const int t0 = g_first[g], t1 = g_first[g + 1];
        const int64_t row0 = (int64_t)t0 * SQ_TQ;                                  // first query row of the group
        const int g_rows = (int)(min((int64_t)t1 * SQ_TQ, prm.nq) - row0);
        const int pt_beg = stripe * prm.ptiles_per_stripe;
        const int pt_end = min(pt_beg + prm.ptiles_per_stripe, prm.n_ptiles);

        named_bar_sync(1, SQ_THREADS);  // previous item: all hits handled, counts flushed
        if (g != cur_g) {               // ---- load the group's entries re-sorted by word id (counting sort)
            cur_g = g;
            const int nw = prm.n_words;
            for (int i = threadIdx.x; i <= nw; i += SQ_THREADS) off[i] = 0u;
            named_bar_sync(1, SQ_THREADS);
            for (int t = t0; t < t1; ++t) {
                const int n = (int)prm.qi.tile_cnt[t];
                const uint16_t* ew = prm.qi.ent_word + (size_t)t * SQ_T1;
                for (int e = threadIdx.x; e < n; e += SQ_THREADS) atoms_add(off_u32 + ((uint32_t)ew[e] + 1u) * 4u, 1u);
            }
            named_bar_sync(1, SQ_THREADS);
            {   // exclusive scan of off[1 .. nw] in place: off[w + 1] = first entry of word w
                const int per = (nw + SQ_THREADS - 1) / SQ_THREADS;  // <= 4
                const int b = 1 + threadIdx.x * per;
                uint32_t v[4], sum = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[j] = (j < per && b + j <= nw) ? off[b + j] : 0u;
                    sum += v[j];
                }
                uint32_t incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += x;
                }
                if (lane == 31) scan_s[warp] = incl;
                named_bar_sync(1, SQ_THREADS);
                uint32_t wbase = 0;
                for (int w = 0; w < warp; ++w) wbase += scan_s[w];
                uint32_t run = wbase + incl - sum;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < per && b + j <= nw) {
                        off[b + j] = run;
                        run += v[j];
                    }
            }
            named_bar_sync(1, SQ_THREADS);
            for (int t = t0; t < t1; ++t) {
                if (prm.qi.tile_cnt[t] == 0u) continue;  // empty or dense-flagged tile (its row offsets are not written)
                const uint16_t* ro = prm.qi.rowoff + (size_t)t * SQ_ROWOFF_LD;
                const uint16_t* ew = prm.qi.ent_word + (size_t)t * SQ_T1;
                const uint32_t* ev = prm.qi.ent_val + (size_t)t * SQ_T1;
                // one thread per (row, entry-of-row): rows are short, so walk rows and let lanes take entries
                for (int i = warp; i < SQ_TQ; i += SQ_WARPS) {
                    const int rb = ro[i], re = ro[i + 1];
                    for (int e = rb + lane; e < re; e += 32) {
                        const uint32_t pos = atoms_add(off_u32 + ((uint32_t)ew[e] + 1u) * 4u, 1u);  // afterwards off[w + 1] = end of w
                        val_s[pos] = ev[e];
                        row_s[pos] = (uint16_t)((t - t0) * SQ_TQ + i);
                    }
                }
            }
        }
        for (int i = threadIdx.x; i < (g_rows + 3) / 4; i += SQ_THREADS)
            reinterpret_cast<volatile uint32_t*>(smem + SQ_SM_COUNT)[i] = 0u;
        for (int i = threadIdx.x; i < (g_rows + 31) / 32; i += SQ_THREADS) lock[i] = 0u;
        named_bar_sync(1, SQ_THREADS);

        // fallback path for one non-zero pool word pv (word id w) of pool row p_rel (relative to the stripe): every
        // intersecting query entry is a hit, queued one per lane and completed 32 at a time by flush_hits
        auto pool_word_slow = [&](uint32_t pv, int w, uint32_t p_rel) {
            const int beg = (int)off[w], end = (int)off[w + 1];
            for (int e0 = beg; e0 < end; e0 += 32) {
                const int e = e0 + lane;
                const bool m = e < end && (val_s[e] & pv) != 0u;
                const uint32_t r = e < end ? (uint32_t)row_s[e] : 0u;
                uint32_t mb = __ballot_sync(0xffffffffu, m);
                while (mb) {
                    const int src = __ffs(mb) - 1;
                    mb &= mb - 1;
                    const uint32_t rr = __shfl_sync(0xffffffffu, r, src);
                    if (lane == qn) {
                        hit_a = rr | ((uint32_t)w << 13);
                        hit_b = p_rel;
                    }
                    if (++qn == 32) {
                        flush_hits(prm, hit_a, hit_b, qn, t0, stripe, pt_beg, lock_u32, count_u32);
                        qn = 0;
                    }
                }
            }
        };
        // replay the words collected so far through the fallback path (the tile turned out too dense for the lists)
        auto replay_slow = [&](int pt) {
#pragma unroll 1
            for (int i = 0; i < pw_n; ++i) {
                const uint32_t pv = lds_u32(pw_a + i * 8), meta = lds_u32(pw_a + i * 8 + 4);
                pool_word_slow(pv, (int)(meta & 0x7ffu), (uint32_t)((pt - pt_beg) * SQ_TP + warp * 8) + (meta >> 11));
            }
            pw_n = 0;
        };

        for (int pt = pt_beg; pt < pt_end; ++pt) {
            bool slow = false;  // this warp's rows of this tile go through the fallback path
            for (int c = 0; c < prm.n_chunks; ++c) {
                mbar_wait(&full_bar[stage], phase);
                const uint32_t sbase = stages_u32 + (uint32_t)stage * SQ_STAGE_BYTES;
                uint4 v0, v1;
                lds128s(v0, sbase + a0);
                lds128s(v1, sbase + a1);
                const uint32_t b0 = __ballot_sync(0xffffffffu, (v0.x | v0.y | v0.z | v0.w) != 0u);
                const uint32_t b1 = __ballot_sync(0xffffffffu, (v1.x | v1.y | v1.z | v1.w) != 0u);
                // the warp's slice of the chunk is in registers: hand the stage back right away
                asm volatile("" ::"r"(b0), "r"(b1) : "memory");  // both ballots (hence every lane's loads) are complete
                if (lane == 0) mbar_arrive(&empty_bar[stage]);
                if (++stage == SQ_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
                if ((b0 | b1) == 0u || prm.debug == 1) continue;
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    uint32_t b = half ? b1 : b0;
                    const uint4 v = half ? v1 : v0;
                    while (b) {
                        const int src = __ffs(b) - 1;
                        b &= b - 1;
                        const uint32_t x0 = __shfl_sync(0xffffffffu, v.x, src), x1 = __shfl_sync(0xffffffffu, v.y, src);
                        const uint32_t x2 = __shfl_sync(0xffffffffu, v.z, src), x3 = __shfl_sync(0xffffffffu, v.w, src);
                        const uint32_t pl = (uint32_t)((src >> 3) + 4 * half);  // pool row inside the warp's 8 rows
                        const int wb = c * SQ_CHUNK_WORDS + (src & 7) * 4;
                        if (!slow && pw_n + 4 > SQ_PW_CAP) {  // word list full: the rest of the tile takes the fallback
                            slow = true;
                            replay_slow(pt);
                        }
                        if (!slow) {  // lanes 0..3 store the unit's non-zero words
                            const uint32_t pv = lane == 0 ? x0 : (lane == 1 ? x1 : (lane == 2 ? x2 : x3));
                            const bool nzw = lane < 4 && pv != 0u;
                            const uint32_t nz = __ballot_sync(0xffffffffu, nzw);
                            if (nzw) {
                                const uint32_t at = pw_a + (uint32_t)(pw_n + __popc(nz & ((1u << lane) - 1u))) * 8u;
                                sts_u32(at, pv);
                                sts_u32(at + 4, (uint32_t)(wb + lane) | (pl << 11));
                            }
                            pw_n += __popc(nz);
                        } else {
                            const uint32_t p_rel = (uint32_t)((pt - pt_beg) * SQ_TP + warp * 8) + pl;
#pragma unroll 1
                            for (int j = 0; j < 4; ++j) {
                                const uint32_t pv = j == 0 ? x0 : (j == 1 ? x1 : (j == 2 ? x2 : x3));
                                if (pv) pool_word_slow(pv, wb + j, p_rel);
                            }
                        }
                    }
                }
            }
            // ---- the tile's lookup phase: every word of the warp's 8 pool rows has been seen
            if (pw_n) {
                __syncwarp();
                if (tile_hits(prm, pw_a, pw_n, hb_a, off_u32, smem_u32(val_s), smem_u32(row_s), t0, stripe,
                              (int64_t)pt * SQ_TP + warp * 8, lock_u32, count_u32))
                    replay_slow(pt);
                pw_n = 0;
            }
        }
        if (qn) {
            flush_hits(prm, hit_a, hit_b, qn, t0, stripe, pt_beg, lock_u32, count_u32);
            qn = 0;
        }

        named_bar_sync(1, SQ_THREADS);  // every hit of this item is stored
        {
            uint8_t* dst = prm.qi.cnt + (size_t)stripe * prm.nq + row0;
            for (int i = threadIdx.x; i < g_rows; i += SQ_THREADS) dst[i] = count[i];
        }
    }
}

// ---------------------------------------------------------------------------- host side
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t sparseq_workspace_bytes(int64_t nq_batch, int32_t n_stripes) {
    const size_t n_qtiles = (size_t)((nq_batch + SQ_TQ - 1) / SQ_TQ);
    return 256 + align256(n_qtiles * 4) * 2 + align256(n_qtiles * SQ_ROWOFF_LD * 2) + align256(n_qtiles * SQ_T1 * 2) +
           align256(n_qtiles * SQ_T1 * 4) + align256((size_t)n_stripes * (size_t)nq_batch);
}

bool sparseq_supported(int32_t words, int32_t k) {
    return options().jaccard_sparse_q != 0 && options().jaccard_skip_zero != 0 && words <= SQ_MAX_WORDS && k <= R4D_TOPK_MAX &&
           k <= 32;
}

QIndex sparseq_carve(void* base, int64_t nq_batch, int32_t n_stripes) {
    const size_t n_qtiles = (size_t)((nq_batch + SQ_TQ - 1) / SQ_TQ);
    uint8_t* p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(base) + 255) & ~uintptr_t(255));
    QIndex qi;
    qi.ent_val = reinterpret_cast<uint32_t*>(p);
    p += align256(n_qtiles * SQ_T1 * 4);
    qi.ent_word = reinterpret_cast<uint16_t*>(p);
    p += align256(n_qtiles * SQ_T1 * 2);
    qi.rowoff = reinterpret_cast<uint16_t*>(p);
    p += align256(n_qtiles * SQ_ROWOFF_LD * 2);
    qi.tile_cnt = reinterpret_cast<uint32_t*>(p);
    p += align256(n_qtiles * 4);
    qi.tile_dense = reinterpret_cast<uint32_t*>(p);
    p += align256(n_qtiles * 4);
    qi.cnt = p;
    (void)n_stripes;
    return qi;
}

int sparseq_build(const uint32_t* qbits, int64_t nq, int32_t words, int32_t pitch_words, int32_t n_stripes,
                  const QIndex& qi, cudaStream_t st) {
    const int n_qtiles = (int)((nq + SQ_TQ - 1) / SQ_TQ);
    R4D_REQUIRE(nq <= SQ_QB, "jaccard query-index path: batch of %lld rows > %d", (long long)nq, SQ_QB);
    R4D_CUDA(cudaMemsetAsync(qi.cnt, 0, (size_t)n_stripes * (size_t)nq, st));
    qindex_kernel<<<n_qtiles, QI_THREADS, 0, st>>>(qbits, nq, words, pitch_words, qi);
    R4D_CUDA(cudaGetLastError());
    return R4D_OK;
}

int sparseq_topk_launch(const uint32_t* pbits, const uint32_t* qcard, const uint32_t* pcard, int64_t nq, int64_t np,
                        int32_t words, int32_t pitch_words, int32_t k, int32_t zero_diag, int64_t query_base,
                        int64_t pool_base, int32_t n_qtiles, int32_t n_ptiles, int32_t n_stripes, int32_t ptiles_per_stripe,
                        uint32_t* part_inter, uint32_t* part_union, int32_t* part_idx, const QIndex& qi, cudaStream_t st) {
    CUtensorMap tm_p;
    int rc = make_tmap_2d(&tm_p, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, pbits, (uint64_t)pitch_words, (uint64_t)np,
                          (uint64_t)pitch_words * 4, SQ_CHUNK_WORDS, SQ_TP, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    SparseParams prm{};
    prm.pbits = pbits;
    prm.qcard = qcard;
    prm.pcard = pcard;
    prm.nq = nq;
    prm.np = np;
    prm.pitch_words = pitch_words;
    prm.n_words = words;
    prm.n_chunks = (words + SQ_CHUNK_WORDS - 1) / SQ_CHUNK_WORDS;
    prm.k = k;
    prm.zero_diag = zero_diag;
    prm.query_base = query_base;
    prm.pool_base = pool_base;
    prm.n_qtiles = n_qtiles;
    prm.n_ptiles = n_ptiles;
    prm.n_stripes = n_stripes;
    prm.ptiles_per_stripe = ptiles_per_stripe;
    prm.part_inter = part_inter;
    prm.part_union = part_union;
    prm.part_idx = part_idx;
    prm.qi = qi;
    prm.debug = options().jaccard_debug;
    static bool attr_done = false;
    if (!attr_done) {
        R4D_CUDA(cudaFuncSetAttribute(jaccard_qindex_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SQ_SM_TOTAL));
        attr_done = true;
    }
    int grid = num_sms();
    if (n_stripes < grid) grid = n_stripes;
    jaccard_qindex_kernel<<<grid, SQ_THREADS + 32, SQ_SM_TOTAL, st>>>(tm_p, prm);
    R4D_CUDA(cudaGetLastError());
    return R4D_OK;
}

}  // namespace r4d
