// jaccard_sparse.cu — fused Jaccard top-K with a QUERY-SIDE BIT INDEX (pool side unchanged: bitset rows streamed from
// HBM by TMA bulk copies, exact integer counts, exact rational ranking).
//
// Why: node-id sets are tiny next to the vocabulary (2.2 of 20 000 bits), so only ~2.4e-4 of all (query, pool) pairs
// intersect at all.  The dense kernel (jaccard.cu) re-streams the pool for every 128-query tile and walks every 8-word
// span of every pair to find that out.  Here the pool is streamed from HBM ONCE per query batch (<= 8 192 rows) and
// the kernel is bound by that stream:
//   1. qindex_kernel turns the query batch into the list of its non-zero words, by row (global memory).
//   2. jaccard_qindex_kernel: every CTA owns a pool stripe.  It expands the index of a whole group of query tiles
//      into shared memory as one entry per SET BIT (query row, bit), sorted by word id (counting sort).
//   3. The 16 warps of a CTA are independent pipelines: warp w owns the 8-row batches w, w+16, ... of the stripe and
//      streams them through its own ring of bulk-copy slots (cp.async.bulk, whole rows, contiguous in HBM, one
//      mbarrier per slot) — no CTA-wide barrier and no shared stage, so a warp that is busy with hits never stalls
//      the feed of the others.  Conflict-free LDS.128 bring a slot into registers, the slot is refilled at once.
//   4. Non-zero pool words of a batch are collected (ballot prefix sums) and looked up: the entries with the same
//      word id whose bit is set in the pool word are hits (query row, pool row).  Every word of the 8 rows has been
//      seen, so the number of hits of a pair IS its intersection; the pair is emitted once.
//   5. Emitted candidates go to the (stripe, query) partial list in global memory: slots are handed out lock-free by
//      a CAS on a per-row counter in shared memory (all lanes in parallel); only a candidate that finds the list
//      full takes the per-row lock and replaces the worst entry if it ranks before it.
// Zero-score candidates are never produced here: they only matter as the lowest-index filler of a short list, which
// the merge kernel adds (jaccard.cu, jaccard_merge_kernel `n_fill`).
//
// Query tiles with more than SQ_T1 non-zero words or SQ_TBITS set bits (dense data, e.g. history sets) are flagged by
// qindex_kernel and handled by the dense kernel in the same launch sequence; both write the same per-stripe partial
// lists.  Pool batches too dense for the per-warp lists fall back to an exact per-hit completion (flush_hits).
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
    __shared__ uint32_t tot_bits;
    const int t = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) tot_bits = 0u;
    __syncthreads();
    for (int i = warp; i < SQ_TQ; i += QI_THREADS / 32) {
        const int64_t gq = (int64_t)t * SQ_TQ + i;
        uint32_t c = 0, bits = 0;
        if (gq < nq) {
            const uint32_t* row = qbits + gq * pitch_words;
            for (int w0 = 0; w0 < words; w0 += 32) {
                const int w = w0 + lane;
                const uint32_t v = w < words ? row[w] : 0u;
                c += __popc(__ballot_sync(0xffffffffu, v != 0u));
                bits += __popc(v);
            }
        }
        bits = __reduce_add_sync(0xffffffffu, bits);
        if (lane == 0) {
            rowcnt[i] = c;
            if (bits) atomicAdd(&tot_bits, bits);
        }
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
    if (total > (uint32_t)SQ_T1 || tot_bits > (uint32_t)SQ_TBITS) {  // too dense for the index path
        if (threadIdx.x == 0) {
            qi.tile_dense[t] = 1u;
            qi.tile_cnt[t] = 0u;
            qi.tile_bits[t] = 0u;
        }
        return;
    }
    if (threadIdx.x == 0) {
        qi.tile_dense[t] = 0u;
        qi.tile_cnt[t] = total;
        qi.tile_bits[t] = tot_bits;
    }
    uint16_t* rowoff = qi.rowoff + (size_t)t * SQ_ROWOFF_LD;
    for (int i = threadIdx.x; i <= SQ_TQ; i += QI_THREADS) rowoff[i] = (uint16_t)rowstart[i];
    uint16_t* ew = qi.ent_word + (size_t)t * SQ_T1;
    uint32_t* ev = qi.ent_val + (size_t)t * SQ_T1;
    uint8_t* er = qi.ent_row + (size_t)t * SQ_T1;
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
                er[pos] = (uint8_t)i;
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
    int32_t pitch_words, n_words, k, zero_diag;
    int64_t query_base, pool_base;
    int32_t n_qtiles, n_stripes;
    int64_t rows_per_stripe;
    int32_t slot_rows;     // pool rows per bulk-copy slot: 8, 4, 2 or 1 (divides SQ_BATCH_ROWS)
    int32_t slot_bytes;    // bytes of a full slot (slot_rows whole pitches, or the live words of one row), 16 B multiple
    int32_t row_units;     // 16-byte units per row inside a slot
    int32_t n_slots;       // slots of a warp's ring (1..SQ_MAX_SLOTS)
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
__device__ __forceinline__ uint32_t atoms_cas(uint32_t addr, uint32_t cmp, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(addr), "r"(cmp), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void reds_add(uint32_t addr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void reds_and(uint32_t addr, uint32_t v) {
    asm volatile("red.shared.and.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_volatile_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
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
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
    asm volatile("{\n\t.reg .u16 t;\n\tcvt.u16.u32 t, %1;\n\tst.shared.u16 [%0], t;\n\t}" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_v2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void lds128s(uint4& v, uint32_t addr) {
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
}
__device__ __forceinline__ uint32_t ldg_volatile_u32(const void* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// TMA bulk copy of `bytes` contiguous bytes (16 B multiple) global -> shared, completion on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// shared-memory addresses (u32) of the candidate bookkeeping of the current item
struct ListState {
    uint32_t lock;   // one bit per group row
    uint32_t alloc;  // u8 per group row: slots handed out (<= k)
    uint32_t pub;    // u8 per group row: slots whose entry is stored (fast path)
};

// Hand the candidates of the lanes with `keep` set to their (stripe, query) partial list in global memory.
// Fast path, all lanes in parallel: a CAS on the row's slot counter hands out an empty slot; the entry is stored and
// then published (counter `pub`).  A candidate that finds all k slots taken goes through the per-row lock: wait until
// the k entries are published, replace the worst one if the candidate ranks before it.
__device__ __noinline__ void append_candidates(const SparseParams& prm, bool keep, uint32_t rr, uint32_t inter, uint32_t uni,
                                               int32_t idx, int t0, int stripe, const ListState ls) {
    const int lane = threadIdx.x & 31;
    const int K = prm.k;
    bool over = false;
    if (keep) {
        const uint32_t wa = (rr & ~3u), sh = (rr & 3u) * 8u;
        uint32_t cur = lds_volatile_u32(ls.alloc + wa), n;
        for (;;) {
            n = (cur >> sh) & 0xffu;
            if (n >= (uint32_t)K) break;
            const uint32_t old = atoms_cas(ls.alloc + wa, cur, cur + (1u << sh));
            if (old == cur) break;
            cur = old;
        }
        if (n < (uint32_t)K) {
            const int64_t at = ((int64_t)stripe * prm.nq + (int64_t)t0 * SQ_TQ + rr) * K + n;
            prm.part_inter[at] = inter;
            prm.part_union[at] = uni;
            prm.part_idx[at] = idx;
            __threadfence_block();  // the entry is visible to the CTA before it counts as published
            reds_add(ls.pub + wa, 1u << sh);
        } else {
            over = true;
        }
    }
    uint32_t pb = __ballot_sync(0xffffffffu, over);
    while (pb) {
        const int src = __ffs(pb) - 1;
        pb &= pb - 1;
        const uint32_t r = __shfl_sync(0xffffffffu, rr, src);
        const JEntry cand{__shfl_sync(0xffffffffu, inter, src), __shfl_sync(0xffffffffu, uni, src),
                          __shfl_sync(0xffffffffu, idx, src)};
        const int64_t base = ((int64_t)stripe * prm.nq + (int64_t)t0 * SQ_TQ + r) * K;
        const uint32_t bit = 1u << (r & 31);
        const uint32_t lock_a = ls.lock + (r >> 5) * 4u;
        if (lane == 0) {
            uint32_t polls = 0;
            while (atoms_or(lock_a, bit) & bit)
                if (++polls > (1u << 28)) __trap();  // a lost unlock must surface as a launch failure, not a hang
            const uint32_t pa = ls.pub + (r & ~3u), sh = (r & 3u) * 8u;
            while (((lds_volatile_u32(pa) >> sh) & 0xffu) < (uint32_t)K)
                if (++polls > (1u << 28)) __trap();
        }
        __syncwarp();
        __threadfence_block();
        if (prm.debug != 5) {  // full list: the candidate replaces the worst entry if it ranks before it
            JEntry wv = lane < K ? JEntry{ldg_volatile_u32(prm.part_inter + base + lane), ldg_volatile_u32(prm.part_union + base + lane),
                                          (int32_t)ldg_volatile_u32(prm.part_idx + base + lane)}
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
        __threadfence_block();  // the replacement is visible to the CTA before the lock is released
        __syncwarp();
        if (lane == 0) reds_and(lock_a, ~bit);
    }
}

// Fallback completion (pool batches too dense for the per-warp lists): up to 32 queued hits, lane i holds hit i
// (a = query row | word << 13 | bit << 24, b = pool row relative to the stripe).  Per lane: the row's entries (by-row
// index, L2) against the pool row give the full intersection and the pair's first intersecting bit; only the hit AT
// that bit is kept, so a pair is emitted exactly once whatever the order the hits arrive in.
__device__ __noinline__ void flush_hits(const SparseParams& prm, uint32_t hit_a, uint32_t hit_b, int qn, int t0, int stripe,
                                        int64_t row_beg, const ListState ls) {
    const int lane = threadIdx.x & 31;
    bool primary = false;
    uint32_t rr = 0, inter = 0, uni = 0;
    int32_t idx = 0;
    if (lane < qn && prm.debug != 2) {
        rr = hit_a & 0x1fffu;
        const uint32_t w = (hit_a >> 13) & 0x7ffu, hb = hit_a >> 24;
        const int64_t gp = row_beg + hit_b;
        const int t = t0 + (int)(rr >> 7), i = (int)(rr & (SQ_TQ - 1));
        const int64_t gq = (int64_t)t0 * SQ_TQ + rr;
        if (!(prm.zero_diag && prm.query_base + gq == prm.pool_base + gp)) {  // the diagonal is a forced zero: a filler
            const uint16_t* ro = prm.qi.rowoff + (size_t)t * SQ_ROWOFF_LD + i;
            const int rb = ro[0], re = ro[1];
            const uint32_t cq = prm.qcard[gq], cp = prm.pcard[gp];
            const uint16_t* ew = prm.qi.ent_word + (size_t)t * SQ_T1;
            const uint32_t* ev = prm.qi.ent_val + (size_t)t * SQ_T1;
            const uint32_t* prow = prm.pbits + gp * prm.pitch_words;
            uint32_t first = 0xffffffffu;  // word << 5 | bit of the first common element
            for (int e = rb; e < re; ++e) {
                const uint32_t ww = ew[e];
                const uint32_t x = ev[e] & __ldg(prow + ww);
                inter += __popc(x);
                if (x) first = min(first, (ww << 5) | (uint32_t)(__ffs(x) - 1));
            }
            primary = first == ((w << 5) | hb) && prm.debug != 3;
            uni = cq + cp - inter;
            idx = (int32_t)(prm.pool_base + gp);
        }
    }
    append_candidates(prm, primary, rr, inter, uni, idx, t0, stripe, ls);
}

// The lookup phase of one batch (8 pool rows of one warp): `pw_n` non-zero pool words (value, word id | row-in-batch
// << 11) wait in shared memory.  Each looks up the index entries with its word id; an entry whose bit is set in the
// pool word is a hit (query row | pool row << 13) and goes to the warp's hit buffer.  Every word of the 8 rows has
// been seen, so the number of hits of a pair IS its intersection; the pair's first hit emits it.
// Returns 1 (nothing emitted) when the hit buffer overflows: the caller replays the batch through flush_hits.
constexpr int SQ_PW_CAP = 64;    // non-zero pool words per warp per batch kept for the lookup phase
constexpr int SQ_HB_CAP = 128;   // hits per warp per batch
__device__ __noinline__ int batch_hits(const SparseParams& prm, uint32_t pw_a, int pw_n, uint32_t hb_a, uint32_t off_a,
                                       uint32_t row_a, uint32_t bit_a, int t0, int stripe, int64_t gp0, uint32_t pc_lane,
                                       const ListState ls) {
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    int hn = 0;
    for (int i = 0; i < pw_n; ++i) {
        const uint32_t pv = lds_u32(pw_a + i * 8), meta = lds_u32(pw_a + i * 8 + 4);
        const uint32_t w = meta & 0x7ffu, pl = meta >> 11;
        const int beg = (int)lds_u32(off_a + w * 4), end = (int)lds_u32(off_a + w * 4 + 4);
        for (int e0 = beg; e0 < end; e0 += 32) {
            const int e = e0 + lane;
            const bool hit = e < end && ((pv >> lds_u8(bit_a + e)) & 1u) != 0u;
            const uint32_t mb = __ballot_sync(0xffffffffu, hit);
            if (!mb) continue;
            const int add = __popc(mb);
            if (hn + add > SQ_HB_CAP) return 1;
            if (hit) sts_u16(hb_a + (hn + __popc(mb & lt)) * 2, lds_u16(row_a + e * 2) | (pl << 13));
            hn += add;
        }
    }
    if (hn == 0 || prm.debug == 2) return 0;
    __syncwarp();
    for (int b = 0; b < hn; b += 32) {
        const int i = b + lane;
        const uint32_t mine = i < hn ? lds_u16(hb_a + i * 2) : 0xffffffffu;
        uint32_t sum = 0;
        bool leader = i < hn;
        for (int j = 0; j < hn; j += 2) {  // two hits per load (the buffer is padded to an even count)
            const uint32_t h2 = lds_u32(hb_a + j * 2);
            const uint32_t ha = h2 & 0xffffu, hc = h2 >> 16;
            if (ha == mine) {
                ++sum;
                if (j < i) leader = false;
            }
            if (hc == mine && j + 1 < hn) {
                ++sum;
                if (j + 1 < i) leader = false;
            }
        }
        const uint32_t rr = mine & 0x1fffu, pl = (mine >> 13) & 7u;
        const int64_t gq = (int64_t)t0 * SQ_TQ + rr, gp = gp0 + pl;
        if (prm.zero_diag && prm.query_base + gq == prm.pool_base + gp) leader = false;  // forced zero: a filler
        if (prm.debug == 3) leader = false;
        const uint32_t cp = __shfl_sync(0xffffffffu, pc_lane, (int)pl);
        uint32_t uni = 0;
        if (leader) uni = prm.qcard[gq] + cp - sum;
        append_candidates(prm, leader, rr, sum, uni, (int32_t)(prm.pool_base + gp), t0, stripe, ls);
    }
    return 0;
}

// shared-memory layout (offsets from the 128-aligned base)
constexpr int SQ_OFF_WORDS = SQ_MAX_WORDS + 5;  // off[0 .. n_words], padded
constexpr size_t SQ_SM_RINGS = 0;                                                  // [16 warps][SQ_RING_BYTES]
constexpr size_t SQ_SM_BARS = SQ_SM_RINGS + (size_t)SQ_WARPS * SQ_RING_BYTES;     // [16 warps][SQ_MAX_SLOTS] mbarriers
constexpr size_t SQ_SM_OFF = SQ_SM_BARS + (size_t)SQ_WARPS * SQ_MAX_SLOTS * 8;
constexpr size_t SQ_SM_ROW = SQ_SM_OFF + (size_t)SQ_OFF_WORDS * 4;
constexpr size_t SQ_SM_BIT = SQ_SM_ROW + (size_t)SQ_E_CAP * 2;
constexpr size_t SQ_SM_ALLOC = SQ_SM_BIT + (size_t)SQ_E_CAP;
constexpr size_t SQ_SM_PUB = SQ_SM_ALLOC + SQ_QB;
constexpr size_t SQ_SM_LOCK = SQ_SM_PUB + SQ_QB;
constexpr size_t SQ_SM_GROUPS = SQ_SM_LOCK + SQ_QB / 8;
constexpr size_t SQ_SM_SCAN = SQ_SM_GROUPS + (size_t)(SQ_MAX_TILES + 2) * 8;
constexpr size_t SQ_SM_PW = SQ_SM_SCAN + 32 * 4;                                   // [16 warps][SQ_PW_CAP] x 8 B
constexpr size_t SQ_SM_HB = SQ_SM_PW + (size_t)SQ_WARPS * SQ_PW_CAP * 8;          // [16 warps][SQ_HB_CAP] x 2 B
constexpr size_t SQ_SM_TOTAL = SQ_SM_HB + (size_t)SQ_WARPS * SQ_HB_CAP * 2 + 128;  // + alignment slack
static_assert(SQ_SM_TOTAL <= 227 * 1024, "query-index kernel: shared memory budget");
static_assert(SQ_SM_OFF % 16 == 0 && SQ_SM_ROW % 4 == 0 && SQ_SM_ALLOC % 4 == 0 && SQ_SM_PW % 8 == 0 && SQ_SM_HB % 4 == 0,
              "query-index kernel: shared memory alignment");

__global__ void __launch_bounds__(SQ_THREADS, 1)
jaccard_qindex_kernel(const __grid_constant__ SparseParams prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SQ_SM_BARS);
    uint32_t* off = reinterpret_cast<uint32_t*>(smem + SQ_SM_OFF);           // entries of word w: [off[w], off[w+1])
    uint16_t* row_s = reinterpret_cast<uint16_t*>(smem + SQ_SM_ROW);         // [E] group-relative query row
    uint8_t* bit_s = smem + SQ_SM_BIT;                                       // [E] bit inside the word
    uint32_t* lock = reinterpret_cast<uint32_t*>(smem + SQ_SM_LOCK);         // one bit per group row
    int32_t* g_first = reinterpret_cast<int32_t*>(smem + SQ_SM_GROUPS);      // [n_groups + 1] first tile of a group
    int32_t* g_ent = g_first + SQ_MAX_TILES + 2;                             // [n_groups] entries of the group
    uint32_t* scan_s = reinterpret_cast<uint32_t*>(smem + SQ_SM_SCAN);       // [16] warp totals, [31] = n_groups

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < SQ_WARPS * SQ_MAX_SLOTS; ++s) mbar_init(&bars[s], 1);
        fence_barrier_init();
        // greedy packing of consecutive query tiles into groups of <= SQ_E_CAP entries (every CTA computes the same)
        int g = 0, sum = 0;
        g_first[0] = 0;
        for (int t = 0; t < prm.n_qtiles; ++t) {
            const int c = (int)prm.qi.tile_bits[t];
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

    const uint32_t off_u32 = smem_u32(off), row_u32 = smem_u32(row_s), bit_u32 = smem_u32(bit_s);
    const ListState ls{smem_u32(lock), smem_u32(smem + SQ_SM_ALLOC), smem_u32(smem + SQ_SM_PUB)};
    const uint32_t ring_u32 = smem_u32(smem + SQ_SM_RINGS) + (uint32_t)warp * SQ_RING_BYTES;   // this warp's ring
    uint64_t* my_bars = bars + warp * SQ_MAX_SLOTS;
    const uint32_t pw_a = smem_u32(smem + SQ_SM_PW) + (uint32_t)warp * SQ_PW_CAP * 8u;  // this warp's pool-word list
    const uint32_t hb_a = smem_u32(smem + SQ_SM_HB) + (uint32_t)warp * SQ_HB_CAP * 2u;  // this warp's hit buffer
    const int R = prm.slot_rows, NS = prm.n_slots;
    const uint32_t upr = (uint32_t)prm.row_units;                        // 16-byte units per row inside a slot
    const uint32_t row_bytes = R == 1 ? (uint32_t)prm.slot_bytes : (uint32_t)prm.pitch_words * 4u;

    int slot = 0;          // ring position of the next slot to consume; slots are issued and consumed cyclically
    uint32_t phase = 0;
    int cur_g = -1;
    uint32_t hit_a = 0, hit_b = 0;  // fallback path: register-resident hit queue, lane i holds hit i
    int qn = 0;
    int pw_n = 0;

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int g = item / prm.n_stripes, stripe = item - g * prm.n_stripes;
        if (g_ent[g] == 0) continue;  // block-uniform: nothing but empty / dense-flagged tiles
        const int t0 = g_first[g], t1 = g_first[g + 1];
        const int64_t row0 = (int64_t)t0 * SQ_TQ;                                  // first query row of the group
        const int g_rows = (int)(min((int64_t)t1 * SQ_TQ, prm.nq) - row0);
        const int64_t r_beg = (int64_t)stripe * prm.rows_per_stripe;               // pool rows [r_beg, r_end)
        const int64_t r_end = min(r_beg + prm.rows_per_stripe, prm.np);
        const int n_rows = (int)(r_end - r_beg);
        const int n_batches = (n_rows + SQ_BATCH_ROWS - 1) / SQ_BATCH_ROWS;

        __syncthreads();  // previous item: all hits handled, counts flushed
        if (g != cur_g) {               // ---- expand the group's entries to one per set bit, sorted by word id
            cur_g = g;
            const int nw = prm.n_words;
            for (int i = threadIdx.x; i <= nw; i += SQ_THREADS) off[i] = 0u;
            __syncthreads();
            for (int t = t0; t < t1; ++t) {
                const int n = (int)prm.qi.tile_cnt[t];
                const uint16_t* ew = prm.qi.ent_word + (size_t)t * SQ_T1;
                const uint32_t* ev = prm.qi.ent_val + (size_t)t * SQ_T1;
                for (int e = threadIdx.x; e < n; e += SQ_THREADS)
                    atoms_add(off_u32 + ((uint32_t)ew[e] + 1u) * 4u, (uint32_t)__popc(ev[e]));
            }
            __syncthreads();
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
                __syncthreads();
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
            __syncthreads();
            for (int t = t0; t < t1; ++t) {
                const int n = (int)prm.qi.tile_cnt[t];  // 0: empty or dense-flagged tile
                const uint16_t* ew = prm.qi.ent_word + (size_t)t * SQ_T1;
                const uint32_t* ev = prm.qi.ent_val + (size_t)t * SQ_T1;
                const uint8_t* er = prm.qi.ent_row + (size_t)t * SQ_T1;
                for (int e = threadIdx.x; e < n; e += SQ_THREADS) {
                    uint32_t v = ev[e];
                    const uint16_t r = (uint16_t)((t - t0) * SQ_TQ + er[e]);
                    // afterwards off[w + 1] = end of word w = start of word w + 1
                    uint32_t pos = atoms_add(off_u32 + ((uint32_t)ew[e] + 1u) * 4u, (uint32_t)__popc(v));
                    while (v) {
                        row_s[pos] = r;
                        bit_s[pos] = (uint8_t)(__ffs(v) - 1);
                        v &= v - 1;
                        ++pos;
                    }
                }
            }
        }
        for (int i = threadIdx.x; i < (g_rows + 3) / 4; i += SQ_THREADS) {
            reinterpret_cast<uint32_t*>(smem + SQ_SM_ALLOC)[i] = 0u;
            reinterpret_cast<uint32_t*>(smem + SQ_SM_PUB)[i] = 0u;
        }
        for (int i = threadIdx.x; i < (g_rows + 31) / 32; i += SQ_THREADS) lock[i] = 0u;
        __syncthreads();

        // ---- this warp's pipeline over the batches warp, warp + 16, ... of the stripe
        // slots of batch b: ceil(rows of b / R); the issue cursor (ib, is) runs n_slots ahead of the consume cursor
        auto batch_slots = [&](int b) { return (min(SQ_BATCH_ROWS, n_rows - b * SQ_BATCH_ROWS) + R - 1) / R; };
        auto issue = [&](int b, int s, int at) {  // lane 0: slot s of batch b into ring position `at`
            const int64_t first = r_beg + (int64_t)b * SQ_BATCH_ROWS + s * R;
            const int nr = (int)min((int64_t)R, r_end - first);
            const uint32_t bytes = R == 1 ? row_bytes : (uint32_t)nr * row_bytes;
            mbar_arrive_expect_tx(&my_bars[at], bytes);
            bulk_load(ring_u32 + (uint32_t)at * (uint32_t)prm.slot_bytes, prm.pbits + first * prm.pitch_words, bytes,
                      smem_u32(&my_bars[at]));
        };
        int ib = warp, is = 0;
        {
            int at = slot;
            for (int j = 0; j < NS && ib < n_batches; ++j) {
                if (lane == 0) issue(ib, is, at);
                if (++at == NS) at = 0;
                if (++is >= batch_slots(ib)) {
                    is = 0;
                    ib += SQ_WARPS;
                }
            }
        }

        // fallback path for one non-zero pool word pv (word id w) of pool row p_rel (relative to the stripe): every
        // entry whose bit is set is a hit, queued one per lane and completed 32 at a time by flush_hits
        auto pool_word_slow = [&](uint32_t pv, int w, uint32_t p_rel) {
            const int beg = (int)off[w], end = (int)off[w + 1];
            for (int e0 = beg; e0 < end; e0 += 32) {
                const int e = e0 + lane;
                const uint32_t eb = e < end ? (uint32_t)bit_s[e] : 0u;
                const bool m = e < end && ((pv >> eb) & 1u) != 0u;
                const uint32_t r = e < end ? (uint32_t)row_s[e] : 0u;
                uint32_t mb = __ballot_sync(0xffffffffu, m);
                while (mb) {
                    const int src = __ffs(mb) - 1;
                    mb &= mb - 1;
                    const uint32_t rr = __shfl_sync(0xffffffffu, r, src);
                    const uint32_t bb = __shfl_sync(0xffffffffu, eb, src);
                    if (lane == qn) {
                        hit_a = rr | ((uint32_t)w << 13) | (bb << 24);
                        hit_b = p_rel;
                    }
                    if (++qn == 32) {
                        flush_hits(prm, hit_a, hit_b, qn, t0, stripe, r_beg, ls);
                        qn = 0;
                    }
                }
            }
        };
        // replay the words collected so far through the fallback path (the batch turned out too dense for the lists)
        auto replay_slow = [&](int b) {
#pragma unroll 1
            for (int i = 0; i < pw_n; ++i) {
                const uint32_t pv = lds_u32(pw_a + i * 8), meta = lds_u32(pw_a + i * 8 + 4);
                pool_word_slow(pv, (int)(meta & 0x7ffu), (uint32_t)(b * SQ_BATCH_ROWS) + (meta >> 11));
            }
            pw_n = 0;
        };

        for (int b = warp; b < n_batches; b += SQ_WARPS) {
            bool slow = false;  // this batch goes through the fallback path
            const int nvs = batch_slots(b);
            const int64_t gp0 = r_beg + (int64_t)b * SQ_BATCH_ROWS;
            // cardinalities of the batch's rows: lane i holds row i (one load, ready long before the lookups)
            const uint32_t pc_lane = (lane < SQ_BATCH_ROWS && gp0 + lane < r_end) ? __ldg(prm.pcard + gp0 + lane) : 0u;
            for (int s = 0; s < nvs; ++s) {
                const int nr = min(R, n_rows - b * SQ_BATCH_ROWS - s * R);   // rows of this slot
                const uint32_t n_units = (uint32_t)nr * upr;
                const uint32_t sbase = ring_u32 + (uint32_t)slot * (uint32_t)prm.slot_bytes;
                mbar_wait(&my_bars[slot], phase);
#pragma unroll 1
                for (uint32_t u0 = 0; u0 < n_units; u0 += 6 * 32) {
                    uint4 v[6];
                    uint32_t any[6];
#pragma unroll
                    for (int j = 0; j < 6; ++j) {
                        const uint32_t u = u0 + j * 32 + lane;
                        v[j] = make_uint4(0u, 0u, 0u, 0u);
                        if (u < n_units) lds128s(v[j], sbase + u * 16u);
                    }
#pragma unroll
                    for (int j = 0; j < 6; ++j) any[j] = __ballot_sync(0xffffffffu, (v[j].x | v[j].y | v[j].z | v[j].w) != 0u);
                    if (u0 + 6 * 32 >= n_units) {
                        // the whole slot is in registers (the ballots above needed every lane's loads): refill it
                        asm volatile("" ::"r"(any[0]), "r"(any[1]), "r"(any[2]), "r"(any[3]), "r"(any[4]), "r"(any[5]) : "memory");
                        if (ib < n_batches) {
                            if (lane == 0) issue(ib, is, slot);
                            if (++is >= batch_slots(ib)) {
                                is = 0;
                                ib += SQ_WARPS;
                            }
                        }
                    }
                    if (prm.debug == 1) continue;
#pragma unroll
                    for (int j = 0; j < 6; ++j) {
                        if (any[j] == 0u) continue;
                        const uint32_t u = u0 + j * 32 + lane;
                        uint32_t rs = 0, uw = u;   // row inside the slot, unit inside the row
                        if (R != 1) {
                            rs = u / upr;
                            uw = u - rs * upr;
                        }
                        const uint32_t pl = (uint32_t)(s * R) + rs;   // row inside the batch
                        const uint32_t meta = (uw * 4u) | (pl << 11);
                        if (uw * 4u + 4u > (uint32_t)prm.n_words) {   // words past the vocabulary are padding, whatever they hold
                            const uint32_t live = uw * 4u < (uint32_t)prm.n_words ? (uint32_t)prm.n_words - uw * 4u : 0u;
                            if (live < 1u) v[j].x = 0u;
                            if (live < 2u) v[j].y = 0u;
                            if (live < 3u) v[j].z = 0u;
                            v[j].w = 0u;
                        }
                        const uint32_t n0 = __ballot_sync(0xffffffffu, v[j].x != 0u), n1 = __ballot_sync(0xffffffffu, v[j].y != 0u);
                        const uint32_t n2 = __ballot_sync(0xffffffffu, v[j].z != 0u), n3 = __ballot_sync(0xffffffffu, v[j].w != 0u);
                        const int c0 = __popc(n0), c1 = c0 + __popc(n1), c2 = c1 + __popc(n2), c3 = c2 + __popc(n3);
                        if (!slow && pw_n + c3 > SQ_PW_CAP) {  // word list full: the rest of the batch takes the fallback
                            slow = true;
                            replay_slow(b);
                        }
                        if (!slow) {
                            if (v[j].x) sts_v2(pw_a + (uint32_t)(pw_n + __popc(n0 & lt)) * 8u, v[j].x, meta);
                            if (v[j].y) sts_v2(pw_a + (uint32_t)(pw_n + c0 + __popc(n1 & lt)) * 8u, v[j].y, meta + 1u);
                            if (v[j].z) sts_v2(pw_a + (uint32_t)(pw_n + c1 + __popc(n2 & lt)) * 8u, v[j].z, meta + 2u);
                            if (v[j].w) sts_v2(pw_a + (uint32_t)(pw_n + c2 + __popc(n3 & lt)) * 8u, v[j].w, meta + 3u);
                            pw_n += c3;
                        } else {
                            uint32_t bm = any[j];
#pragma unroll 1
                            while (bm) {
                                const int src = __ffs(bm) - 1;
                                bm &= bm - 1;
                                const uint32_t x0 = __shfl_sync(0xffffffffu, v[j].x, src), x1 = __shfl_sync(0xffffffffu, v[j].y, src);
                                const uint32_t x2 = __shfl_sync(0xffffffffu, v[j].z, src), x3 = __shfl_sync(0xffffffffu, v[j].w, src);
                                const uint32_t ms = __shfl_sync(0xffffffffu, meta, src);
                                const uint32_t p_rel = (uint32_t)(b * SQ_BATCH_ROWS) + (ms >> 11);
#pragma unroll 1
                                for (int q = 0; q < 4; ++q) {
                                    const uint32_t pv = q == 0 ? x0 : (q == 1 ? x1 : (q == 2 ? x2 : x3));
                                    if (pv) pool_word_slow(pv, (int)(ms & 0x7ffu) + q, p_rel);
                                }
                            }
                        }
                    }
                }
                if (++slot == NS) {
                    slot = 0;
                    phase ^= 1;
                }
            }
            // ---- the batch's lookup phase: every word of the warp's 8 pool rows has been seen
            if (pw_n) {
                __syncwarp();
                if (batch_hits(prm, pw_a, pw_n, hb_a, off_u32, row_u32, bit_u32, t0, stripe, gp0, pc_lane, ls)) replay_slow(b);
                pw_n = 0;
            }
        }
        if (qn) {
            flush_hits(prm, hit_a, hit_b, qn, t0, stripe, r_beg, ls);
            qn = 0;
        }

        __syncthreads();  // every hit of this item is stored
        {
            uint8_t* dst = prm.qi.cnt + (size_t)stripe * prm.nq + row0;
            const uint8_t* alloc = smem + SQ_SM_ALLOC;
            for (int i = threadIdx.x; i < g_rows; i += SQ_THREADS) dst[i] = alloc[i];
        }
    }
}

// ---------------------------------------------------------------------------- host side
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t sparseq_workspace_bytes(int64_t nq_batch, int32_t n_stripes) {
    const size_t n_qtiles = (size_t)((nq_batch + SQ_TQ - 1) / SQ_TQ);
    return 256 + align256(n_qtiles * 4) * 3 + align256(n_qtiles * SQ_ROWOFF_LD * 2) + align256(n_qtiles * SQ_T1 * 2) +
           align256(n_qtiles * SQ_T1 * 4) + align256(n_qtiles * SQ_T1) + align256((size_t)n_stripes * (size_t)nq_batch);
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
    qi.ent_row = p;
    p += align256(n_qtiles * SQ_T1);
    qi.rowoff = reinterpret_cast<uint16_t*>(p);
    p += align256(n_qtiles * SQ_ROWOFF_LD * 2);
    qi.tile_cnt = reinterpret_cast<uint32_t*>(p);
    p += align256(n_qtiles * 4);
    qi.tile_bits = reinterpret_cast<uint32_t*>(p);
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
    R4D_REQUIRE((reinterpret_cast<uintptr_t>(pbits) & 15) == 0 && pitch_words % 4 == 0,
                "jaccard query-index path: pool bitsets must be 16-byte aligned (base %p, pitch %d words)", (const void*)pbits,
                pitch_words);
    (void)n_ptiles;
    SparseParams prm{};
    prm.pbits = pbits;
    prm.qcard = qcard;
    prm.pcard = pcard;
    prm.nq = nq;
    prm.np = np;
    prm.pitch_words = pitch_words;
    prm.n_words = words;
    prm.k = k;
    prm.zero_diag = zero_diag;
    prm.query_base = query_base;
    prm.pool_base = pool_base;
    prm.n_qtiles = n_qtiles;
    prm.n_stripes = n_stripes;
    prm.rows_per_stripe = (int64_t)ptiles_per_stripe * SQ_TP;
    // slot geometry: several whole pitches per bulk copy while they fit a third of the ring, else the live words of
    // one row (the zero padding of the pitch is not read)
    const int64_t pitch_bytes = (int64_t)pitch_words * 4;
    int R = SQ_BATCH_ROWS;
    while (R > 1 && R * pitch_bytes > SQ_RING_BYTES / 3) R >>= 1;
    if (R > 1) {
        prm.slot_rows = R;
        prm.slot_bytes = (int32_t)(R * pitch_bytes);
        prm.row_units = pitch_words / 4;
    } else {
        prm.slot_rows = 1;
        prm.slot_bytes = ((words + 3) / 4) * 16;
        prm.row_units = prm.slot_bytes / 16;
    }
    prm.n_slots = SQ_RING_BYTES / prm.slot_bytes;
    if (prm.n_slots > SQ_MAX_SLOTS) prm.n_slots = SQ_MAX_SLOTS;
    R4D_REQUIRE(prm.n_slots >= 1, "jaccard query-index path: a row of %d words does not fit the ring", words);
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
    jaccard_qindex_kernel<<<grid, SQ_THREADS, SQ_SM_TOTAL, st>>>(prm);
    R4D_CUDA(cudaGetLastError());
    return R4D_OK;
}

}  // namespace r4d
