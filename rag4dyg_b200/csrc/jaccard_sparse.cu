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
//   5. Emitted candidates go to the query's own contiguous array in global memory (slot = one global atomicAdd, all
//      lanes in parallel; the merge kernel reads it coalesced).  Only when that array is full (> 1 024 candidates of a
//      query in one batch) do they go to the (stripe, query) list of the CTA: slots handed out lock-free by a CAS on a
//      per-row counter in shared memory, and a candidate that finds that list full takes the per-row lock and
//      replaces the worst entry if it ranks before it.
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
// grid = n_qtiles CTAs of 1024 threads (32 warps, 4 rows each).  One pass over the tile's bitsets: the non-zero words
// are staged in shared memory in arrival order, then written out grouped by row (the order inside a row is arbitrary;
// every consumer is order independent).
constexpr int QI_THREADS = 1024;
__global__ void __launch_bounds__(QI_THREADS)
qindex_kernel(const uint32_t* __restrict__ qbits, int64_t nq, int32_t words, int32_t pitch_words, QIndex qi) {
    __shared__ uint32_t rowcnt[SQ_TQ];
    __shared__ uint32_t rowstart[SQ_TQ + 1];
    __shared__ uint32_t rowfill[SQ_TQ];
    __shared__ uint32_t st_val[SQ_T1];
    __shared__ uint16_t st_word[SQ_T1];
    __shared__ uint8_t st_row[SQ_T1];
    __shared__ uint32_t tot_bits, n_ent;
    const int t = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tot_bits = 0u;
        n_ent = 0u;
    }
    if (threadIdx.x < SQ_TQ) rowfill[threadIdx.x] = 0u;
    __syncthreads();
    for (int i = warp; i < SQ_TQ; i += QI_THREADS / 32) {
        const int64_t gq = (int64_t)t * SQ_TQ + i;
        uint32_t c = 0, bits = 0;
        if (gq < nq) {
            const uint32_t* row = qbits + gq * pitch_words;
            for (int w0 = 0; w0 < words; w0 += 8 * 32) {   // eight independent loads in flight per lane
                uint32_t v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int w = w0 + j * 32 + lane;
                    v[j] = w < words ? __ldg(row + w) : 0u;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t b = __ballot_sync(0xffffffffu, v[j] != 0u);
                    if (b == 0u) continue;
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(&n_ent, (uint32_t)__popc(b));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (v[j] != 0u) {
                        const uint32_t pos = base + __popc(b & ((1u << lane) - 1u));
                        if (pos < (uint32_t)SQ_T1) {
                            st_val[pos] = v[j];
                            st_word[pos] = (uint16_t)(w0 + j * 32 + lane);
                            st_row[pos] = (uint8_t)i;
                        }
                    }
                    c += __popc(b);
                    bits += __popc(v[j]);
                }
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
            atomicOr(qi.any_dense + t / SQ_MAX_TILES, 1u);   // one flag per 8 192-row batch
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
    for (uint32_t e = threadIdx.x; e < total; e += QI_THREADS) {
        const uint32_t r = st_row[e];
        const uint32_t pos = rowstart[r] + atomicAdd(&rowfill[r], 1u);
        ew[pos] = st_word[e];
        ev[pos] = st_val[e];
        er[pos] = (uint8_t)r;
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
    int32_t slot_rows_log2;
    int32_t bucket_shift;  // index bucket of a bit id = id >> bucket_shift (0 while the vocabulary fits SQ_NBK buckets)
    int32_t n_buckets;
    uint4* part;           // [n_stripes][nq][k] {inter, |pool set|, idx, 0}: the merge kernel forms union = |q| + |p| - inter
    QIndex qi;
    int32_t debug;  // stage bypass for measurements (tools/probe_qindex.py; results are wrong unless 0): 1 = scan only,
                    // 2 = lookups but no hit completion, 3 = no candidate is emitted, 5 = full lists are never updated,
                    // 6 = no fence before a per-stripe entry is published, 7 = candidate stores skipped
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
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
    const uint16_t h = (uint16_t)v;
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(h) : "memory");
}
__device__ __forceinline__ void lds128s(uint4& v, uint32_t addr) {
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
}
__device__ __forceinline__ uint4 ldg_volatile_v4(const void* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// TMA bulk copy of `bytes` contiguous bytes (16 B multiple) global -> shared, completion on an mbarrier (SASS: UBLKCP).
// `policy`: an L2 cache policy (createpolicy); the pool is read once per launch, so its lines are marked evict-first
// and the candidate lists written meanwhile stay in L2 for the merge kernel.
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// shared-memory addresses (u32) of the candidate bookkeeping of the current item
struct ListState {
    uint32_t lock;   // one bit per group row
    uint32_t alloc;  // u8 per group row: slots handed out (<= k)
    uint32_t pub;    // u8 per group row: slots whose entry is stored (fast path)
};

// Store the candidates of the lanes with `keep` set, all lanes in parallel.
//  1. The query's own contiguous array glist[q][SQ_GC]: a global atomicAdd on gcount[q] hands out the slot.
//  2. Array full: the (stripe, query) list of this CTA (k slots).  A CAS on the row's slot counter in shared memory
//     hands out an empty slot; the entry is stored and then published (counter `pub`).
//  3. That list full as well: the per-row lock — wait until its k entries are published, replace the worst one if the
//     candidate ranks before it.
// top-k(all candidates) is contained in glist[q] + the per-stripe top-k of the rest, so the merge stays exact.
__device__ __noinline__ void append_candidates(const SparseParams& prm, bool keep, uint32_t rr, uint32_t inter, uint32_t pcard,
                                               int32_t idx, int t0, int stripe, const ListState ls) {
    const int lane = threadIdx.x & 31;
    const int K = prm.k;
    bool over = false;
    if (keep) {   // first choice: the next free slot of the query's own contiguous array (one global atomic)
        const int64_t gq = (int64_t)t0 * SQ_TQ + rr;
        const uint32_t slot = atomicAdd(prm.qi.gcount + gq, 1u);
        if (slot < (uint32_t)SQ_GC) {
            if (prm.debug != 7) prm.qi.glist[gq * SQ_GC + slot] = make_uint4(inter, pcard, (uint32_t)idx, 0u);
            keep = false;
        }
    }
    if (keep) {   // the array is full: this stripe's k-entry list of the query
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
            const int64_t at = (((int64_t)t0 * SQ_TQ + rr) * prm.n_stripes + stripe) * K + n;   // query-major lists
            if (prm.debug != 7) {
                prm.part[at] = make_uint4(inter, pcard, (uint32_t)idx, 0u);
            }
            if (prm.debug != 6) __threadfence_block();  // the entry is visible to the CTA before it counts as published
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
        const uint32_t cq = prm.qcard[(int64_t)t0 * SQ_TQ + r];   // entries hold |pool set|: union = |q| + |p| - inter
        const uint32_t c_inter = __shfl_sync(0xffffffffu, inter, src), c_pcard = __shfl_sync(0xffffffffu, pcard, src);
        const JEntry cand{c_inter, cq + c_pcard - c_inter, __shfl_sync(0xffffffffu, idx, src)};
        const int64_t base = (((int64_t)t0 * SQ_TQ + r) * prm.n_stripes + stripe) * K;
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
            JEntry wv{0xffffffffu, 1u, -1};  // ranks before every real entry
            if (lane < K) {
                const uint4 e = ldg_volatile_v4(prm.part + base + lane);
                wv = JEntry{e.x, cq + e.y - e.x, (int32_t)e.z};
            }
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
                prm.part[base + wl] = make_uint4(c_inter, c_pcard, (uint32_t)cand.idx, 0u);
            }
        }
        __threadfence_block();  // the replacement is visible to the CTA before the lock is released
        __syncwarp();
        if (lane == 0) reds_and(lock_a, ~bit);
    }
}

// Fallback completion (pool batches too dense for the per-warp lists): up to 32 queued hits, lane i holds hit i
// (a = query row | bit id << 13, b = pool row relative to the stripe).  Per lane: the row's entries (by-row index, L2)
// against the pool row give the full intersection and the pair's first common bit; only the hit AT that bit is kept,
// so a pair is emitted exactly once whatever the order the hits arrive in.
__device__ __noinline__ void flush_hits(const SparseParams& prm, uint32_t hit_a, uint32_t hit_b, int qn, int t0, int stripe,
                                        int64_t row_beg, const ListState ls) {
    const int lane = threadIdx.x & 31;
    bool primary = false;
    uint32_t rr = 0, inter = 0, uni = 0;
    int32_t idx = 0;
    if (lane < qn && prm.debug != 2) {
        rr = hit_a & 0x1fffu;
        const uint32_t hbit = hit_a >> 13;
        const int64_t gp = row_beg + hit_b;
        const int t = t0 + (int)(rr >> 7), i = (int)(rr & (SQ_TQ - 1));
        const int64_t gq = (int64_t)t0 * SQ_TQ + rr;
        if (!(prm.zero_diag && prm.query_base + gq == prm.pool_base + gp)) {  // the diagonal is a forced zero: a filler
            const uint16_t* ro = prm.qi.rowoff + (size_t)t * SQ_ROWOFF_LD + i;
            const int rb = ro[0], re = ro[1];
            const uint32_t cp = prm.pcard[gp];
            const uint16_t* ew = prm.qi.ent_word + (size_t)t * SQ_T1;
            const uint32_t* ev = prm.qi.ent_val + (size_t)t * SQ_T1;
            const uint32_t* prow = prm.pbits + gp * prm.pitch_words;
            uint32_t first = 0xffffffffu;  // bit id of the first common element
            for (int e = rb; e < re; ++e) {
                const uint32_t ww = ew[e];
                const uint32_t x = ev[e] & __ldg(prow + ww);
                inter += __popc(x);
                if (x) first = min(first, (ww << 5) | (uint32_t)(__ffs(x) - 1));
            }
            primary = first == hbit && prm.debug != 3;
            uni = cp;   // the list entry holds |pool set|
            idx = (int32_t)(prm.pool_base + gp);
        }
    }
    append_candidates(prm, primary, rr, inter, uni, idx, t0, stripe, ls);
}

// The lookup phase of one batch (8 pool rows of one warp): `nb` set pool bits (bit id | row-in-batch << 16) wait in
// shared memory.  Lane i looks bit i up: the entries of its bucket whose low bits match are hits
// (query row | pool row << 13) and go to the warp's hit buffer.  Every bit of the 8 rows has been seen, so the
// number of hits of a pair IS its intersection; the pair's first hit emits it.
// Returns 1 (nothing emitted) when the hit buffer overflows: the caller replays the batch through flush_hits.
constexpr int SQ_BL_CAP = 64;    // set pool bits per warp per batch kept for the lookup phase
constexpr int SQ_HB_CAP = 128;   // hits per warp per batch
constexpr int SQ_LANE_ENT = 8;   // bucket entries a lane walks on its own; longer buckets are walked by the whole warp
__device__ __noinline__ int batch_hits(const SparseParams& prm, uint32_t bl_a, int nb, uint32_t hb_a, uint32_t off_a,
                                       uint32_t row_a, int t0, int stripe, int64_t gp0, uint32_t pc_lane,
                                       const ListState ls) {
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t sh = (uint32_t)prm.bucket_shift, lomask = (1u << sh) - 1u;
    int hn = 0;
    for (int i0 = 0; i0 < nb; i0 += 32) {
        const int i = i0 + lane;
        const bool have = i < nb;
        const uint32_t ent = have ? lds_u32(bl_a + i * 4) : 0u;
        const uint32_t bitid = ent & 0xffffu, pl = ent >> 16;
        const uint32_t bucket = bitid >> sh, lo = bitid & lomask;
        const int beg = have ? (int)lds_u16(off_a + bucket * 2) : 0;
        const int end = have ? (int)lds_u16(off_a + bucket * 2 + 2) : 0;
        const int n_it = min(__reduce_max_sync(0xffffffffu, end - beg), SQ_LANE_ENT);
        for (int kk = 0; kk < n_it; ++kk) {
            const int e = beg + kk;
            const uint32_t x = e < end ? lds_u16(row_a + e * 2) : 0u;
            const bool hit = e < end && (x & 7u) == lo;
            const uint32_t mb = __ballot_sync(0xffffffffu, hit);
            if (!mb) continue;
            const int add = __popc(mb);
            if (hn + add > SQ_HB_CAP) return 1;
            if (hit) sts_u16(hb_a + (hn + __popc(mb & lt)) * 2, (x >> 3) | (pl << 13));
            hn += add;
        }
        uint32_t big = __ballot_sync(0xffffffffu, end - beg > SQ_LANE_ENT);   // long buckets: 32 entries at a time
        while (big) {
            const int src = __ffs(big) - 1;
            big &= big - 1;
            const int b2 = __shfl_sync(0xffffffffu, beg, src) + SQ_LANE_ENT, e2 = __shfl_sync(0xffffffffu, end, src);
            const uint32_t lo2 = __shfl_sync(0xffffffffu, lo, src), pl2 = __shfl_sync(0xffffffffu, pl, src);
            for (int e0 = b2; e0 < e2; e0 += 32) {
                const int e = e0 + lane;
                const uint32_t x = e < e2 ? lds_u16(row_a + e * 2) : 0u;
                const bool hit = e < e2 && (x & 7u) == lo2;
                const uint32_t mb = __ballot_sync(0xffffffffu, hit);
                if (!mb) continue;
                const int add = __popc(mb);
                if (hn + add > SQ_HB_CAP) return 1;
                if (hit) sts_u16(hb_a + (hn + __popc(mb & lt)) * 2, (x >> 3) | (pl2 << 13));
                hn += add;
            }
        }
    }
    if (hn == 0 || prm.debug == 2) return 0;
    __syncwarp();
    for (int b = 0; b < hn; b += 32) {
        const int i = b + lane;
        const uint32_t mine = i < hn ? lds_u16(hb_a + i * 2) : 0xffffffffu;
        uint32_t sum = 0;
        bool leader = i < hn;
        for (int j = 0; j < hn; j += 2) {  // two hits per load (entries past hn are ignored)
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
        const uint32_t cp = __shfl_sync(0xffffffffu, pc_lane, (int)pl);   // |pool set|; the union is formed at the merge
        append_candidates(prm, leader, rr, sum, cp, (int32_t)(prm.pool_base + gp), t0, stripe, ls);
    }
    return 0;
}

// shared-memory layout (offsets from the 128-aligned base)
constexpr size_t SQ_SM_RINGS = 0;                                                  // [16 warps][SQ_RING_BYTES]
constexpr size_t SQ_SM_BARS = SQ_SM_RINGS + (size_t)SQ_WARPS * SQ_RING_BYTES;     // [16 warps][SQ_MAX_SLOTS] mbarriers
constexpr size_t SQ_SM_OFF = SQ_SM_BARS + (size_t)SQ_WARPS * SQ_MAX_SLOTS * 8;    // u16 off[SQ_NBK + 2], padded
constexpr size_t SQ_SM_ROW = SQ_SM_OFF + (((size_t)SQ_NBK + 2) * 2 + 15) / 16 * 16;
constexpr size_t SQ_SM_ALLOC = SQ_SM_ROW + (size_t)SQ_E_CAP * 2;
constexpr size_t SQ_SM_PUB = SQ_SM_ALLOC + SQ_QB;
constexpr size_t SQ_SM_LOCK = SQ_SM_PUB + SQ_QB;
constexpr size_t SQ_SM_GROUPS = SQ_SM_LOCK + SQ_QB / 8;
constexpr size_t SQ_SM_TCNT = SQ_SM_GROUPS + (size_t)(SQ_MAX_TILES + 2) * 8;    // [SQ_MAX_TILES] u16 entries of a tile
constexpr size_t SQ_SM_TBITS = SQ_SM_TCNT + (size_t)SQ_MAX_TILES * 2;           // [SQ_MAX_TILES] u16 set bits of a tile
constexpr size_t SQ_SM_SCAN = SQ_SM_TBITS + (size_t)SQ_MAX_TILES * 2;
// per-warp scratch of 512 B: while a batch is scanned it holds the batch's non-zero words (values [80] u32, then tags
// [80] u16 = row in batch << 11 | word id); the lookup phase takes them into registers and reuses the space for the
// bit list ([SQ_BL_CAP] x 4 B) and the hit buffer ([SQ_HB_CAP] x 2 B)
constexpr int SQ_IB = 9;         // index build: entries a lane loads per round (9 * 32 = 288 covers a typical tile)
constexpr int SQ_UL_CAP = 80;
constexpr int SQ_SCRATCH = 512;
constexpr size_t SQ_SM_SCR = (SQ_SM_SCAN + 32 * 4 + 15) / 16 * 16;
constexpr size_t SQ_SM_CNT = SQ_SM_SCR + (size_t)SQ_WARPS * SQ_SCRATCH;             // [16 warps] {units, bits} counters
constexpr size_t SQ_SM_TOTAL = SQ_SM_CNT + (size_t)SQ_WARPS * 8 + 128;              // + alignment slack
static_assert(SQ_UL_CAP * 6 <= SQ_SCRATCH && SQ_UL_CAP <= 96 && SQ_BL_CAP * 4 + SQ_HB_CAP * 2 <= SQ_SCRATCH, "per-warp scratch");
static_assert(SQ_SM_TOTAL <= 227 * 1024, "query-index kernel: shared memory budget");
static_assert(SQ_SM_OFF % 16 == 0 && SQ_SM_ROW % 4 == 0 && SQ_SM_ALLOC % 4 == 0 && SQ_SM_SCR % 16 == 0 && SQ_SM_CNT % 8 == 0,
              "query-index kernel: shared memory alignment");
static_assert(SQ_E_CAP < 65536, "bucket offsets are 16-bit");

// NU > 0 ("ROW1"): one pool row per slot and at most NU * 32 16-byte units per row (vocabularies up to 24 K bits): the
// slot geometry is a compile-time shape and every lane's shared-memory offsets are immediates.  NU == 0: any geometry.
template <int NU>
__global__ void __launch_bounds__(SQ_THREADS, 1)
jaccard_qindex_kernel(const __grid_constant__ SparseParams prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SQ_SM_BARS);
    uint16_t* off = reinterpret_cast<uint16_t*>(smem + SQ_SM_OFF);           // entries of bucket b: [off[b], off[b+1])
    uint16_t* row_s = reinterpret_cast<uint16_t*>(smem + SQ_SM_ROW);         // [E] group-relative query row << 3 | low bits
    uint32_t* lock = reinterpret_cast<uint32_t*>(smem + SQ_SM_LOCK);         // one bit per group row
    int32_t* g_first = reinterpret_cast<int32_t*>(smem + SQ_SM_GROUPS);      // [n_groups + 1] first tile of a group
    int32_t* g_ent = g_first + SQ_MAX_TILES + 2;                             // [n_groups] entries of the group
    uint32_t* scan_s = reinterpret_cast<uint32_t*>(smem + SQ_SM_SCAN);       // [16] warp totals, [31] = n_groups
    uint16_t* t_cnt = reinterpret_cast<uint16_t*>(smem + SQ_SM_TCNT);        // [n_qtiles] word entries of every tile

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr bool ROW1 = NU > 0;
    constexpr int NJ = ROW1 ? NU : 6;   // 16-byte units a lane loads per pass over a slot

    uint16_t* t_bits = reinterpret_cast<uint16_t*>(smem + SQ_SM_TBITS);
    for (int t = threadIdx.x; t < prm.n_qtiles; t += SQ_THREADS) {   // one round of parallel loads, not a serial chain
        t_cnt[t] = (uint16_t)prm.qi.tile_cnt[t];
        t_bits[t] = (uint16_t)prm.qi.tile_bits[t];
    }
    if (threadIdx.x < SQ_WARPS * SQ_MAX_SLOTS) {
        mbar_init(&bars[threadIdx.x], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // greedy packing of consecutive query tiles into groups of <= SQ_E_CAP entries (every CTA computes the same)
        int g = 0, sum = 0;
        g_first[0] = 0;
        for (int t = 0; t < prm.n_qtiles; ++t) {
            const int c = (int)t_bits[t];
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

    const uint32_t off_u32 = smem_u32(off), row_u32 = smem_u32(row_s);
    const ListState ls{smem_u32(lock), smem_u32(smem + SQ_SM_ALLOC), smem_u32(smem + SQ_SM_PUB)};
    const uint32_t ring_u32 = smem_u32(smem + SQ_SM_RINGS) + (uint32_t)warp * SQ_RING_BYTES;   // this warp's ring
    uint64_t* my_bars = bars + warp * SQ_MAX_SLOTS;
    const uint32_t ul_a = smem_u32(smem + SQ_SM_SCR) + (uint32_t)warp * SQ_SCRATCH;     // this warp's word list ...
    const uint32_t ut_a = ul_a + SQ_UL_CAP * 4u;                                        // ... and its tags
    const uint32_t bl_a = ul_a;                                                          // pool-bit list (lookup phase)
    const uint32_t hb_a = ul_a + SQ_BL_CAP * 4u;                                        // hit buffer (lookup phase)
    const uint32_t uln_a = smem_u32(smem + SQ_SM_CNT) + (uint32_t)warp * 8u;            // units in the list
    const uint32_t bln_a = uln_a + 4u;                                                  // bits in the list
    const int R = ROW1 ? 1 : prm.slot_rows, lgR = ROW1 ? 0 : prm.slot_rows_log2, NS = prm.n_slots;
    const uint32_t upr = (uint32_t)prm.row_units;                        // 16-byte units per row inside a slot
    const uint32_t row_bytes = R == 1 ? (uint32_t)prm.slot_bytes : (uint32_t)prm.pitch_words * 4u;
    const uint32_t bsh = (uint32_t)prm.bucket_shift, blo = (1u << bsh) - 1u;

    const uint64_t l2_stream = l2_policy_evict_first();
    int slot = 0;          // ring position of the next slot to consume; slots are issued and consumed cyclically
    uint32_t phase = 0;
    int cur_g = -1;
    uint32_t hit_a = 0, hit_b = 0;  // fallback path: register-resident hit queue, lane i holds hit i
    int qn = 0;
    if (lane == 0) {
        sts_u32(uln_a, 0u);
        sts_u32(bln_a, 0u);
    }
    // every slot is followed by 16 zero bytes: lanes past the end of a slot read those instead of branching
    const uint32_t slot_stride = (uint32_t)prm.slot_bytes + 16u;
    if (lane < NS) {
        const uint32_t za = ring_u32 + (uint32_t)lane * slot_stride + (uint32_t)prm.slot_bytes;
        sts_u32(za, 0u);
        sts_u32(za + 4, 0u);
        sts_u32(za + 8, 0u);
        sts_u32(za + 12, 0u);
    }
    __syncwarp();
    const uint32_t full_units = (uint32_t)R * upr;   // units of a full slot
    uint32_t loff[NJ];                                // this lane's byte offsets inside a full slot (first 192 units)
#pragma unroll
    for (int j = 0; j < NJ; ++j) loff[j] = (uint32_t)(j * 32 + lane) < full_units ? (uint32_t)(j * 32 + lane) * 16u : (uint32_t)prm.slot_bytes;

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
        if (g != cur_g) {  // ---- expand the group's entries to one per set bit, sorted by bucket (counting sort)
            cur_g = g;
            const int nbk = prm.n_buckets;                      // off[0 .. nbk] are used
            for (int i = threadIdx.x; i < (nbk + 2) / 2 + 1; i += SQ_THREADS) reinterpret_cast<uint32_t*>(off)[i] = 0u;
            __syncthreads();
            // warps take the group's tiles round-robin; nine entries per lane (a typical tile in one round) are loaded
            // before they are used
            for (int t = t0 + warp; t < t1; t += SQ_WARPS) {
                const int n = (int)t_cnt[t];
                const uint16_t* ew = prm.qi.ent_word + (size_t)t * SQ_T1;
                const uint32_t* ev = prm.qi.ent_val + (size_t)t * SQ_T1;
                for (int e0 = lane; e0 < n; e0 += SQ_IB * 32) {
                    uint32_t v[SQ_IB], wb[SQ_IB];
#pragma unroll
                    for (int j = 0; j < SQ_IB; ++j) {
                        const int e = e0 + j * 32;
                        v[j] = e < n ? ev[e] : 0u;
                        wb[j] = e < n ? (uint32_t)ew[e] << 5 : 0u;
                    }
#pragma unroll
                    for (int j = 0; j < SQ_IB; ++j)
                        while (v[j]) {   // count into off[bucket + 1] (16-bit halves of 32-bit words; sums stay < 2^16)
                            const uint32_t at = (((wb[j] | (uint32_t)(__ffs(v[j]) - 1)) >> bsh) + 1u);
                            v[j] &= v[j] - 1;
                            atoms_add(off_u32 + (at >> 1) * 4u, 1u << ((at & 1u) * 16u));
                        }
                }
            }
            __syncthreads();
            {   // exclusive scan of off[1 .. nbk] in place: off[b + 1] = first entry of bucket b
                const int per = (nbk + SQ_THREADS - 1) / SQ_THREADS;
                const int b0 = 1 + threadIdx.x * per, b1 = min(b0 + per, nbk + 1);
                uint32_t sum = 0;
                for (int i = b0; i < b1; ++i) sum += off[i];
                uint32_t incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += x;
                }
                if (lane == 31) scan_s[warp] = incl;
                __syncthreads();
                uint32_t run = incl - sum;
                for (int w = 0; w < warp; ++w) run += scan_s[w];
                for (int i = b0; i < b1; ++i) {
                    const uint32_t c = off[i];
                    off[i] = (uint16_t)run;
                    run += c;
                }
            }
            __syncthreads();
            for (int t = t0 + warp; t < t1; t += SQ_WARPS) {
                const int n = (int)t_cnt[t];  // 0: empty or dense-flagged tile
                const uint16_t* ew = prm.qi.ent_word + (size_t)t * SQ_T1;
                const uint32_t* ev = prm.qi.ent_val + (size_t)t * SQ_T1;
                const uint8_t* er = prm.qi.ent_row + (size_t)t * SQ_T1;
                for (int e0 = lane; e0 < n; e0 += SQ_IB * 32) {
                    uint32_t v[SQ_IB], wb[SQ_IB], r8[SQ_IB];
#pragma unroll
                    for (int j = 0; j < SQ_IB; ++j) {
                        const int e = e0 + j * 32;
                        v[j] = e < n ? ev[e] : 0u;
                        wb[j] = e < n ? (uint32_t)ew[e] << 5 : 0u;
                        r8[j] = e < n ? (uint32_t)((t - t0) * SQ_TQ + er[e]) << 3 : 0u;
                    }
#pragma unroll
                    for (int j = 0; j < SQ_IB; ++j)
                        while (v[j]) {   // afterwards off[b + 1] = end of bucket b = start of bucket b + 1
                            const uint32_t bitid = wb[j] | (uint32_t)(__ffs(v[j]) - 1);
                            v[j] &= v[j] - 1;
                            const uint32_t at = (bitid >> bsh) + 1u, hs = (at & 1u) * 16u;
                            const uint32_t pos = (atoms_add(off_u32 + (at >> 1) * 4u, 1u << hs) >> hs) & 0xffffu;
                            row_s[pos] = (uint16_t)(r8[j] | (bitid & blo));
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
        auto batch_slots = [&](int b) { return (min(SQ_BATCH_ROWS, n_rows - b * SQ_BATCH_ROWS) + R - 1) >> lgR; };
        int ib = warp, is = 0, ib_slots = ib < n_batches ? batch_slots(ib) : 0;
        const uint32_t* isrc = prm.pbits + (r_beg + (int64_t)ib * SQ_BATCH_ROWS) * prm.pitch_words;   // next slot's rows
        auto issue_next = [&](int at) {   // the next slot of this warp's sequence into ring position `at`
            if (lane == 0) {
                uint32_t bytes = row_bytes;
                if (R != 1) bytes = (uint32_t)min(R, n_rows - ib * SQ_BATCH_ROWS - (is << lgR)) * row_bytes;
                mbar_arrive_expect_tx(&my_bars[at], bytes);
                bulk_load(ring_u32 + (uint32_t)at * slot_stride, isrc, bytes, smem_u32(&my_bars[at]), l2_stream);
            }
            isrc += (size_t)R * prm.pitch_words;
            if (++is >= ib_slots) {
                is = 0;
                ib += SQ_WARPS;
                ib_slots = ib < n_batches ? batch_slots(ib) : 0;
                isrc = prm.pbits + (r_beg + (int64_t)ib * SQ_BATCH_ROWS) * prm.pitch_words;
            }
        };
        {
            int at = slot;
            for (int j = 0; j < NS && ib < n_batches; ++j) {
                issue_next(at);
                if (++at == NS) at = 0;
            }
        }

        // fallback path for one set pool bit `bitid` of pool row p_rel (relative to the stripe): every matching entry
        // of its bucket is a hit, queued one per lane and completed 32 at a time by flush_hits
        auto pool_bit_slow = [&](uint32_t bitid, uint32_t p_rel) {
            const uint32_t bucket = bitid >> bsh, lo = bitid & blo;
            const int beg = (int)off[bucket], end = (int)off[bucket + 1];
            for (int e0 = beg; e0 < end; e0 += 32) {
                const int e = e0 + lane;
                const uint32_t x = e < end ? (uint32_t)row_s[e] : 0u;
                uint32_t mb = __ballot_sync(0xffffffffu, e < end && (x & 7u) == lo);
                while (mb) {
                    const int src = __ffs(mb) - 1;
                    mb &= mb - 1;
                    const uint32_t rr = __shfl_sync(0xffffffffu, x, src) >> 3;
                    if (lane == qn) {
                        hit_a = rr | (bitid << 13);
                        hit_b = p_rel;
                    }
                    if (++qn == 32) {
                        flush_hits(prm, hit_a, hit_b, qn, t0, stripe, r_beg, ls);
                        qn = 0;
                    }
                }
            }
        };
        // a batch too dense for the per-warp lists: its rows are read again (global memory, L2) and every set bit
        // takes the fallback path
        auto slow_batch = [&](int b) {
            const int nr = min(SQ_BATCH_ROWS, n_rows - b * SQ_BATCH_ROWS);
#pragma unroll 1
            for (int r = 0; r < nr; ++r) {
                const uint32_t p_rel = (uint32_t)(b * SQ_BATCH_ROWS + r);
                const uint32_t* prow = prm.pbits + (r_beg + p_rel) * prm.pitch_words;
#pragma unroll 1
                for (int w0 = 0; w0 < prm.n_words; w0 += 32) {
                    const uint32_t v = w0 + lane < prm.n_words ? __ldg(prow + w0 + lane) : 0u;
                    uint32_t bm = __ballot_sync(0xffffffffu, v != 0u);
#pragma unroll 1
                    while (bm) {
                        const int src = __ffs(bm) - 1;
                        bm &= bm - 1;
                        uint32_t pv = __shfl_sync(0xffffffffu, v, src);
#pragma unroll 1
                        while (pv) {
                            pool_bit_slow(((uint32_t)(w0 + src) << 5) | (uint32_t)(__ffs(pv) - 1), p_rel);
                            pv &= pv - 1;
                        }
                    }
                }
            }
        };

        for (int b = warp; b < n_batches; b += SQ_WARPS) {
            const int nvs = batch_slots(b);
            const int64_t gp0 = r_beg + (int64_t)b * SQ_BATCH_ROWS;
            // cardinalities of the batch's rows: lane i holds row i (one load, ready long before the lookups)
            const uint32_t pc_lane = (lane < SQ_BATCH_ROWS && gp0 + lane < r_end) ? __ldg(prm.pcard + gp0 + lane) : 0u;
            for (int s = 0; s < nvs; ++s) {
                const int nr = ROW1 ? 1 : min(R, n_rows - b * SQ_BATCH_ROWS - (s << lgR));   // rows of this slot
                const uint32_t n_units = (uint32_t)nr * upr;
                const uint32_t sbase = ring_u32 + (uint32_t)slot * slot_stride;
                const bool fast = ROW1 || (nr == R && full_units <= 6 * 32);   // a full slot of <= 192 units: precomputed offsets
                mbar_wait(&my_bars[slot], phase);
                uint32_t u0 = 0;
                do {   // 192 units at a time (ROW1: exactly one pass)
                    uint4 v[NJ];
                    uint32_t nzu[NJ];
                    if (ROW1) {   // units lane, lane + 32, ...: immediates off one base; only the last may be past the row
                        const uint32_t a0 = sbase + (uint32_t)lane * 16u;
#pragma unroll
                        for (int j = 0; j < NJ - 1; ++j) lds128s(v[j], a0 + (uint32_t)j * 512u);
                        lds128s(v[NJ - 1], sbase + loff[NJ - 1]);
                    } else {
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            uint32_t a = loff[j];
                            if (!fast) {
                                const uint32_t u = u0 + j * 32 + lane;
                                a = u < n_units ? u * 16u : (uint32_t)prm.slot_bytes;
                            }
                            lds128s(v[j], sbase + a);
                        }
                    }
                    uint32_t m = 0;
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        nzu[j] = v[j].x | v[j].y | v[j].z | v[j].w;
                        m |= nzu[j];
                    }
                    // the ballot needs every lane's loads: after it the slot's words are in registers
                    const uint32_t anyb = __ballot_sync(0xffffffffu, m != 0u);
                    if ((ROW1 || u0 + 6 * 32 >= n_units) && ib < n_batches) issue_next(slot);   // refill the slot at once
                    if (anyb == 0u || prm.debug == 1) continue;
                    if (m != 0u) {   // divergent: the few lanes holding non-zero words push them to the warp's word list
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            if (nzu[j] == 0u) continue;
                            const uint32_t u = u0 + j * 32 + lane;
                            uint32_t rs = 0, uw = u;   // row inside the slot, unit inside the row
                            if (R != 1) {
                                rs = u / upr;
                                uw = u - rs * upr;
                            }
                            const uint32_t tag = ((((uint32_t)(s << lgR) + rs)) << 11) | (uw * 4u);   // row in batch | word id
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint32_t x = q == 0 ? v[j].x : (q == 1 ? v[j].y : (q == 2 ? v[j].z : v[j].w));
                                if (x == 0u) continue;
                                const uint32_t pos = atoms_add(uln_a, 1u);
                                if (pos < (uint32_t)SQ_UL_CAP) {
                                    sts_u32(ul_a + pos * 4u, x);
                                    sts_u16(ut_a + pos * 2u, tag + q);
                                }
                            }
                        }
                    }
                } while (!ROW1 && (u0 += 6 * 32) < n_units);
                if (++slot == NS) {
                    slot = 0;
                    phase ^= 1;
                }
            }
            // ---- the batch's lookup phase: every unit of the warp's 8 pool rows has been seen
            __syncwarp();
            const int nu = (int)lds_volatile_u32(uln_a);
            if (nu) {
                int nb = SQ_BL_CAP + 1;
                if (nu <= SQ_UL_CAP) {
                    // lanes take the words into registers, then the scratch is reused: set bits -> bit list
                    uint32_t xv[3], xt[3];
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const int i = q * 32 + lane;
                        xv[q] = 0u;
                        xt[q] = 0u;
                        if (i < nu) {
                            xt[q] = lds_u16(ut_a + (uint32_t)i * 2u);
                            xv[q] = (xt[q] & 0x7ffu) < (uint32_t)prm.n_words ? lds_u32(ul_a + (uint32_t)i * 4u) : 0u;   // past the vocabulary: padding
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        uint32_t x = xv[q];
                        const uint32_t wb = (xt[q] & 0x7ffu) << 5, ptag = (xt[q] >> 11) << 16;
                        while (x) {
                            const uint32_t bitid = wb | (uint32_t)(__ffs(x) - 1);
                            x &= x - 1;
                            const uint32_t pos = atoms_add(bln_a, 1u);
                            if (pos < (uint32_t)SQ_BL_CAP) sts_u32(bl_a + pos * 4u, bitid | ptag);
                        }
                    }
                    __syncwarp();
                    nb = (int)lds_volatile_u32(bln_a);
                }
                if (nb > SQ_BL_CAP || (nb > 0 && batch_hits(prm, bl_a, nb, hb_a, off_u32, row_u32, t0, stripe, gp0, pc_lane, ls)))
                    slow_batch(b);
                __syncwarp();
                if (lane == 0) {
                    sts_u32(uln_a, 0u);
                    sts_u32(bln_a, 0u);
                }
                __syncwarp();
            }
        }
        if (qn) {
            flush_hits(prm, hit_a, hit_b, qn, t0, stripe, r_beg, ls);
            qn = 0;
        }

        __syncthreads();  // every hit of this item is stored
        {
            uint8_t* dst = prm.qi.cnt + (size_t)row0 * prm.n_stripes + stripe;   // query-major [nq][n_stripes]
            const uint8_t* alloc = smem + SQ_SM_ALLOC;
            for (int i = threadIdx.x; i < g_rows; i += SQ_THREADS) dst[(size_t)i * prm.n_stripes] = alloc[i];
        }
    }
}

// ---------------------------------------------------------------------------- host side
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Workspace of the query-index path = [index of the WHOLE call (by-row word lists of every tile, one any-dense flag per
// 8 192-row batch)] [per-batch region, reused by every batch: candidate arrays, their counts, per-stripe counts].
size_t sparseq_index_bytes(int64_t nq_total) {
    const size_t n_qtiles = (size_t)((nq_total + SQ_TQ - 1) / SQ_TQ);
    const size_t n_batches = (size_t)((nq_total + SQ_QB - 1) / SQ_QB);
    return 256 + align256(n_qtiles * 4) * 3 + align256(n_qtiles * SQ_ROWOFF_LD * 2) + align256(n_qtiles * SQ_T1 * 2) +
           align256(n_qtiles * SQ_T1 * 4) + align256(n_qtiles * SQ_T1) + align256(n_batches * 4);
}

size_t sparseq_batch_bytes(int64_t nq_batch, int32_t n_stripes) {
    return 256 + align256((size_t)nq_batch * SQ_GC * 16) + align256((size_t)nq_batch * 4) +
           align256((size_t)n_stripes * (size_t)nq_batch);
}

bool sparseq_supported(int32_t words, int32_t k) {
    return options().jaccard_sparse_q != 0 && options().jaccard_skip_zero != 0 && words <= SQ_MAX_WORDS && k <= R4D_TOPK_MAX &&
           k <= 32;
}

QIndex sparseq_carve_index(void* base, int64_t nq_total) {
    const size_t n_qtiles = (size_t)((nq_total + SQ_TQ - 1) / SQ_TQ);
    const size_t n_batches = (size_t)((nq_total + SQ_QB - 1) / SQ_QB);
    uint8_t* p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(base) + 255) & ~uintptr_t(255));
    QIndex qi{};
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
    qi.any_dense = reinterpret_cast<uint32_t*>(p);   // [n_batches]
    (void)n_batches;
    return qi;
}

// The view of batch `q0 / SQ_QB` (rows q0 .. q0 + nq_batch): its tiles of the call-wide index + the per-batch arrays.
QIndex sparseq_batch_view(const QIndex& all, int64_t q0, void* batch_base, int64_t nq_batch, int32_t n_stripes) {
    const size_t t0 = (size_t)(q0 / SQ_TQ);
    QIndex qi = all;
    qi.ent_val += t0 * SQ_T1;
    qi.ent_word += t0 * SQ_T1;
    qi.ent_row += t0 * SQ_T1;
    qi.rowoff += t0 * SQ_ROWOFF_LD;
    qi.tile_cnt += t0;
    qi.tile_bits += t0;
    qi.tile_dense += t0;
    qi.any_dense += (size_t)(q0 / SQ_QB);
    uint8_t* p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(batch_base) + 255) & ~uintptr_t(255));
    qi.glist = reinterpret_cast<uint4*>(p);
    p += align256((size_t)nq_batch * SQ_GC * 16);
    qi.gcount = reinterpret_cast<uint32_t*>(p);   // gcount and cnt are adjacent: one memset clears them
    p += align256((size_t)nq_batch * 4);
    qi.cnt = p;
    (void)n_stripes;
    return qi;
}

// ONE launch builds the by-row index of every tile of the call (all batches).
int sparseq_build_all(const uint32_t* qbits, int64_t nq_total, int32_t words, int32_t pitch_words, const QIndex& all,
                      cudaStream_t st) {
    const int64_t n_qtiles = (nq_total + SQ_TQ - 1) / SQ_TQ;
    const size_t n_batches = (size_t)((nq_total + SQ_QB - 1) / SQ_QB);
    R4D_CUDA(cudaMemsetAsync(all.any_dense, 0, n_batches * 4, st));
    qindex_kernel<<<(unsigned)n_qtiles, QI_THREADS, 0, st>>>(qbits, nq_total, words, pitch_words, all); note_launch();
    R4D_CUDA(cudaGetLastError());
    return R4D_OK;
}

// Clears the candidate counts of one batch.
int sparseq_clear_batch(const QIndex& qi, int64_t nq, int32_t n_stripes, cudaStream_t st) {
    R4D_REQUIRE(nq <= SQ_QB, "jaccard query-index path: batch of %lld rows > %d", (long long)nq, SQ_QB);
    R4D_CUDA(cudaMemsetAsync(qi.gcount, 0, align256((size_t)nq * 4) + (size_t)n_stripes * (size_t)nq, st));
    return R4D_OK;
}

int sparseq_topk_launch(const uint32_t* pbits, const uint32_t* qcard, const uint32_t* pcard, int64_t nq, int64_t np,
                        int32_t words, int32_t pitch_words, int32_t k, int32_t zero_diag, int64_t query_base,
                        int64_t pool_base, int32_t n_qtiles, int32_t n_ptiles, int32_t n_stripes, int32_t ptiles_per_stripe,
                        uint4* part, const QIndex& qi, cudaStream_t st) {
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
    while (R > 1 && R * pitch_bytes + 16 > SQ_RING_BYTES / 3) R >>= 1;
    if (R > 1) {
        prm.slot_rows = R;
        prm.slot_bytes = (int32_t)(R * pitch_bytes);
        prm.row_units = pitch_words / 4;
    } else {
        prm.slot_rows = 1;
        prm.slot_bytes = ((words + 3) / 4) * 16;
        prm.row_units = prm.slot_bytes / 16;
    }
    prm.slot_rows_log2 = 0;
    while ((1 << prm.slot_rows_log2) < prm.slot_rows) ++prm.slot_rows_log2;
    prm.bucket_shift = 0;
    while ((((int64_t)words * 32 + (1 << prm.bucket_shift) - 1) >> prm.bucket_shift) > SQ_NBK) ++prm.bucket_shift;
    prm.n_buckets = (int32_t)(((int64_t)words * 32 + (1 << prm.bucket_shift) - 1) >> prm.bucket_shift);
    prm.n_slots = SQ_RING_BYTES / (prm.slot_bytes + 16);   // every slot is followed by 16 zero bytes
    if (prm.n_slots > SQ_MAX_SLOTS) prm.n_slots = SQ_MAX_SLOTS;
    R4D_REQUIRE(prm.n_slots >= 1, "jaccard query-index path: a row of %d words does not fit the ring", words);
    prm.part = part;
    prm.qi = qi;
    prm.debug = options().jaccard_debug;
    // NU = units per lane of a one-row slot (0: generic geometry)
    const int nu = (prm.slot_rows == 1 && prm.row_units <= 6 * 32) ? (prm.row_units + 31) / 32 : 0;
    void (*kern)(const SparseParams) = jaccard_qindex_kernel<0>;
    switch (nu) {
        case 1: kern = jaccard_qindex_kernel<1>; break;
        case 2: kern = jaccard_qindex_kernel<2>; break;
        case 3: kern = jaccard_qindex_kernel<3>; break;
        case 4: kern = jaccard_qindex_kernel<4>; break;
        case 5: kern = jaccard_qindex_kernel<5>; break;
        case 6: kern = jaccard_qindex_kernel<6>; break;
        default: break;
    }
    static SmemOptIn opt_in[7];   // per kernel instantiation, keyed by device inside
    if (int rc = ensure_dyn_smem(kern, SQ_SM_TOTAL, opt_in[nu])) return rc;
    int grid = num_sms();
    if (n_stripes < grid) grid = n_stripes;
    prof_begin(PROF_JACCARD_QINDEX, st);
    kern<<<grid, SQ_THREADS, SQ_SM_TOTAL, st>>>(prm); note_launch();
    prof_end(PROF_JACCARD_QINDEX, st);
    R4D_CUDA(cudaGetLastError());
    return R4D_OK;
}

}  // namespace r4d
