// jaccard_postings.cu — fused Jaccard top-K as a sparse x sparse join over POOL-SIDE POSTINGS.
//
// Replaces occurrence_matrix + np.argsort(-row)[:k] (retrieval_data_annotation.py:36-41, :97-103) for SPARSE sets
// (label sets: 2.2 of 20 000 ids): only pairs that share an id can score > 0, and an inverted index of the pool
// names exactly those pairs.  Where the query-index kernel (jaccard_sparse.cu) still streams every pool bitset row
// once per 8 192 queries (2.5 GB per pass at C4), this path touches only the ~110 postings of each of a query's ids.
//
//   pool state (built once per pool shard from its bitsets, r4d_postings_build):
//     post[]   8-byte entries {pool row, |pool set|}, grouped by (node id, row window): bucket b = id * n_win + (row >> win_shift)
//     off[]    bucket offsets (exclusive scan of the bucket sizes); an id's whole posting list is the contiguous
//              range off[id * n_win] .. off[(id + 1) * n_win], and any run of windows of it is contiguous too
//     best[]   per id: its 32 postings with the smallest (|pool set|, row), ascending (the head of the id's ranking)
//   r4d_jaccard_topk_postings (queries arrive as CSR id lists; duplicates inside a row collapse, like Python's set()) — a
//   chain of kernels, each serving what it can and handing the rest to the next through a list in the workspace:
//     postings_big_kernel     label-like sets, queries of more than 8 ids: the head kernel's algorithm with a whole CTA
//                             per query (listed first by postings_big_scan_kernel).
//     postings_head_kernel    label-like sets (the headline): one WARP per query.  Rows holding ONE of the query's ids are
//                             ranked by merging the per-id best lists; only rows holding SEVERAL ids need the join: the
//                             posting rows are streamed through a Bloom filter in shared memory, the few noted rows are
//                             counted exactly by bucket probes.
//     postings_reg_kernel     first stage when k > 16 (and the stage behind the head kernel): one WARP per query, the
//                             pass's <= 256 postings held in REGISTERS, repeats of a pool row found with an 8 192-bit
//                             filter in shared memory and resolved by a warp-wide compare; warp-level sorted top-K list.
//     postings_light_kernel   history-like sets (hundreds to thousands of postings per query): one WARP per query.  The
//                             (<= 64) distinct ids of the query select their posting ranges; every posting is inserted
//                             into the warp's 512 / 1 024-slot hash table in shared memory (64-bit CAS claims a slot for
//                             a pool row, a 32-bit atomic add counts further hits), so a slot's count IS |Q n P| and
//                             union = |Q| + |P| - count needs no second look at the pool.  The claimed slots are scanned
//                             into the warp-level sorted top-K list.
//                             Both walk the pool's row windows in several passes when a query has more postings than a
//                             pass holds (disjoint row ranges => a pair never spans two passes).
//     postings_heavy_kernel   one CTA per query the others handed over (more than 64 ids, a row window with more
//                             postings than a pass holds): the query becomes a bitmap over the ids, every window of the
//                             pool gets one 16-bit counter per row in shared memory, postings increment them, a scan
//                             turns non-zero counters into candidates.  Any set size, any skew; slower.
//   All write the final [nq][k] lists themselves (zero-score fillers included), or store them into the peers' gather
//   buffers (fused exchange) — no candidate ever goes through global memory.
//
// Exactness: counts are integers, ranking is the same (inter * union' vs inter' * union, index) comparator as everywhere
// else; results are bit-identical to the bitset kernels (tests/test_gpu_postings.py).
#include "jaccard_common.cuh"

namespace r4d {

constexpr int PJ_WIN_SHIFT_MIN = 13;            // 8 192 pool rows per window
constexpr int PJ_WIN_SHIFT_MAX = 15;            // heavy kernel: a window's 16-bit counters fill 64 KB of shared memory
constexpr int64_t PJ_MAX_BUCKETS = 48ll << 20;  // (id, window) buckets: at most 192 MB of offsets
constexpr int PJ_MAX_BITS = 65535;              // ids fit 16 bits; a heavy counter (<= |Q n P| <= 65 535) cannot wrap
// A warp's hash table has 1 << LOG_T slots of 8 B (LOG_T = 9 or 10); a pass plans for a load factor of ~0.44.
// 512 slots keep 32 warps per SM resident (label-like sets: one pass per query); 1 024 slots halve the number of passes
// of a query whose posting lists are long (history-like sets).
constexpr int PJ_LOG_T_SMALL = 9, PJ_LOG_T_LARGE = 10;
__host__ __device__ constexpr int pj_cap(int log_t) { return (7 << log_t) / 16; }   // postings planned per pass
constexpr int PJ_IDS = 64;                      // distinct ids of a light query: two per lane
constexpr int PJ_LIGHT_WARPS = 8;
constexpr int PJ_CHUNK = 8;                     // queries a warp takes per grab of the work counter
constexpr int PJ_HEAVY_THREADS = 256;
constexpr int PJ_HEAVY_UID = 2048;              // distinct ids of a heavy query enumerated in shared memory
constexpr int PJ_BEST = 32;                      // per id: its 32 best postings by (|pool set| asc, row asc), one per lane
constexpr uint32_t PJ_MAGIC = 0x52344451u;      // "R4DQ" (layout with the per-id best lists)

struct PostingsHeader {   // first 256 bytes of the index blob (device memory)
    uint32_t magic, status;   // status != 0: the build overflowed nnz_cap (index unusable)
    int64_t np, nnz_cap;
    int32_t n_bits, win_shift, n_win;
    uint32_t max_bucket;   // most postings of one id inside one row window (written by the build)
};

struct PostingsLayout {
    int win_shift, n_win;
    int64_t n_buckets;
    size_t off_at, post_at, best_at, total;
    bool ok;
};

static PostingsLayout postings_layout(int64_t np, int32_t n_bits, int64_t nnz) {
    PostingsLayout L{};
    L.ok = np >= 0 && n_bits > 0 && n_bits <= PJ_MAX_BITS && nnz >= 0 && nnz < ((int64_t)1 << 32) && np < ((int64_t)1 << 31);
    if (!L.ok) return L;
    int ws = PJ_WIN_SHIFT_MIN;
    auto wins = [&](int s) { return np == 0 ? (int64_t)1 : ((np + ((int64_t)1 << s) - 1) >> s); };
    while (ws < PJ_WIN_SHIFT_MAX && wins(ws) * n_bits > PJ_MAX_BUCKETS) ++ws;
    L.win_shift = ws;
    L.n_win = (int)wins(ws);
    L.n_buckets = (int64_t)L.n_win * n_bits;
    if (L.n_buckets > PJ_MAX_BUCKETS) {
        L.ok = false;
        return L;
    }
    L.off_at = 256;
    L.post_at = L.off_at + (((size_t)(L.n_buckets + 1) * 4 + 255) / 256) * 256;
    L.best_at = L.post_at + (((size_t)(nnz > 0 ? nnz : 1) * 8 + 255) / 256) * 256;
    L.total = L.best_at + (size_t)n_bits * PJ_BEST * 8;
    return L;
}

// ---------------------------------------------------------------------------- index build
// One warp per pool row: every set bit is one posting.  COUNT: bucket sizes (atomicAdd into off[b + 1]);
// FILL: slot = off[b] + atomicAdd(cursor[b]).  The order inside a bucket is arbitrary; every consumer is order independent.
template <bool FILL>
__global__ void __launch_bounds__(256)
postings_scan_rows_kernel(const uint32_t* __restrict__ pbits, const uint32_t* __restrict__ pcard, int64_t np,
                          int32_t words, int32_t pitch_words, int32_t n_bits, int32_t win_shift, int32_t n_win,
                          uint32_t* __restrict__ off, uint32_t* __restrict__ cursor, uint2* __restrict__ post,
                          int64_t nnz_cap, PostingsHeader* __restrict__ hdr) {
    const int lane = threadIdx.x & 31;
    const int64_t wpg = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < np; row += wpg) {
        const uint32_t* r = pbits + row * pitch_words;
        const uint32_t win = (uint32_t)(row >> win_shift);
        const uint32_t card = FILL ? pcard[row] : 0u;
        for (int w0 = 0; w0 < words; w0 += 4 * 32) {   // four independent loads in flight per lane
            uint32_t v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int w = w0 + j * 32 + lane;
                v[j] = w < words ? r[w] : 0u;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t x = v[j];
                const int w = w0 + j * 32 + lane;
                while (x) {
                    const int b = __ffs(x) - 1;
                    x &= x - 1;
                    const int32_t id = w * 32 + b;
                    if (id >= n_bits) break;
                    const int64_t bucket = (int64_t)id * n_win + win;
                    if (!FILL) {
                        atomicAdd(off + bucket + 1, 1u);
                    } else {
                        const uint32_t at = off[bucket] + atomicAdd(cursor + bucket, 1u);
                        if ((int64_t)at < nnz_cap)
                            post[at] = make_uint2((uint32_t)row, card);
                        else
                            hdr->status = 1u;
                    }
                }
            }
        }
    }
}

__global__ void postings_header_kernel(PostingsHeader* hdr, const PostingsHeader h) { *hdr = h; }

__global__ void __launch_bounds__(256) postings_max_bucket_kernel(const uint32_t* __restrict__ off, int64_t n_buckets,
                                                                 PostingsHeader* __restrict__ hdr) {
    uint32_t m = 0;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_buckets; b += (int64_t)gridDim.x * blockDim.x)
        m = max(m, off[b + 1] - off[b]);
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(&hdr->max_bucket, m);
}

// best[id][0 .. PJ_BEST): the postings of `id` with the smallest (|pool set|, row), ascending, padded with
// {0xffffffff, 0}.  A query that holds ONE id scores every pool row of that list 1 / |pool set| (intersection 1, union
// |pool set|), so its top-K is simply the head of this list — no join at all (postings_reg_kernel).  One warp per id,
// built once with the index.
struct BestEntry {
    uint32_t card, row;
    __device__ __forceinline__ static BestEntry worst() { return BestEntry{0xffffffffu, 0xffffffffu}; }
    __device__ __forceinline__ static bool better(const BestEntry& a, const BestEntry& b) {
        return a.card < b.card || (a.card == b.card && a.row < b.row);
    }
    __device__ __forceinline__ BestEntry shfl(int src) const {
        return BestEntry{__shfl_sync(0xffffffffu, card, src), __shfl_sync(0xffffffffu, row, src)};
    }
    __device__ __forceinline__ BestEntry shfl_up1() const {
        return BestEntry{__shfl_up_sync(0xffffffffu, card, 1), __shfl_up_sync(0xffffffffu, row, 1)};
    }
};

__global__ void __launch_bounds__(256) postings_best_kernel(const uint32_t* __restrict__ off, const uint2* __restrict__ post,
                                                           int32_t n_bits, int32_t n_win, uint2* __restrict__ best) {
    const int lane = threadIdx.x & 31;
    const int64_t wpg = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t id = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); id < n_bits; id += wpg) {
        const uint32_t s = off[id * n_win], e = off[(id + 1) * n_win];
        WarpTopK<BestEntry> tk;
        tk.init(PJ_BEST);
        for (uint32_t t0 = s; t0 < e; t0 += 32) {
            BestEntry c = BestEntry::worst();
            if (t0 + lane < e) {
                const uint2 v = post[t0 + lane];
                c = BestEntry{v.y, v.x};
            }
            uint32_t m = __ballot_sync(0xffffffffu, BestEntry::better(c, tk.kth));
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                tk.insert(c.shfl(src));
            }
        }
        best[id * PJ_BEST + lane] = make_uint2(tk.mine.row, tk.mine.card == 0xffffffffu ? 0u : tk.mine.card);
    }
}

// In-place inclusive scan of x[0 .. n) (uint32), three launches: block sums, scan of the sums, rescan + offset.
constexpr int SCAN_THREADS = 256, SCAN_PER_THREAD = 16, SCAN_TILE = SCAN_THREADS * SCAN_PER_THREAD;

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* warp_tot /*[8] smem*/, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        const uint32_t t = warp_tot[w];
        if (w < warp) base += t;
        tot += t;
    }
    __syncthreads();
    total = tot;
    return base + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums_kernel(const uint32_t* __restrict__ x, int64_t n,
                                                                      uint32_t* __restrict__ tile_sum) {
    __shared__ uint32_t wt[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
    for (int i = threadIdx.x; i < SCAN_TILE; i += SCAN_THREADS)
        if (base + i < n) s += x[base + i];
    uint32_t tot;
    block_excl_scan(s, wt, tot);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(uint32_t* __restrict__ tile_sum, int64_t n_tiles) {
    __shared__ uint32_t wt[SCAN_THREADS / 32];
    uint32_t carry = 0;
    for (int64_t t0 = 0; t0 < n_tiles; t0 += SCAN_THREADS) {
        const int64_t t = t0 + threadIdx.x;
        const uint32_t v = t < n_tiles ? tile_sum[t] : 0u;
        uint32_t tot;
        const uint32_t ex = block_excl_scan(v, wt, tot);
        if (t < n_tiles) tile_sum[t] = carry + ex;   // exclusive prefix of the tile sums
        carry += tot;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(uint32_t* __restrict__ x, int64_t n,
                                                                  const uint32_t* __restrict__ tile_excl) {
    __shared__ uint32_t wt[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_PER_THREAD;
    uint32_t v[SCAN_PER_THREAD], s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; ++i) {
        v[i] = base + i < n ? x[base + i] : 0u;
        s += v[i];
    }
    uint32_t tot;
    uint32_t run = block_excl_scan(s, wt, tot) + tile_excl[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; ++i) {
        run += v[i];
        if (base + i < n) x[base + i] = run;
    }
}

// ---------------------------------------------------------------------------- scoring kernels
struct PJParams {
    const int32_t* q_ids;
    const int64_t* q_off;
    int64_t nq;
    const uint32_t* off;
    const uint2* post;
    const uint2* best;      // [n_bits][PJ_BEST] per-id best postings (nullptr: not used)
    const PostingsHeader* hdr;
    const uint32_t* pcard;
    int64_t np;
    int32_t n_bits, win_shift, n_win, k, zero_diag, n_fill;
    int64_t query_base, pool_base;
    uint32_t* out_inter;
    uint32_t* out_union;
    int32_t* out_idx;
    // Packed results (r4d_jaccard_topk_postings_packed; out_qcard != nullptr): out_inter holds `inter << 16 | |pool set|`
    // per entry, out_qcard[q] = |query set|, out_union is not written: 8 bytes per entry instead of 12 for results that
    // cross PCIe.  union = |query set| + |pool set| - inter; a padding entry (idx R4D_IDX_NONE) packs as 0.
    uint32_t* out_qcard;
    // Kernel chain of a call: first-stage light kernel (all queries) -> [hash-table kernel on the queries the register
    // kernel handed over] -> heavy kernel on what is left.  Every stage takes its work from `work`, serves the queries
    // in_list[0 .. *in_count) (in_list == nullptr: queries 0 .. nq-1) and appends what it cannot serve to hand_list.
    const uint32_t* in_list;
    const uint32_t* in_count;
    uint32_t* work;
    uint32_t* hand_list;
    uint32_t* hand_count;
    // Queries with more than big_ids ids (listed by postings_big_scan_kernel) are served by postings_big_kernel, a whole CTA
    // per query, before the head kernel (which skips them): one warp takes ~100 us for a 25-id query — alone that is the
    // length of the whole launch.
    const uint32_t* big_list;
    const uint32_t* big_count;
    uint32_t* big_work;
    int32_t big_ids;
    // Head kernel, packed lists bound for pinned HOST memory (relay_done != nullptr): out_inter / out_idx / out_qcard point
    // at a staging copy in HBM, and the warp that completes a block of PJ_RELAY queries (relay_done[block] counts them)
    // copies the block to its final place with full 128-byte stores — PCIe writes from an SM are bound by their NUMBER
    // (~400 per us, measured: tools/probe_host_store.cu), so 40-byte rows or 320-byte chunks waste most of the link.
    uint32_t* relay_pair;
    int32_t* relay_idx;
    uint32_t* relay_qcard;
    uint32_t* relay_done;
    PeerOut peers;
    int64_t q_out_off, nq_total;   // fused exchange: row offset / rows of the whole call
    int32_t chunk;                 // queries a light warp takes per grab of the work counter (<= PJ_CHUNK)
};

// Candidate with a 32-bit exact compare: valid while inter * union < 2^32 (light path: inter <= 64, union < 2^17).
struct PEntry {
    uint32_t inter, uni;
    int32_t idx;
    __device__ __forceinline__ static PEntry worst() { return PEntry{0u, 1u, R4D_IDX_NONE}; }
    __device__ __forceinline__ static bool better(const PEntry& a, const PEntry& b) {
        const uint32_t l = a.inter * b.uni, r = b.inter * a.uni;
        return (l > r) || (l == r && a.idx < b.idx);
    }
    __device__ __forceinline__ PEntry shfl(int src) const {
        return PEntry{__shfl_sync(0xffffffffu, inter, src), __shfl_sync(0xffffffffu, uni, src),
                      __shfl_sync(0xffffffffu, idx, src)};
    }
    __device__ __forceinline__ PEntry shfl_up1() const {
        return PEntry{__shfl_up_sync(0xffffffffu, inter, 1), __shfl_up_sync(0xffffffffu, uni, 1),
                      __shfl_up_sync(0xffffffffu, idx, 1)};
    }
    __device__ __forceinline__ PEntry shfl_xor(int m) const {
        return PEntry{__shfl_xor_sync(0xffffffffu, inter, m), __shfl_xor_sync(0xffffffffu, uni, m),
                      __shfl_xor_sync(0xffffffffu, idx, m)};
    }
};

// Zero-score fillers (pool rows 0 .. n_fill-1 unless listed already; they only matter while the k-th entry scores 0)
// and the final store: plain [nq][k] planes, or slot `rank` of every peer's gather buffer (fused exchange).
// stage != nullptr (light kernel, plain output, k <= PJ_OBUF_K): the list goes to the warp's staging buffer [3][chunk][k];
// the rows of a whole chunk of consecutive queries are then written out together (pj_flush_chunk).
template <class E>
__device__ __forceinline__ void pj_finish(WarpTopK<E>& tk, const PJParams& p, int64_t q, uint32_t cq, const bool packed,
                                          uint32_t* stage = nullptr, int stage_row = 0, int stage_rows = 0) {
    const int lane = threadIdx.x & 31;
    if (p.n_fill > 0 && tk.kth.inter == 0u) {
        const uint32_t cp = lane < p.n_fill ? p.pcard[lane] : 0u;
        for (int i = 0; i < p.n_fill; ++i) {
            const E c{0u, max(cq + __shfl_sync(0xffffffffu, cp, i), 1u), (int32_t)(p.pool_base + i)};
            if (__ballot_sync(0xffffffffu, lane < p.k && tk.mine.idx == c.idx)) continue;
            tk.insert(c);
        }
    }
    // packed entry: |pool set| = union + inter - |query set| (also for the fillers: an empty pair has union 1 = "|pool set| 1")
    const uint32_t first = !packed ? tk.mine.inter
                                   : (tk.mine.idx == R4D_IDX_NONE ? 0u : (tk.mine.inter << 16) | (tk.mine.uni + tk.mine.inter - cq));
    if (stage != nullptr) {
        if (lane < p.k) {
            stage[stage_row * p.k + lane] = first;
            if (!packed) stage[(stage_rows + stage_row) * p.k + lane] = tk.mine.uni;
            stage[(2 * stage_rows + stage_row) * p.k + lane] = (uint32_t)tk.mine.idx;
        }
        if (packed && lane == 0) stage[stage_rows * p.k + stage_row] = cq;   // the union plane's place holds |query set|
        return;
    }
    if (lane < p.k) {
        if (p.peers.world == 0) {
            p.out_inter[q * p.k + lane] = first;
            if (!packed) p.out_union[q * p.k + lane] = tk.mine.uni;
            p.out_idx[q * p.k + lane] = tk.mine.idx;
            if (packed && lane == 0) p.out_qcard[q] = cq;
        } else {
            const int64_t nq_all = p.nq_total > 0 ? p.nq_total : p.nq;
            const int64_t plane = (int64_t)p.peers.world * nq_all * p.k;
            const int64_t at = ((int64_t)p.peers.rank * nq_all + p.q_out_off + q) * p.k + lane;
            for (int r = 0; r < p.peers.world; ++r) {
                uint32_t* dst = reinterpret_cast<uint32_t*>(p.peers.base[r]);
                dst[at] = tk.mine.inter;
                dst[plane + at] = tk.mine.uni;
                dst[2 * plane + at] = (uint32_t)tk.mine.idx;
            }
        }
    }
}

__constant__ uint32_t ph_inv[33] = {   // ceil(2^32 / m): t / m = umulhi(t, ph_inv[m]) for t < 2^16
    0u, 0u, 2147483648u, 1431655766u, 1073741824u, 858993460u, 715827883u, 613566757u, 536870912u, 477218589u, 429496730u,
    390451573u, 357913942u, 330382100u, 306783379u, 286331154u, 268435456u, 252645136u, 238609295u, 226050911u, 214748365u,
    204522253u, 195225787u, 186737709u, 178956971u, 171798692u, 165191050u, 159072863u, 153391690u, 148102321u, 143165577u,
    138547333u, 134217728u};

// The finished lists of a chunk's consecutive queries, staged as [plane][row][k] (pj_finish), go out together: n_here * k
// consecutive words per plane.  The store loop walks the 128-byte line grid of the destination (a warp-wide store that
// straddles two lines becomes two partial writes — in HBM two partial sectors, over PCIe two short TLPs).
__device__ __forceinline__ void pj_flush_chunk(const PJParams& p, const uint32_t* stage, int64_t q0, int n_here,
                                               const bool packed, const uint32_t skip = 0u) {
    const int lane = threadIdx.x & 31;
    const int n_words = n_here * p.k;
    const int64_t at = q0 * p.k;
    const int mis = (int)((reinterpret_cast<uintptr_t>(p.out_idx + at) >> 2) & 31);   // planes of one call are aligned alike
    if (skip == 0u) {
        for (int i = lane - mis; i < n_words; i += 32)
            if (i >= 0) {
                p.out_inter[at + i] = stage[i];
                if (!packed) p.out_union[at + i] = stage[n_words + i];
                p.out_idx[at + i] = (int32_t)stage[2 * n_words + i];
            }
    } else {   // rows of the chunk that another warp has written already (bit r of `skip`) are left alone
        const uint32_t inv = ph_inv[p.k];
        for (int i = lane - mis; i < n_words; i += 32)
            if (i >= 0) {
                const uint32_t r = p.k == 1 ? (uint32_t)i : __umulhi((uint32_t)i, inv);
                if ((skip >> r) & 1u) continue;
                p.out_inter[at + i] = stage[i];
                if (!packed) p.out_union[at + i] = stage[n_words + i];
                p.out_idx[at + i] = (int32_t)stage[2 * n_words + i];
            }
    }
    if (packed && lane < n_here && !((skip >> lane) & 1u)) p.out_qcard[q0 + lane] = stage[n_words + lane];
}

// Relay of finished blocks (see PJParams::relay_done).  `cnt` queries of block q0 / PJ_RELAY are final in the staging copy
// (a chunk never straddles two blocks: PJ_RELAY is a multiple of every chunk size); the warp whose count completes the
// block copies it out.  A block's words are contiguous and 128-byte aligned in every plane (PJ_RELAY * k * 4 = 256 k bytes).
constexpr int PJ_RELAY = 64;
__device__ __noinline__ void pj_relay_copy(const uint32_t* __restrict__ src_pair, const int32_t* __restrict__ src_idx,
                                           const uint32_t* __restrict__ src_qcard, uint32_t* __restrict__ dst_pair,
                                           int32_t* __restrict__ dst_idx, uint32_t* __restrict__ dst_qcard, int64_t first,
                                           int bn, int k) {
    const int lane = threadIdx.x & 31;
    const int64_t w0 = first * k;
    const int nw = bn * k;
    for (int i0 = 0; i0 < nw; i0 += 128) {   // four lines per plane in flight
        uint32_t a[4], c[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = i0 + j * 32 + lane;
            if (i < nw) {
                a[j] = __ldcg(src_pair + w0 + i);
                c[j] = (uint32_t)__ldcg(src_idx + w0 + i);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = i0 + j * 32 + lane;
            if (i < nw) {
                dst_pair[w0 + i] = a[j];
                dst_idx[w0 + i] = (int32_t)c[j];
            }
        }
    }
    for (int i = lane; i < bn; i += 32) dst_qcard[first + i] = __ldcg(src_qcard + first + i);
}
__device__ __forceinline__ void pj_relay(const PJParams& p, int64_t q0, uint32_t cnt) {
    if (cnt == 0u) return;   // warp-uniform
    const int lane = threadIdx.x & 31;
    const int64_t b = q0 / PJ_RELAY;
    __threadfence();   // this warp's staging stores before the count
    __syncwarp();
    uint32_t old = 0u;
    if (lane == 0) old = atomicAdd(p.relay_done + b, cnt);
    old = __shfl_sync(0xffffffffu, old, 0);
    const int64_t first = b * PJ_RELAY;
    const uint32_t bn = (uint32_t)min((int64_t)PJ_RELAY, p.nq - first);
    if (old + cnt != bn) return;
    __threadfence();   // the other warps' staging stores after their counts
    pj_relay_copy(p.out_inter, p.out_idx, p.out_qcard, p.relay_pair, p.relay_idx, p.relay_qcard, first, (int)bn, p.k);
}

constexpr unsigned long long PJ_EMPTY = 0xffffffffffffffffull;
constexpr int PJ_OBUF_K = 12;   // lists up to this width are staged per chunk (wider ones are stored row by row)
constexpr int PJ_OWN = 16;   // table slots a lane may claim per pass before the pass falls back to scanning the whole table

// slot = {row : 32 | card : 24 | count : 8}.  Returns the claimed slot (>= 0) when this call created the entry, -1 when it
// counted one more hit of an existing entry, -2 when the table is full.
template <int LOG_T>
__device__ __forceinline__ int pj_insert(unsigned long long* tab, uint32_t row, uint32_t card) {
    constexpr int PJ_T = 1 << LOG_T;
    uint32_t h = (row * 2654435761u) >> (32 - LOG_T);
    const unsigned long long fresh = ((unsigned long long)row << 32) | (unsigned long long)((card << 8) | 1u);
    for (int probe = 0; probe < PJ_T; ++probe) {
        const unsigned long long old = atomicCAS(tab + h, PJ_EMPTY, fresh);
        if (old == PJ_EMPTY) return (int)h;
        if ((uint32_t)(old >> 32) == row) {
            atomicAdd(reinterpret_cast<unsigned int*>(tab + h), 1u);   // low word: card << 8 | count
            return -1;
        }
        h = (h + 1) & (PJ_T - 1);
    }
    return -2;
}

template <int LOG_T>
struct PJWarpSmem {
    unsigned long long tab[1 << LOG_T];
    uint32_t start[PJ_IDS];
    uint32_t pref[PJ_IDS + 1];
    int32_t ids[PJ_IDS];
    uint16_t own[PJ_OWN * 32];   // own[i * 32 + lane]: i-th slot claimed by the lane in this pass
    uint32_t obuf[3 * PJ_CHUNK * PJ_OBUF_K];   // finished lists of the chunk's queries: [plane][row][k], flushed together
    uint32_t pad[3];
};

template <int LOG_T>
__global__ void __launch_bounds__(PJ_LIGHT_WARPS * 32, LOG_T == PJ_LOG_T_SMALL ? 4 : 2) postings_light_kernel(const PJParams p) {
    extern __shared__ __align__(16) uint8_t pj_smem[];
    constexpr int PJ_T = 1 << LOG_T, PJ_CAP = pj_cap(LOG_T);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    PJWarpSmem<LOG_T>& sm = reinterpret_cast<PJWarpSmem<LOG_T>*>(pj_smem)[warp];
    const bool packed = p.out_qcard != nullptr;
    const uint4 ones = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    auto clear_table = [&]() {
        uint4* t4 = reinterpret_cast<uint4*>(sm.tab);
#pragma unroll
        for (int i = 0; i < PJ_T / 2 / 32; ++i) t4[i * 32 + lane] = ones;
    };
    auto hand_over = [&](int64_t q) {   // the heavy kernel serves this query
        if (lane == 0) p.hand_list[atomicAdd(p.hand_count, 1u)] = (uint32_t)q;
    };
    // second stage of a chain: the queries are the ones the register kernel handed over, one per grab
    const bool listed = p.in_list != nullptr;
    const int64_t n_work = listed ? (int64_t)*p.in_count : p.nq;
    if (n_work == 0) return;   // (an empty list: no table to clear, no grab)
    clear_table();
    bool table_clean = true;
    const int chunk = listed ? 1 : p.chunk;
    for (;;) {
        int64_t q0 = 0;
        if (lane == 0) q0 = (int64_t)atomicAdd(p.work, (uint32_t)chunk);
        q0 = __shfl_sync(0xffffffffu, q0, 0);
        if (q0 >= n_work) break;
        const int64_t q_first = listed ? (int64_t)p.in_list[q0] : q0;
        // row offsets of the whole chunk with one load
        int64_t my_off = 0;
        if (lane <= chunk && q_first + lane <= p.nq) my_off = p.q_off[q_first + lane];
        const int n_here = (int)min((int64_t)chunk, n_work - q0);
        // plain output of narrow lists: the chunk's rows are contiguous in every output plane, so they are staged in
        // shared memory and written out together — n_here * k consecutive words per plane instead of k-word pieces
        // (full sectors in HBM; when the caller's buffers are pinned HOST memory, far fewer and larger PCIe writes)
        const bool staged = !listed && p.peers.world == 0 && p.k <= PJ_OBUF_K;
        for (int qi = 0; qi < n_here; ++qi) {
            const int64_t q = q_first + qi;
            const int64_t beg = __shfl_sync(0xffffffffu, my_off, qi), end = __shfl_sync(0xffffffffu, my_off, qi + 1);
            const int64_t m_raw = end - beg;
            if (m_raw > PJ_IDS) {
                hand_over(q);
                continue;
            }
            // ---- the query's distinct ids, at most two per lane (set semantics: duplicates collapse)
            int32_t id0 = -1, id1 = -1;
            if (lane < m_raw) id0 = p.q_ids[beg + lane];
            if (lane + 32 < m_raw) id1 = p.q_ids[beg + 32 + lane];
            if (id0 < 0 || id0 >= p.n_bits) id0 = -1;
            if (id1 < 0 || id1 >= p.n_bits) id1 = -1;
            if (m_raw <= 32) {
                const uint32_t same = __match_any_sync(0xffffffffu, id0);
                if (id0 >= 0 && (__ffs(same) - 1) != lane) id0 = -1;
            } else {
                sm.ids[lane] = id0;
                sm.ids[lane + 32] = id1;
                __syncwarp();
                bool d0 = false, d1 = false;
                for (int t = 0; t < (int)m_raw; ++t) {
                    const int32_t v = sm.ids[t];
                    d0 |= (t < lane) && (v == id0);
                    d1 |= (t < lane + 32) && (v == id1);
                }
                if (d0) id0 = -1;
                if (d1) id1 = -1;
                __syncwarp();
            }
            const uint32_t cq = __popc(__ballot_sync(0xffffffffu, id0 >= 0)) + __popc(__ballot_sync(0xffffffffu, id1 >= 0));
            const bool two = m_raw > 32;   // warp-uniform: the second id register is in use
            // ---- whole posting lists: how many passes over the pool's row windows?
            uint32_t s0 = 0, e0 = 0, s1 = 0, e1 = 0;
            if (id0 >= 0) {
                s0 = p.off[(int64_t)id0 * p.n_win];
                e0 = p.off[(int64_t)(id0 + 1) * p.n_win];
            }
            if (two && id1 >= 0) {
                s1 = p.off[(int64_t)id1 * p.n_win];
                e1 = p.off[(int64_t)(id1 + 1) * p.n_win];
            }
            const uint32_t hits = __reduce_add_sync(0xffffffffu, (e0 - s0) + (e1 - s1));
            const int passes = (int)((hits + PJ_CAP - 1) / PJ_CAP);
            if (passes > p.n_win) {
                hand_over(q);
                continue;
            }
            WarpTopK<PEntry> tk;
            tk.init(p.k);
            bool failed = false, have_list = false;   // have_list: the sorted list holds entries of an earlier pass
            const bool diag_on = p.zero_diag != 0;
            const int64_t diag_row = p.query_base + q - p.pool_base;   // pool row forced to score 0
            const int w_base = passes > 1 ? p.n_win / passes : 0, w_rem = passes > 1 ? p.n_win - w_base * passes : 0;
            for (int ps = 0; ps < passes; ++ps) {
                if (passes > 1) {   // this pass: windows [wlo, whi) of every list (balanced split of the n_win windows)
                    const int wlo = ps * w_base + min(ps, w_rem), whi = (ps + 1) * w_base + min(ps + 1, w_rem);
                    if (id0 >= 0) {
                        s0 = p.off[(int64_t)id0 * p.n_win + wlo];
                        e0 = p.off[(int64_t)id0 * p.n_win + whi];
                    }
                    if (two && id1 >= 0) {
                        s1 = p.off[(int64_t)id1 * p.n_win + wlo];
                        e1 = p.off[(int64_t)id1 * p.n_win + whi];
                    }
                }
                const uint32_t l0 = e0 - s0, l1 = two ? e1 - s1 : 0u;
                if (!__ballot_sync(0xffffffffu, (l0 | l1) != 0u)) continue;
                if (!table_clean) clear_table();
                table_clean = false;
                // every posting of the pass goes into the hash table; a lane remembers the slots it claimed
                int n_own = 0, bad = 0;   // bad: 1 = table full, 2 = more than PJ_OWN claims (scan the whole table instead)
                auto put = [&](const uint2 e) {
                    if (diag_on && (int64_t)e.x == diag_row) return;
                    const int r = pj_insert<LOG_T>(sm.tab, e.x, e.y);
                    if (r >= 0) {
                        if (n_own < PJ_OWN) sm.own[n_own * 32 + lane] = (uint16_t)r; else bad |= 2;
                        ++n_own;
                    } else if (r == -2) {
                        bad |= 1;
                    }
                };
                // ---- (a) the full 32-entry blocks of every list, list by list: coalesced, no search
                const uint32_t f0 = l0 & ~31u, f1 = l1 & ~31u;
                __syncwarp();
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    uint32_t mask = __ballot_sync(0xffffffffu, (half ? f1 : f0) != 0u);
                    while (mask) {
                        const int j = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const uint32_t s = __shfl_sync(0xffffffffu, half ? s1 : s0, j);
                        const uint32_t f = __shfl_sync(0xffffffffu, half ? f1 : f0, j);
                        for (uint32_t t0 = 0; t0 < f; t0 += 128) {   // up to four loads in flight per lane
                            uint2 e[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (t0 + u * 32 < f) e[u] = p.post[s + t0 + u * 32 + lane];
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (t0 + u * 32 < f) put(e[u]);
                        }
                    }
                    if (!two) break;
                }
                // ---- (b) the remainders (< 32 entries per list), flattened: prefix sums + a search per posting
                const uint32_t r0 = l0 - f0, r1 = l1 - f1;
                uint32_t i0 = r0, i1 = r1;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t0 = __shfl_up_sync(0xffffffffu, i0, o), t1 = __shfl_up_sync(0xffffffffu, i1, o);
                    if (lane >= o) {
                        i0 += t0;
                        i1 += t1;
                    }
                }
                const uint32_t tot0 = __shfl_sync(0xffffffffu, i0, 31), tot1 = __shfl_sync(0xffffffffu, i1, 31);
                const uint32_t n_rem = tot0 + tot1;
                if (n_rem != 0) {
                    sm.start[lane] = s0 + f0;
                    sm.pref[lane] = i0 - r0;
                    sm.start[lane + 32] = s1 + f1;
                    sm.pref[lane + 32] = tot0 + i1 - r1;
                    __syncwarp();
                    const int n_lists = two ? PJ_IDS : 32;
                    for (uint32_t h0 = 0; h0 < n_rem; h0 += 32) {
                        const uint32_t h = h0 + lane;
                        if (h < n_rem) {
                            int lo = 0, hi = n_lists;   // last list j with pref[j] <= h (empty lists share a prefix value)
                            while (hi - lo > 1) {
                                const int mid = (lo + hi) >> 1;
                                if (sm.pref[mid] <= h) lo = mid; else hi = mid;
                            }
                            put(p.post[sm.start[lo] + (h - sm.pref[lo])]);
                        }
                    }
                }
                __syncwarp();
                const uint32_t any_bad = __reduce_or_sync(0xffffffffu, (uint32_t)bad);
                if (any_bad & 1u) {   // a hot window overflowed the table
                    failed = true;
                    break;
                }
                // ---- candidates -> sorted top-K list.  A candidate is a slot claimed in this pass: lane L looks at its own
                // claims (or, if some lane claimed more than PJ_OWN slots, at slots L, L + 32, ... of the whole table)
                const bool scan_all = (any_bad & 2u) != 0u;
                const int n_cand = scan_all ? PJ_T / 32 : __reduce_max_sync(0xffffffffu, n_own);
                auto cand = [&](int i) {
                    unsigned long long s = PJ_EMPTY;
                    if (scan_all) s = sm.tab[i * 32 + lane];
                    else if (i < n_own) s = sm.tab[sm.own[i * 32 + lane]];
                    if (s == PJ_EMPTY) return PEntry::worst();
                    const uint32_t lw = (uint32_t)s, cnt = lw & 0xffu, card = lw >> 8;
                    return PEntry{cnt, cq + card - cnt, (int32_t)(p.pool_base + (int64_t)(uint32_t)(s >> 32))};
                };
                int32_t seeded = R4D_IDX_NONE;
                if (!have_list) {
                    // empty list: every lane finds the best of its candidates, a bitonic sort ranks the 32 lane-bests and the
                    // first k of them seed the list (the others cannot be in the top k); the rest is inserted below only
                    // if it beats the k-th
                    PEntry lb = PEntry::worst();
                    for (int i = 0; i < n_cand; ++i) {
                        const PEntry c = cand(i);
                        if (PEntry::better(c, lb)) lb = c;
                    }
                    PEntry v = lb;
#pragma unroll
                    for (int k2 = 2; k2 <= 32; k2 <<= 1)
#pragma unroll
                        for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
                            const PEntry o = v.shfl_xor(j2);
                            const bool want_better = ((lane & j2) == 0) == ((lane & k2) == 0);
                            if (PEntry::better(o, v) == want_better && o.idx != v.idx) v = o;
                        }
                    if (lane < p.k) tk.mine = v;
                    tk.refresh_kth();
                    seeded = lb.idx;
                    have_list = true;
                }
                for (int i = 0; i < n_cand; ++i) {
                    const PEntry c = cand(i);
                    uint32_t m = __ballot_sync(0xffffffffu, c.idx != seeded && PEntry::better(c, tk.kth));
                    while (m) {
                        const int src = __ffs(m) - 1;
                        m &= m - 1;
                        tk.insert(c.shfl(src));
                    }
                }
                __syncwarp();
            }
            if (failed) {
                hand_over(q);
                continue;
            }
            pj_finish(tk, p, q, cq, packed, staged ? sm.obuf : nullptr, qi, n_here);
        }
        if (staged) {
            // rows of queries handed to the heavy kernel hold stale words here; that kernel runs afterwards and rewrites them
            __syncwarp();
            pj_flush_chunk(p, sm.obuf, q0, n_here, packed);
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------- light kernel, register-resident variant
// For label-like sets (a couple of hundred postings per query) the hash table of postings_light_kernel is mostly
// overhead: a posting's row almost never repeats (two ids of a query rarely share a pool row), yet every posting pays a
// CAS probe loop, a claimed-slot list and a table read-back.  Here a pass keeps its <= 256 postings IN REGISTERS, eight
// per lane ({row, card << 8 | count}), and finds the rare repeats with a per-warp 8 192-bit filter in shared memory: a
// posting whose filter bit was already set MAY repeat an earlier row; only those (a handful per query, false positives
// included) are resolved, by broadcasting the row and letting every lane compare it with the live rows it holds — the
// holder absorbs the count, the repeat dies.  Counts are exact; everything after that (candidate -> sorted top-K list,
// passes over row windows for longer lists, hand-over to the heavy kernel) is the same as in postings_light_kernel.
// Postings a lane holds per pass: 5 (160 per pass) keeps the unrolled per-slot code short — the kernel is instruction-fetch
// sensitive: 8 slots ran 17 % slower on uniform ids — while 8 (256 per pass) serves pools with hot ids, where a single
// row window of two hot lists would overflow the smaller pass (such queries then take the slow hand-over chain).  The
// index header knows the largest (id, window) bucket; the kernel picks its body from it, uniformly for the whole grid.
constexpr int PR_SLOTS_SMALL = 5, PR_SLOTS_LARGE = 8;
constexpr int PR_BM_WORDS = 256;          // 8 192-bit repeat filter per warp
constexpr int PR_WARPS = 8;

struct PRWarpSmem {
    uint32_t bm[PR_BM_WORDS];
    uint32_t start[PJ_IDS];
    uint32_t pref[PJ_IDS + 1];
    int32_t ids[PJ_IDS];
    uint32_t obuf[3 * PJ_CHUNK * PJ_OBUF_K];
    uint32_t pad[3];
};

template <int PR_SLOTS, bool PACKED>
__device__ __forceinline__ void postings_reg_body(const PJParams& p) {
    constexpr bool packed = PACKED;   // packed results are a kernel variant of their own: the plain one pays nothing for them
    constexpr int PR_CAP = PR_SLOTS * 32;            // postings per pass
    constexpr int PR_PLAN = PR_SLOTS * 27;           // postings PLANNED per pass (row windows are not perfectly even)
    extern __shared__ __align__(16) uint8_t pj_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    PRWarpSmem& sm = reinterpret_cast<PRWarpSmem*>(pj_smem)[warp];
    auto clear_filter = [&]() {
        uint4* b4 = reinterpret_cast<uint4*>(sm.bm);
#pragma unroll
        for (int i = 0; i < PR_BM_WORDS / 4 / 32; ++i) b4[i * 32 + lane] = make_uint4(0u, 0u, 0u, 0u);
    };
    auto hand_over = [&](int64_t q) {   // the heavy kernel serves this query
        if (lane == 0) p.hand_list[atomicAdd(p.hand_count, 1u)] = (uint32_t)q;
    };
    // later stage of a chain: the queries are the ones the head kernel handed over, one per grab
    const bool listed = p.in_list != nullptr;
    const int64_t n_work = listed ? (int64_t)*p.in_count : p.nq;
    if (n_work == 0) return;   // (an empty list: no filter to clear, no grab)
    clear_filter();
    bool filter_clean = true;
    const int chunk = listed ? 1 : p.chunk;
    for (;;) {
        int64_t q0 = 0;
        if (lane == 0) q0 = (int64_t)atomicAdd(p.work, (uint32_t)chunk);
        q0 = __shfl_sync(0xffffffffu, q0, 0);
        if (q0 >= n_work) break;
        const int64_t q_first = listed ? (int64_t)p.in_list[q0] : q0;
        int64_t my_off = 0;
        if (lane <= chunk && q_first + lane <= p.nq) my_off = p.q_off[q_first + lane];
        const int n_here = (int)min((int64_t)chunk, n_work - q0);
        const bool staged = !listed && p.peers.world == 0 && p.k <= PJ_OBUF_K;   // see postings_light_kernel
        for (int qi = 0; qi < n_here; ++qi) {
            const int64_t q = q_first + qi;
            const int64_t beg = __shfl_sync(0xffffffffu, my_off, qi), end = __shfl_sync(0xffffffffu, my_off, qi + 1);
            const int64_t m_raw = end - beg;
            if (m_raw > 32) {   // more ids than lanes: the hash-table kernel (second stage) holds two ids per lane
                hand_over(q);
                continue;
            }
            // ---- the query's distinct ids, at most two per lane (set semantics: duplicates collapse)
            int32_t id0 = -1, id1 = -1;
            if (lane < m_raw) id0 = p.q_ids[beg + lane];
            if (lane + 32 < m_raw) id1 = p.q_ids[beg + 32 + lane];
            if (id0 < 0 || id0 >= p.n_bits) id0 = -1;
            if (id1 < 0 || id1 >= p.n_bits) id1 = -1;
            if (m_raw <= 32) {
                const uint32_t same = __match_any_sync(0xffffffffu, id0);
                if (id0 >= 0 && (__ffs(same) - 1) != lane) id0 = -1;
            } else {
                sm.ids[lane] = id0;
                sm.ids[lane + 32] = id1;
                __syncwarp();
                bool d0 = false, d1 = false;
                for (int t = 0; t < (int)m_raw; ++t) {
                    const int32_t v = sm.ids[t];
                    d0 |= (t < lane) && (v == id0);
                    d1 |= (t < lane + 32) && (v == id1);
                }
                if (d0) id0 = -1;
                if (d1) id1 = -1;
                __syncwarp();
            }
            const uint32_t cq = __popc(__ballot_sync(0xffffffffu, id0 >= 0)) + __popc(__ballot_sync(0xffffffffu, id1 >= 0));
            constexpr bool two = false;    // (this kernel keeps one id per lane; the shared code below folds away)
            const bool few = m_raw <= 4;   // warp-uniform: the lists are found by three compares instead of a search
            // ---- a query of ONE id needs no join: every row of the id's list scores 1 / |pool set|, and the index holds
            // the head of that list in (|pool set| asc, row asc) = (score desc, index asc) order
            if (cq == 1u && p.best != nullptr && p.k + (p.zero_diag ? 1 : 0) <= PJ_BEST) {
                const uint32_t has = __ballot_sync(0xffffffffu, id0 >= 0) | 0u;
                const int32_t id = has ? __shfl_sync(0xffffffffu, id0, __ffs(has) - 1)
                                       : __shfl_sync(0xffffffffu, id1, __ffs(__ballot_sync(0xffffffffu, id1 >= 0)) - 1);
                const uint2 e = p.best[(int64_t)id * PJ_BEST + lane];
                const int64_t diag_row1 = p.query_base + q - p.pool_base;
                const bool is_diag = p.zero_diag != 0 && e.x != 0xffffffffu && (int64_t)e.x == diag_row1;
                const uint32_t dm = __ballot_sync(0xffffffffu, is_diag);
                const int d_at = dm ? __ffs(dm) - 1 : 32;                       // the forced-zero row, if it is listed
                const int from = lane < d_at ? lane : lane + 1;                 // close the gap it leaves
                const uint32_t row = __shfl_sync(0xffffffffu, e.x, from & 31), card = __shfl_sync(0xffffffffu, e.y, from & 31);
                const bool live = from < 32 && row != 0xffffffffu;
                const int n_live = __popc(__ballot_sync(0xffffffffu, live));
                if (n_live >= p.k) {   // (a shorter list falls through to the general path, which also adds the fillers)
                    WarpTopK<PEntry> tk1;
                    tk1.init(p.k);
                    if (lane < p.k) tk1.mine = PEntry{1u, card, (int32_t)(p.pool_base + (int64_t)row)};
                    tk1.refresh_kth();
                    pj_finish(tk1, p, q, cq, packed, staged ? sm.obuf : nullptr, qi, n_here);
                    continue;
                }
            }
            uint32_t s0 = 0, e0 = 0, s1 = 0, e1 = 0;
            if (id0 >= 0) {
                s0 = p.off[(int64_t)id0 * p.n_win];
                e0 = p.off[(int64_t)(id0 + 1) * p.n_win];
            }
            if (two && id1 >= 0) {
                s1 = p.off[(int64_t)id1 * p.n_win];
                e1 = p.off[(int64_t)(id1 + 1) * p.n_win];
            }
            const uint32_t hits = __reduce_add_sync(0xffffffffu, (e0 - s0) + (e1 - s1));
            // One pass holds PR_CAP postings.  A query with more walks the pool's row windows, `wpp` windows per pass
            // (planned for PR_PLAN postings: windows are not perfectly even); a pass that still overflows is halved until
            // it fits — only a SINGLE window with more than PR_CAP postings sends the query to the heavy kernel.
            const bool single = hits <= (uint32_t)PR_CAP;
            int wpp = p.n_win;
            if (!single) {
                wpp = (int)(((uint64_t)PR_PLAN * (uint64_t)p.n_win) / (uint64_t)hits);
                if (wpp < 1) wpp = 1;
            }
            if (!single && (uint64_t)hits > (uint64_t)PR_CAP * (uint64_t)p.n_win) {   // even one window per pass cannot fit
                hand_over(q);
                continue;
            }
            WarpTopK<PEntry> tk;
            tk.init(p.k);
            bool failed = false, have_list = false;   // have_list: the sorted list holds entries of an earlier pass
            const bool diag_on = p.zero_diag != 0;
            const int64_t diag_row = p.query_base + q - p.pool_base;   // pool row forced to score 0
            for (int w = 0; w < p.n_win;) {
                int wn = single ? p.n_win : min(w + wpp, p.n_win);
                uint32_t l0, l1, i0, i1, tot0, tot1, n_post;
                for (;;) {
                    if (!single) {   // this pass: windows [w, wn) of every list
                        if (id0 >= 0) {
                            s0 = p.off[(int64_t)id0 * p.n_win + w];
                            e0 = p.off[(int64_t)id0 * p.n_win + wn];
                        }
                        if (two && id1 >= 0) {
                            s1 = p.off[(int64_t)id1 * p.n_win + w];
                            e1 = p.off[(int64_t)id1 * p.n_win + wn];
                        }
                    }
                    // ---- the pass's postings as ONE sequence: posting g of the sequence goes to lane g % 32, slot g / 32
                    l0 = e0 - s0;
                    l1 = two ? e1 - s1 : 0u;
                    i0 = l0;
                    i1 = l1;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t t0 = __shfl_up_sync(0xffffffffu, i0, o), t1 = __shfl_up_sync(0xffffffffu, i1, o);
                        if (lane >= o) {
                            i0 += t0;
                            i1 += t1;
                        }
                    }
                    tot0 = __shfl_sync(0xffffffffu, i0, 31);
                    tot1 = __shfl_sync(0xffffffffu, i1, 31);
                    n_post = tot0 + tot1;
                    if (n_post <= (uint32_t)PR_CAP || wn - w <= 1) break;
                    wn = w + ((wn - w) >> 1);   // too many postings for the registers of a pass: take half the windows
                }
                w = wn;
                if (n_post == 0u) continue;
                if (n_post > (uint32_t)PR_CAP) {   // one hot window: more postings than a pass holds
                    failed = true;
                    break;
                }
                const uint32_t ex0 = i0 - l0, ex1 = tot0 + i1 - l1;   // first sequence position of the lane's lists
                uint32_t b1 = 0, b2 = 0, b3 = 0, st0 = 0, st1 = 0, st2 = 0, st3 = 0;
                if (few) {
                    b1 = __shfl_sync(0xffffffffu, ex0, 1);
                    b2 = __shfl_sync(0xffffffffu, ex0, 2);
                    b3 = __shfl_sync(0xffffffffu, ex0, 3);
                    st0 = __shfl_sync(0xffffffffu, s0, 0);
                    st1 = __shfl_sync(0xffffffffu, s0, 1);
                    st2 = __shfl_sync(0xffffffffu, s0, 2);
                    st3 = __shfl_sync(0xffffffffu, s0, 3);
                } else {
                    sm.start[lane] = s0;
                    sm.pref[lane] = ex0;
                    sm.start[lane + 32] = s1;
                    sm.pref[lane + 32] = ex1;
                }
                if (!filter_clean) clear_filter();
                filter_clean = false;
                __syncwarp();
                const int n_it = (int)((n_post + 31u) >> 5);
                const int n_lists = two ? PJ_IDS : 32;
                uint2 ent[PR_SLOTS];
#pragma unroll
                for (int sl = 0; sl < PR_SLOTS; ++sl) {   // all loads of the pass in flight
                    ent[sl] = make_uint2(0xffffffffu, 0u);
                    const uint32_t g = (uint32_t)sl * 32u + (uint32_t)lane;
                    if (sl < n_it && g < n_post) {
                        uint32_t at;
                        if (few) {
                            // lists past the query's ids are empty and start at n_post: an empty list shares its start
                            // with its successor, so counting "starts <= g" lands on the list that holds g
                            const int li = (g >= b1) + (g >= b2) + (g >= b3);
                            const uint32_t base = li == 0 ? 0u : (li == 1 ? b1 : (li == 2 ? b2 : b3));
                            const uint32_t st = li == 0 ? st0 : (li == 1 ? st1 : (li == 2 ? st2 : st3));
                            at = st + (g - base);
                        } else {
                            int lo = 0, hi = n_lists;   // last list j with pref[j] <= g (empty lists share a prefix value)
                            while (hi - lo > 1) {
                                const int mid = (lo + hi) >> 1;
                                if (sm.pref[mid] <= g) lo = mid; else hi = mid;
                            }
                            at = sm.start[lo] + (g - sm.pref[lo]);
                        }
                        ent[sl] = p.post[at];
                    }
                }
                uint32_t rrow[PR_SLOTS], rcc[PR_SLOTS];   // row; card << 8 | count (count 0: no posting here / merged away)
                uint32_t flags = 0u;                      // bit sl: the posting may repeat an earlier row
#pragma unroll
                for (int sl = 0; sl < PR_SLOTS; ++sl) {
                    rrow[sl] = ent[sl].x;
                    rcc[sl] = 0u;
                    const uint32_t g = (uint32_t)sl * 32u + (uint32_t)lane;
                    if (sl < n_it && g < n_post && !(diag_on && (int64_t)ent[sl].x == diag_row)) {
                        rcc[sl] = (ent[sl].y << 8) | 1u;
                        const uint32_t h = (ent[sl].x * 2654435761u) >> 19;   // 13 bits
                        const uint32_t bit = 1u << (h & 31u);
                        if (atomicOr(&sm.bm[h >> 5], bit) & bit) flags |= 1u << sl;
                    }
                }
                // Only the cheap per-slot work above and below is unrolled; everything with a large body (the resolution
                // loop, the list insertions) exists ONCE and picks its slot with an 8-way select — the fully unrolled
                // version was instruction-fetch bound (ncu: 7.3 of 12 stall cycles "no instruction").
                auto sel8 = [&](const uint32_t (&a)[PR_SLOTS], int i) {
                    uint32_t v = a[0];
#pragma unroll
                    for (int t = 1; t < PR_SLOTS; ++t) v = (i == t) ? a[t] : v;
                    return v;
                };
                // ---- resolve the possible repeats (rare): the live holder of the same row absorbs the count
                if (__ballot_sync(0xffffffffu, flags != 0u)) {
#pragma unroll 1
                    for (int sl = 0; sl < n_it; ++sl) {
                        uint32_t m = __ballot_sync(0xffffffffu, (flags >> sl) & 1u);
                        if (!m) continue;
                        const uint32_t my_row = sel8(rrow, sl);
                        while (m) {
                            const int src = __ffs(m) - 1;
                            m &= m - 1;
                            const uint32_t drow = __shfl_sync(0xffffffffu, my_row, src);
                            const uint32_t dcnt = __shfl_sync(0xffffffffu, sel8(rcc, sl), src) & 0xffu;   // counts move: re-read
                            int hs = -1;
#pragma unroll
                            for (int t = 0; t < PR_SLOTS; ++t)
                                if (rrow[t] == drow && (rcc[t] & 0xffu) != 0u && !(t == sl && lane == src)) hs = t;
                            const uint32_t hm = __ballot_sync(0xffffffffu, hs >= 0);
                            if (hm) {
                                const bool owner = lane == __ffs(hm) - 1;
#pragma unroll
                                for (int t = 0; t < PR_SLOTS; ++t) {
                                    if (owner && t == hs) rcc[t] += dcnt;
                                    if (lane == src && t == sl) rcc[t] &= ~0xffu;
                                }
                            }
                        }
                    }
                }
                // ---- candidates (live postings) -> sorted top-K list
                auto cand = [&](uint32_t row, uint32_t cc) {
                    const uint32_t cnt = cc & 0xffu, card = cc >> 8;
                    if (cnt == 0u) return PEntry::worst();
                    return PEntry{cnt, cq + card - cnt, (int32_t)(p.pool_base + (int64_t)row)};
                };
                int32_t seeded = R4D_IDX_NONE;
                if (!have_list) {
                    // empty list: every lane finds the best of its candidates, a bitonic sort ranks the 32 lane-bests and the
                    // first k of them seed the list (the others cannot be in the top k); the rest is inserted below only
                    // if it beats the k-th
                    PEntry lb = PEntry::worst();
#pragma unroll
                    for (int t = 0; t < PR_SLOTS; ++t) {
                        const PEntry c = cand(rrow[t], rcc[t]);   // slots past n_it hold count 0 = worst()
                        if (PEntry::better(c, lb)) lb = c;
                    }
                    PEntry v = lb;
#pragma unroll
                    for (int k2 = 2; k2 <= 32; k2 <<= 1)
#pragma unroll
                        for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
                            const PEntry o = v.shfl_xor(j2);
                            const bool want_better = ((lane & j2) == 0) == ((lane & k2) == 0);
                            if (PEntry::better(o, v) == want_better && o.idx != v.idx) v = o;
                        }
                    if (lane < p.k) tk.mine = v;
                    tk.refresh_kth();
                    seeded = lb.idx;
                    have_list = true;
                }
                uint32_t todo = 0u;   // my slots whose candidate still beats the k-th entry
#pragma unroll
                for (int t = 0; t < PR_SLOTS; ++t) {
                    const PEntry c = cand(rrow[t], rcc[t]);
                    if (c.idx != seeded && PEntry::better(c, tk.kth)) todo |= 1u << t;
                }
#pragma unroll 1
                for (;;) {
                    const uint32_t any = __ballot_sync(0xffffffffu, todo != 0u);
                    if (!any) break;
                    const int src = __ffs(any) - 1;
                    const int t = __ffs(todo) - 1;                  // meaningful in lane src only
                    const PEntry c = cand(sel8(rrow, t), sel8(rcc, t));
                    if (lane == src) todo &= todo - 1u;
                    tk.insert(c.shfl(src));                         // a candidate the list has outgrown is dropped inside
                }
                __syncwarp();
            }
            if (failed) {
                hand_over(q);
                continue;
            }
            pj_finish(tk, p, q, cq, packed, staged ? sm.obuf : nullptr, qi, n_here);
        }
        if (staged) {
            // rows of queries handed to the heavy kernel hold stale words here; that kernel runs afterwards and rewrites them
            __syncwarp();
            pj_flush_chunk(p, sm.obuf, q0, n_here, packed);
            __syncwarp();
        }
    }
}

// Two kernels, one of which exits at once (an idle launch costs ~3 us): each body gets its own register budget, and the
// short 5-slot body fits 5 CTAs per SM.
__device__ __forceinline__ bool pr_wants_large(const PJParams& p) {
    return 2u * p.hdr->max_bucket > (uint32_t)(PR_SLOTS_SMALL * 32);
}
template <bool PACKED>
__global__ void __launch_bounds__(PR_WARPS * 32, 6) postings_reg_kernel(const PJParams p) {
    if (!pr_wants_large(p)) postings_reg_body<PR_SLOTS_SMALL, PACKED>(p);
}
template <bool PACKED>
__global__ void __launch_bounds__(PR_WARPS * 32, 4) postings_reg_large_kernel(const PJParams p) {
    // (behind the head kernel this launch serves every pool: one stage for the few queries handed over, not two)
    if (p.in_list != nullptr || pr_wants_large(p)) postings_reg_body<PR_SLOTS_LARGE, PACKED>(p);
}

// ---------------------------------------------------------------------------- first stage for label-like sets: heads + repeat check
// A pool row that holds exactly ONE of the query's m ids scores 1 / (m + |pool set| - 1): among such rows the ranking is
// (|pool set| asc, row asc) — the order the per-id best lists are kept in.  So the top-K over the single-hit rows is the
// merge of the m list heads, and the join only has to find the rows that hold SEVERAL of the query's ids (rare: two ids of a
// query seldom share a pool row).  Per query:
//   1. stream the rows of all postings, list by list, through a per-warp 16 384-bit Bloom filter in shared memory (two bits
//      of one word per row: one ATOMS.OR); a posting whose bits were already set MAY repeat an earlier row — it is noted
//      (rarely: for a few hundred postings the false-positive rate is ~1e-3).  Nothing is kept in registers, there are no
//      passes over row windows, no candidate handling per posting;
//   2. every noted row is counted exactly: one lane per (row, id) pair looks the row up in the id's (row window) bucket;
//      rows with a count >= 2 are the multi-hit rows (with |pool set| from the posting);
//   3. the m heads (32 entries each, one coalesced load) lose the multi-hit rows and the forced-zero row, and are merged
//      by a 5-stage bitonic network per list (two sorted 16-entry lists -> sorted 32); the multi-hit rows are inserted
//      with their exact score.  A head that is left with fewer than k entries although its list is longer than the head
//      cannot vouch for the k best single-hit rows of its id: the query goes to the next stage.
// Queries this kernel cannot serve (more than 32 ids, more than PH_MAX_HITS postings, more than PH_MULTI multi-hit rows,
// a probed bucket with more than PH_MAX_BUCKET entries, a depleted head) are handed to the register kernel.  k <= 16.
constexpr int PH_BM_WORDS = 512;     // 16 384-bit repeat filter per warp
constexpr int PH_FLAG = 64;          // noted rows buffered (resolved in batches, whenever more than half of them are in use)
constexpr int PH_MULTI = 32;         // multi-hit rows per query
constexpr int PH_MAX_HITS = 4096;    // postings streamed per query
constexpr int PH_MAX_BUCKET = 64;    // bucket entries scanned per probe
constexpr int PH_K = 16;             // the merge network ranks two sorted 16-entry lists
constexpr int PH_WARPS = 8;
constexpr int PH_BIG_IDS = 8;         // queries with more ids are served first

struct PHWarpSmem {
    uint32_t bm[PH_BM_WORDS];
    int32_t ids[32];
    uint32_t frow[PH_FLAG];    // noted rows
    uint32_t fcnt[PH_FLAG];    // ids of the query each of them holds
    uint32_t fcard[PH_FLAG];   // |pool set|
    uint32_t mrow[PH_MULTI], mcnt[PH_MULTI], mcard[PH_MULTI];   // the multi-hit rows among them
    uint32_t obuf[3 * PJ_CHUNK * PJ_OBUF_K];
};

__global__ void __launch_bounds__(256) postings_big_scan_kernel(const int64_t* __restrict__ q_off, int64_t nq, int32_t big_ids,
                                                               uint32_t* __restrict__ list, uint32_t* __restrict__ count) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = q_off[q + 1] - q_off[q];
        if (m > (int64_t)big_ids && m <= 32) list[atomicAdd(count, 1u)] = (uint32_t)q;
    }
}

template <bool PACKED>
__global__ void __launch_bounds__(PH_WARPS * 32, 6) postings_head_kernel(const PJParams p) {
    constexpr bool packed = PACKED;
    extern __shared__ __align__(16) uint8_t pj_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    PHWarpSmem& sm = reinterpret_cast<PHWarpSmem*>(pj_smem)[warp];
    const uint32_t lt = (1u << lane) - 1u;
    constexpr uint32_t FULL = 0xffffffffu, NONE = 0xffffffffu;
    auto clear_filter = [&]() {
        uint4* b4 = reinterpret_cast<uint4*>(sm.bm);
#pragma unroll
        for (int i = 0; i < PH_BM_WORDS / 4 / 32; ++i) b4[i * 32 + lane] = make_uint4(0u, 0u, 0u, 0u);
    };
    auto hand_over = [&](int64_t q) {
        if (lane == 0) p.hand_list[atomicAdd(p.hand_count, 1u)] = (uint32_t)q;
    };
    clear_filter();
    bool filter_clean = true;
    // head entries are ranked as 32-bit keys |pool set| << key_shift | row
    const uint32_t key_shift = p.np > 1 ? 32u - (uint32_t)__clz((uint32_t)(p.np - 1)) : 1u;
    const uint32_t row_mask = (1u << key_shift) - 1u, card_lim = (1u << (32u - key_shift)) - 1u;
    const int chunk = p.chunk;
    const bool staged = p.peers.world == 0 && p.k <= PJ_OBUF_K;   // see postings_light_kernel
    for (;;) {
        int64_t q0 = 0;
        if (lane == 0) q0 = (int64_t)atomicAdd(p.work, (uint32_t)chunk);
        q0 = __shfl_sync(FULL, q0, 0);
        if (q0 >= p.nq) break;
        int64_t my_off = 0;
        if (lane <= chunk && q0 + lane <= p.nq) my_off = p.q_off[q0 + lane];
        const int n_here = (int)min((int64_t)chunk, p.nq - q0);
        uint32_t skip = 0u;
        for (int qi = 0; qi < n_here; ++qi) {
            const int64_t q = q0 + qi;
            const int64_t beg = __shfl_sync(FULL, my_off, qi), end = __shfl_sync(FULL, my_off, qi + 1);
            const int64_t m_raw = end - beg;
            if (m_raw > 32) {
                hand_over(q);
                continue;
            }
            if (p.big_list != nullptr && m_raw > (int64_t)p.big_ids) {   // served by postings_big_kernel
                skip |= 1u << qi;
                continue;
            }
            // ---- the query's distinct ids (set semantics: duplicates collapse), the j-th of them in lane j
            int32_t id = -1;
            if (lane < m_raw) id = p.q_ids[beg + lane];
            if (id < 0 || id >= p.n_bits) id = -1;
            {
                const uint32_t same = __match_any_sync(FULL, id);
                if (id >= 0 && (__ffs(same) - 1) != lane) id = -1;
            }
            const uint32_t act = __ballot_sync(FULL, id >= 0);
            const uint32_t cq = __popc(act);
            if (act & (act + 1u)) {   // not a prefix of the lanes yet
                __syncwarp();
                if (id >= 0) sm.ids[__popc(act & lt)] = id;
                __syncwarp();
                id = lane < (int)cq ? sm.ids[lane] : -1;
            }
            uint32_t s = 0, len = 0;
            if (id >= 0) {
                s = p.off[(int64_t)id * p.n_win];
                len = p.off[(int64_t)(id + 1) * p.n_win] - s;
            }
            const int64_t diag_row = p.query_base + q - p.pool_base;   // pool row forced to score 0
            const uint32_t diag = (p.zero_diag != 0 && diag_row >= 0 && diag_row < p.np) ? (uint32_t)diag_row : NONE;
            uint32_t n_flag = 0, n_multi = 0;
            if (cq >= 2u) {
                const uint32_t hits = __reduce_add_sync(FULL, len);
                if (hits > (uint32_t)PH_MAX_HITS) {
                    hand_over(q);
                    continue;
                }
                if (hits != 0u) {
                    if (!filter_clean) clear_filter();
                    filter_clean = false;
                    __syncwarp();
                }
                const uint32_t inv = ph_inv[cq];   // t / cq = umulhi(t, inv) for t < 2^16
                // ---- 2. the noted rows, counted exactly: lane <-> (row j, id i), the row is looked up in the id's (row
                // window) bucket; rows with a count >= 2 (never the forced-zero row) join the multi-hit list.  Runs
                // whenever the noted rows pile up and once at the end.  false: the query must be handed over.
                auto resolve = [&]() -> bool {
                    __syncwarp();
                    for (uint32_t j = lane; j < n_flag; j += 32u) sm.fcnt[j] = 0u;
                    __syncwarp();
                    const uint32_t n_probe = n_flag * cq;
                    bool bad = false;
                    for (uint32_t t0 = 0; t0 < n_probe; t0 += 32u) {
                        const uint32_t t = t0 + (uint32_t)lane;
                        const bool on = t < n_probe;
                        const uint32_t j = on ? __umulhi(t, inv) : 0u, i = on ? t - j * cq : 0u;
                        const int32_t pid = __shfl_sync(FULL, id, (int)i);
                        if (on) {
                            const uint32_t r = sm.frow[j];
                            const int64_t b = (int64_t)pid * p.n_win + (int64_t)(r >> p.win_shift);
                            const uint32_t bs = p.off[b], be = p.off[b + 1];
                            if (be - bs > (uint32_t)PH_MAX_BUCKET) {
                                bad = true;   // a hot bucket
                            } else {
                                for (uint32_t u = bs; u < be; ++u) {
                                    const uint2 v = p.post[u];
                                    if (v.x == r) {
                                        atomicAdd(&sm.fcnt[j], 1u);
                                        sm.fcard[j] = v.y;
                                        break;
                                    }
                                }
                            }
                        }
                    }
                    __syncwarp();
                    for (uint32_t j0 = 0; j0 < n_flag; j0 += 32u) {
                        const uint32_t j = j0 + (uint32_t)lane;
                        const bool multi = j < n_flag && sm.fcnt[j] >= 2u && sm.frow[j] != diag;
                        const uint32_t mm = __ballot_sync(FULL, multi);
                        if (multi) {
                            const uint32_t at = n_multi + __popc(mm & lt);
                            if (at < (uint32_t)PH_MULTI) {
                                sm.mrow[at] = sm.frow[j];
                                sm.mcnt[at] = sm.fcnt[j];
                                sm.mcard[at] = sm.fcard[j];
                            }
                        }
                        n_multi += __popc(mm);
                    }
                    n_flag = 0u;
                    __syncwarp();
                    return !__ballot_sync(FULL, bad) && n_multi <= (uint32_t)PH_MULTI;
                };
                // ---- 1. the rows of every list through the filter: list by list, lanes stride the list, four loads in flight
                bool ok = true;
                for (uint32_t i = 0; i < cq && ok; ++i) {
                    const uint32_t ls = __shfl_sync(FULL, s, (int)i), ll = __shfl_sync(FULL, len, (int)i);
                    const uint2* lp = p.post + ls;
                    for (uint32_t t0 = 0; t0 < ll && ok; t0 += 128u) {
                        uint32_t row[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t t = t0 + (uint32_t)j * 32u + (uint32_t)lane;
                            row[j] = t < ll ? lp[t].x : NONE;
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (t0 + (uint32_t)j * 32u >= ll) break;   // warp-uniform
                            bool fl = false;
                            if (row[j] != NONE) {
                                // two bits of one filter word per row (a blocked Bloom filter: one ATOMS)
                                const uint32_t h = row[j] * 2654435761u;
                                const uint32_t mask = (1u << ((h >> 18) & 31u)) | (1u << ((h >> 13) & 31u));
                                fl = (atomicOr(&sm.bm[h >> 23], mask) & mask) == mask;
                            }
                            const uint32_t fm = __ballot_sync(FULL, fl);
                            if (fm) {
                                const uint32_t slot = n_flag + __popc(fm & lt);
                                if (fl && slot < (uint32_t)PH_FLAG) sm.frow[slot] = row[j];
                                n_flag += __popc(fm);
                            }
                        }
                        // noted rows are resolved in batches; a group of 128 postings that notes more than the buffer has
                        // room for (a saturated filter, a pool full of repeats) sends the query to the next stage
                        if (n_flag > (uint32_t)PH_FLAG) ok = false;
                        else if (n_flag > (uint32_t)PH_FLAG / 2u) ok = resolve();
                    }
                }
                if (ok && n_flag != 0u) ok = resolve();
                if (!ok) {
                    hand_over(q);
                    continue;
                }
            }
            // ---- 3. heads: 32-bit keys (|pool set| << key_shift | row), ascending = best first; lanes 0..15 carry the
            // merged list.  (A |pool set| too large for the key — none among label-like sets — sends the query on.)
            uint32_t cur = NONE;
            bool depleted = false;
            for (uint32_t i = 0; i < cq; ++i) {
                const int32_t hid = __shfl_sync(FULL, id, (int)i);
                const uint32_t hlen = __shfl_sync(FULL, len, (int)i);
                const uint2 e = p.best[(int64_t)hid * PJ_BEST + lane];
                bool alive = e.x != NONE && e.x != diag;
                for (uint32_t j = 0; j < n_multi; ++j)   // multi-hit rows are ranked on their own below
                    if (sm.mrow[j] == e.x) alive = false;
                const uint32_t am = __ballot_sync(FULL, alive);
                if ((__popc(am) < p.k && hlen > (uint32_t)PJ_BEST) || __ballot_sync(FULL, alive && e.y >= card_lim)) {
                    depleted = true;
                    break;
                }
                uint32_t key = alive ? ((e.y << key_shift) | e.x) : NONE;
                if (am & (am + 1u)) {   // holes: the t-th survivor moves to lane t (the order is kept)
                    const uint32_t src = __fns(am, 0u, lane + 1);
                    const uint32_t moved = __shfl_sync(FULL, key, (int)(src & 31u));
                    key = src < 32u ? moved : NONE;
                }
                if (i == 0u) {
                    cur = key;
                } else {
                    const uint32_t rev = __shfl_sync(FULL, key, 31 - lane);   // lanes 16..31: the new list, worst first
                    uint32_t v = lane < PH_K ? cur : rev;
#pragma unroll
                    for (int j2 = 16; j2 > 0; j2 >>= 1) {
                        const uint32_t o = __shfl_xor_sync(FULL, v, j2);
                        v = ((lane & j2) == 0) ? min(v, o) : max(v, o);
                    }
                    cur = v;
                }
            }
            if (depleted) {
                hand_over(q);
                continue;
            }
            WarpTopK<PEntry> tk;
            tk.init(p.k);
            if (lane < p.k && cur != NONE)
                tk.mine = PEntry{1u, cq + (cur >> key_shift) - 1u, (int32_t)(p.pool_base + (int64_t)(cur & row_mask))};
            tk.refresh_kth();
            for (uint32_t j = 0; j < n_multi; ++j) {
                const uint32_t c = sm.mcnt[j];
                const PEntry cnd{c, cq + sm.mcard[j] - c, (int32_t)(p.pool_base + (int64_t)sm.mrow[j])};
                if (__ballot_sync(FULL, lane < p.k && tk.mine.idx == cnd.idx)) continue;   // noted twice
                tk.insert(cnd);
            }
            pj_finish(tk, p, q, cq, packed, staged ? sm.obuf : nullptr, qi, n_here);
        }
        if (staged) {
            // rows of handed-over queries hold stale words here; the later stages rewrite them
            __syncwarp();
            pj_flush_chunk(p, sm.obuf, q0, n_here, packed, skip);
            __syncwarp();
        }
        // (a handed-over query counts too: its rows are rewritten by a later stage, after this kernel's relays)
        if (p.relay_done != nullptr) pj_relay(p, q0, (uint32_t)n_here - (uint32_t)__popc(skip));
    }
}

// ---------------------------------------------------------------------------- the head kernel's algorithm, one CTA per LONG query
// A query of 25 ids streams ~2 800 postings; in one warp that is 25 dependent list walks, a per-warp filter that fills
// up (false positives grow with the square of the postings) and hundreds of bucket probes: ~100 us, as long as the whole
// head launch for the other 99 000 queries.  Here the eight warps of a CTA share the query: lists are dealt out to the
// warps, the Bloom filter is CTA wide (131 072 bits: almost no false positives), probes are spread over 256 threads,
// every warp merges the heads of its lists and warp 0 merges the eight partial rankings.  Exact for the same reasons.
constexpr int PB_THREADS = 256, PB_WARPS = PB_THREADS / 32;
constexpr int PB_BM_WORDS = 4096;    // 131 072-bit repeat filter per CTA
constexpr int PB_FLAG = 512;         // noted rows per query
constexpr int PB_MULTI = 64;         // multi-hit rows per query
constexpr int PB_MAX_HITS = 16384;   // postings streamed per query

struct PBSmem {
    uint32_t bm[PB_BM_WORDS];
    int32_t ids[32];
    uint32_t start[32], len[32];
    uint32_t frow[PB_FLAG], fcnt[PB_FLAG], fcard[PB_FLAG];
    uint32_t mrow[PB_MULTI], mcnt[PB_MULTI], mcard[PB_MULTI];
    uint32_t keys[PB_WARPS][PH_K];
    uint32_t cq, n_flag, n_multi, work, fail;
};

template <bool PACKED>
__global__ void __launch_bounds__(PB_THREADS) postings_big_kernel(const PJParams p) {
    constexpr bool packed = PACKED;
    __shared__ __align__(16) PBSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    constexpr uint32_t FULL = 0xffffffffu, NONE = 0xffffffffu;
    const uint32_t key_shift = p.np > 1 ? 32u - (uint32_t)__clz((uint32_t)(p.np - 1)) : 1u;
    const uint32_t row_mask = (1u << key_shift) - 1u, card_lim = (1u << (32u - key_shift)) - 1u;
    const uint32_t n_big = *p.big_count;
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            sm.work = atomicAdd(p.big_work, 1u);
            sm.n_flag = 0u;
            sm.n_multi = 0u;
            sm.fail = 0u;
        }
        for (int i = tid; i < PB_BM_WORDS / 4; i += PB_THREADS) reinterpret_cast<uint4*>(sm.bm)[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        if (sm.work >= n_big) break;
        const int64_t q = (int64_t)p.big_list[sm.work];
        if (warp == 0) {
            // ---- the query's distinct ids, the j-th of them in lane j (the list holds queries of 9 .. 32 ids)
            const int64_t beg = p.q_off[q], m_raw = p.q_off[q + 1] - beg;
            int32_t id = -1;
            if (lane < m_raw) id = p.q_ids[beg + lane];
            if (id < 0 || id >= p.n_bits) id = -1;
            const uint32_t same = __match_any_sync(FULL, id);
            if (id >= 0 && (__ffs(same) - 1) != lane) id = -1;
            const uint32_t act = __ballot_sync(FULL, id >= 0);
            const uint32_t cq = __popc(act);
            if (id >= 0) sm.ids[__popc(act & lt)] = id;
            __syncwarp();
            id = lane < (int)cq ? sm.ids[lane] : -1;
            uint32_t s = 0, len = 0;
            if (id >= 0) {
                s = p.off[(int64_t)id * p.n_win];
                len = p.off[(int64_t)(id + 1) * p.n_win] - s;
            }
            sm.start[lane] = s;
            sm.len[lane] = len;
            if (lane == 0) sm.cq = cq;
            if (__reduce_add_sync(FULL, len) > (uint32_t)PB_MAX_HITS && lane == 0) sm.fail = 1u;
        }
        __syncthreads();
        const uint32_t cq = sm.cq;
        const int64_t diag_row = p.query_base + q - p.pool_base;   // pool row forced to score 0
        const uint32_t diag = (p.zero_diag != 0 && diag_row >= 0 && diag_row < p.np) ? (uint32_t)diag_row : NONE;
        // ---- 1. rows through the CTA's filter: list i is walked by warp i % 8
        if (sm.fail == 0u) {
            for (uint32_t i = warp; i < cq; i += PB_WARPS) {
                const uint32_t ll = sm.len[i];
                const uint2* lp = p.post + sm.start[i];
                for (uint32_t t0 = 0; t0 < ll; t0 += 128u) {
                    uint32_t row[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t t = t0 + (uint32_t)j * 32u + (uint32_t)lane;
                        row[j] = t < ll ? lp[t].x : NONE;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (row[j] != NONE) {
                            const uint32_t h = row[j] * 2654435761u;
                            const uint32_t mask = (1u << ((h >> 15) & 31u)) | (1u << ((h >> 10) & 31u));
                            if ((atomicOr(&sm.bm[h >> 20], mask) & mask) == mask) {
                                const uint32_t slot = atomicAdd(&sm.n_flag, 1u);
                                if (slot < (uint32_t)PB_FLAG) sm.frow[slot] = row[j];
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
        const uint32_t n_flag = sm.n_flag;
        if (n_flag > (uint32_t)PB_FLAG && tid == 0) sm.fail = 1u;
        // ---- 2. the noted rows, counted exactly: thread <-> (row j, id i)
        for (uint32_t j = tid; j < n_flag && j < (uint32_t)PB_FLAG; j += PB_THREADS) sm.fcnt[j] = 0u;
        __syncthreads();
        if (sm.fail == 0u && n_flag != 0u) {
            const uint32_t n_probe = n_flag * cq, inv = ph_inv[cq];
            for (uint32_t t = tid; t < n_probe; t += PB_THREADS) {
                const uint32_t j = __umulhi(t, inv), i = t - j * cq;
                const uint32_t r = sm.frow[j];
                const int64_t b = (int64_t)sm.ids[i] * p.n_win + (int64_t)(r >> p.win_shift);
                const uint32_t bs = p.off[b], be = p.off[b + 1];
                if (be - bs > (uint32_t)PH_MAX_BUCKET) {
                    sm.fail = 1u;   // a hot bucket
                } else {
                    for (uint32_t u = bs; u < be; ++u) {
                        const uint2 v = p.post[u];
                        if (v.x == r) {
                            atomicAdd(&sm.fcnt[j], 1u);
                            sm.fcard[j] = v.y;
                            break;
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (sm.fail == 0u)
            for (uint32_t j = tid; j < n_flag; j += PB_THREADS)
                if (sm.fcnt[j] >= 2u && sm.frow[j] != diag) {
                    const uint32_t at = atomicAdd(&sm.n_multi, 1u);
                    if (at < (uint32_t)PB_MULTI) {
                        sm.mrow[at] = sm.frow[j];
                        sm.mcnt[at] = sm.fcnt[j];
                        sm.mcard[at] = sm.fcard[j];
                    }
                }
        __syncthreads();
        const uint32_t n_multi = sm.n_multi;
        if (n_multi > (uint32_t)PB_MULTI && tid == 0) sm.fail = 1u;
        // ---- 3. every warp merges the heads of its lists (see postings_head_kernel)
        uint32_t cur = NONE;
        if (sm.fail == 0u && n_multi <= (uint32_t)PB_MULTI) {
            bool have = false;
            for (uint32_t i = warp; i < cq; i += PB_WARPS) {
                const uint2 e = p.best[(int64_t)sm.ids[i] * PJ_BEST + lane];
                bool alive = e.x != NONE && e.x != diag;
                for (uint32_t j = 0; j < n_multi; ++j)
                    if (sm.mrow[j] == e.x) alive = false;
                const uint32_t am = __ballot_sync(FULL, alive);
                if ((__popc(am) < p.k && sm.len[i] > (uint32_t)PJ_BEST) || __ballot_sync(FULL, alive && e.y >= card_lim)) {
                    if (lane == 0) sm.fail = 1u;   // a depleted head / a |pool set| too large for the key
                    break;
                }
                uint32_t key = alive ? ((e.y << key_shift) | e.x) : NONE;
                if (am & (am + 1u)) {
                    const uint32_t src = __fns(am, 0u, lane + 1);
                    const uint32_t moved = __shfl_sync(FULL, key, (int)(src & 31u));
                    key = src < 32u ? moved : NONE;
                }
                if (!have) {
                    cur = key;
                    have = true;
                } else {
                    const uint32_t rev = __shfl_sync(FULL, key, 31 - lane);
                    uint32_t v = lane < PH_K ? cur : rev;
#pragma unroll
                    for (int j2 = 16; j2 > 0; j2 >>= 1) {
                        const uint32_t o = __shfl_xor_sync(FULL, v, j2);
                        v = ((lane & j2) == 0) ? min(v, o) : max(v, o);
                    }
                    cur = v;
                }
            }
        }
        if (lane < PH_K) sm.keys[warp][lane] = cur;
        __syncthreads();
        if (warp == 0) {
            if (sm.fail != 0u) {
                if (lane == 0) p.hand_list[atomicAdd(p.hand_count, 1u)] = (uint32_t)q;
            } else {
                cur = lane < PH_K ? sm.keys[0][lane] : NONE;
                for (int w = 1; w < PB_WARPS; ++w) {
                    const uint32_t key = lane < PH_K ? sm.keys[w][lane] : NONE;
                    const uint32_t rev = __shfl_sync(FULL, key, 31 - lane);
                    uint32_t v = lane < PH_K ? cur : rev;
#pragma unroll
                    for (int j2 = 16; j2 > 0; j2 >>= 1) {
                        const uint32_t o = __shfl_xor_sync(FULL, v, j2);
                        v = ((lane & j2) == 0) ? min(v, o) : max(v, o);
                    }
                    cur = v;
                }
                WarpTopK<PEntry> tk;
                tk.init(p.k);
                if (lane < p.k && cur != NONE)
                    tk.mine = PEntry{1u, cq + (cur >> key_shift) - 1u, (int32_t)(p.pool_base + (int64_t)(cur & row_mask))};
                tk.refresh_kth();
                for (uint32_t j = 0; j < n_multi; ++j) {
                    const uint32_t c = sm.mcnt[j];
                    const PEntry cnd{c, cq + sm.mcard[j] - c, (int32_t)(p.pool_base + (int64_t)sm.mrow[j])};
                    if (__ballot_sync(FULL, lane < p.k && tk.mine.idx == cnd.idx)) continue;   // noted twice
                    tk.insert(cnd);
                }
                pj_finish(tk, p, q, cq, packed);
            }
            // (a handed-over query counts too: a later stage rewrites its rows after the relays)
            if (p.relay_done != nullptr) pj_relay(p, q, 1u);
        }
    }
}

constexpr int PJ_HWIN_SHIFT = PJ_WIN_SHIFT_MAX;   // the heavy kernel walks the pool 32 768 rows at a time (several index windows)

struct PJHeavySmem {
    uint32_t cnt[(1 << PJ_HWIN_SHIFT) / 2];      // one 16-bit counter per row of the window
    uint32_t ubits[(PJ_MAX_BITS + 32) / 32];     // the query as a bitmap over the ids
    uint16_t uid[PJ_HEAVY_UID];                  // its distinct ids, enumerated (when they fit)
    uint32_t seg_start[PJ_HEAVY_UID];            // per enumerated id: its postings inside the current window
    uint32_t seg_pref[PJ_HEAVY_UID + 8];         // exclusive prefix sums of the segment lengths
    uint32_t lst_inter[PJ_HEAVY_THREADS / 32][32], lst_uni[PJ_HEAVY_THREADS / 32][32];
    int32_t lst_idx[PJ_HEAVY_THREADS / 32][32];
    uint32_t warp_tot[PJ_HEAVY_THREADS / 32];
    uint32_t n_uid, cq, work;
};

__global__ void __launch_bounds__(PJ_HEAVY_THREADS) postings_heavy_kernel(const PJParams p) {
    extern __shared__ __align__(16) uint8_t pj_smem[];
    PJHeavySmem& sm = *reinterpret_cast<PJHeavySmem*>(pj_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = PJ_HEAVY_THREADS / 32;
    static_assert(PJ_HEAVY_THREADS == SCAN_THREADS, "block_excl_scan is written for SCAN_THREADS threads");
    const uint32_t n_heavy = *p.in_count;
    if (n_heavy == 0u) return;
    constexpr int HW_ROWS = 1 << PJ_HWIN_SHIFT;
    const int wins_per_hw = 1 << (PJ_HWIN_SHIFT - p.win_shift);
    const int n_hwin = (p.n_win + wins_per_hw - 1) / wins_per_hw;
    const int n_uw = (p.n_bits + 31) / 32;
    for (int i = tid; i < HW_ROWS / 2; i += PJ_HEAVY_THREADS) sm.cnt[i] = 0u;
    for (;;) {
        __syncthreads();
        if (tid == 0) sm.work = atomicAdd(p.work, 1u);
        __syncthreads();
        const uint32_t hq = sm.work;
        if (hq >= n_heavy) break;
        const int64_t q = p.in_list[hq];
        const int64_t beg = p.q_off[q], end = p.q_off[q + 1];
        // ---- the query as a set: bitmap over the ids, |Q| = its popcount, distinct ids enumerated
        for (int i = tid; i < n_uw; i += PJ_HEAVY_THREADS) sm.ubits[i] = 0u;
        if (tid == 0) {
            sm.n_uid = 0u;
            sm.cq = 0u;
        }
        __syncthreads();
        for (int64_t e = beg + tid; e < end; e += PJ_HEAVY_THREADS) {
            const int32_t id = p.q_ids[e];
            if (id >= 0 && id < p.n_bits) {
                const uint32_t bit = 1u << (id & 31);
                const uint32_t old = atomicOr(&sm.ubits[id >> 5], bit);
                if (!(old & bit)) {
                    atomicAdd(&sm.cq, 1u);
                    const uint32_t at = atomicAdd(&sm.n_uid, 1u);
                    if (at < (uint32_t)PJ_HEAVY_UID) sm.uid[at] = (uint16_t)id;
                }
            }
        }
        __syncthreads();
        const uint32_t cq = sm.cq;
        const int n_uid = (int)sm.n_uid;
        const bool listed = n_uid <= PJ_HEAVY_UID;
        const int64_t diag_row = p.zero_diag ? p.query_base + q - p.pool_base : -1;
        WarpTopK<JEntry> tk;
        tk.init(p.k);
        for (int g = 0; g < n_hwin; ++g) {
            const int wlo = g * wins_per_hw, whi = min(wlo + wins_per_hw, p.n_win);
            auto count = [&](uint32_t row) {
                const uint32_t r = row & (uint32_t)(HW_ROWS - 1);
                atomicAdd(&sm.cnt[r >> 1], 1u << (16 * (r & 1)));
            };
            if (listed) {
                // ---- the ids' postings inside this window: segments, their prefix sums, then one thread per posting
                constexpr int PER = PJ_HEAVY_UID / PJ_HEAVY_THREADS;
                uint32_t len[PER], sum = 0;
#pragma unroll
                for (int j = 0; j < PER; ++j) {
                    const int i = tid * PER + j;
                    len[j] = 0u;
                    if (i < n_uid) {
                        const int64_t b = (int64_t)sm.uid[i] * p.n_win;
                        const uint32_t s = p.off[b + wlo];
                        sm.seg_start[i] = s;
                        len[j] = p.off[b + whi] - s;
                    }
                    sum += len[j];
                }
                uint32_t total;
                uint32_t run = block_excl_scan(sum, sm.warp_tot, total);
#pragma unroll
                for (int j = 0; j < PER; ++j) {
                    const int i = tid * PER + j;
                    if (i < n_uid) sm.seg_pref[i] = run;
                    run += len[j];
                }
                __syncthreads();
                for (uint32_t h = tid; h < total; h += PJ_HEAVY_THREADS) {
                    int lo = 0, hi = n_uid;   // last segment with pref <= h (empty segments share a prefix value)
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (sm.seg_pref[mid] <= h) lo = mid; else hi = mid;
                    }
                    count(p.post[sm.seg_start[lo] + (h - sm.seg_pref[lo])].x);
                }
            } else {
                // more distinct ids than the list holds: one warp per set bit of the bitmap, lanes stride its bucket
                for (int wi = warp; wi < n_uw; wi += NW) {
                    uint32_t x = sm.ubits[wi];
                    while (x) {
                        const int b = __ffs(x) - 1;
                        x &= x - 1;
                        const int64_t bk = (int64_t)(wi * 32 + b) * p.n_win;
                        const uint32_t s = p.off[bk + wlo], e = p.off[bk + whi];
                        for (uint32_t t = s + lane; t < e; t += 32) count(p.post[t].x);
                    }
                }
            }
            __syncthreads();
            // ---- non-zero counters are the candidates of this window (the counters are cleared on the way)
            const int64_t row0 = (int64_t)g << PJ_HWIN_SHIFT;
            uint4* c4 = reinterpret_cast<uint4*>(sm.cnt);
            for (int i0 = 0; i0 < HW_ROWS / 8; i0 += PJ_HEAVY_THREADS) {
                const int i = i0 + tid;
                if (row0 + (int64_t)i0 * 8 >= p.np) break;
                uint4 v = c4[i];
                const bool nz = (v.x | v.y | v.z | v.w) != 0u;
                if (!__ballot_sync(0xffffffffu, nz)) continue;
                if (nz) c4[i] = make_uint4(0u, 0u, 0u, 0u);
                const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
                // a candidate with count c scores at most c / |Q| (|P| >= c): only those that could still enter the list
                // fetch |P|; the (up to eight) fetches of a thread are issued together
                JEntry cand[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t c = (wv[j >> 1] >> (16 * (j & 1))) & 0xffffu;
                    const int64_t row = row0 + (int64_t)i * 8 + j;
                    cand[j] = JEntry::worst();
                    if (c != 0u && row != diag_row && row < p.np &&
                        (uint64_t)c * tk.kth.uni >= (uint64_t)tk.kth.inter * cq)
                        cand[j] = JEntry{c, cq + p.pcard[row] - c, (int32_t)(p.pool_base + row)};
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint32_t m = __ballot_sync(0xffffffffu, cand[j].inter != 0u && JEntry::better(cand[j], tk.kth));
                    while (m) {
                        const int src = __ffs(m) - 1;
                        m &= m - 1;
                        tk.insert(cand[j].shfl(src));
                    }
                }
            }
            __syncthreads();
        }
        // ---- merge the warps' lists: warp 0 inserts the others' entries
        sm.lst_inter[warp][lane] = tk.mine.inter;
        sm.lst_uni[warp][lane] = tk.mine.uni;
        sm.lst_idx[warp][lane] = tk.mine.idx;
        __syncthreads();
        if (warp == 0) {
            for (int w2 = 1; w2 < NW; ++w2) {
                const JEntry c{sm.lst_inter[w2][lane], sm.lst_uni[w2][lane], sm.lst_idx[w2][lane]};
                uint32_t m = __ballot_sync(0xffffffffu, lane < p.k && c.inter != 0u && JEntry::better(c, tk.kth));
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    tk.insert(c.shfl(src));
                }
            }
            pj_finish(tk, p, q, cq, p.out_qcard != nullptr);
        }
    }
}

}  // namespace r4d

extern "C" {

size_t r4d_postings_index_bytes(int64_t np, int32_t n_bits, int64_t nnz) {
    const r4d::PostingsLayout L = r4d::postings_layout(np, n_bits, nnz);
    return L.ok ? L.total : 0;
}

size_t r4d_postings_build_workspace_bytes(int64_t np, int32_t n_bits) {
    const r4d::PostingsLayout L = r4d::postings_layout(np, n_bits, 0);
    if (!L.ok) return 0;
    const size_t tiles = (size_t)((L.n_buckets + r4d::SCAN_TILE - 1) / r4d::SCAN_TILE);
    return ((size_t)L.n_buckets * 4 + 255) / 256 * 256 + (tiles + 1) * 4 + 256;
}

int r4d_postings_build(const uint32_t* pbits, const uint32_t* pcard, int64_t np, int32_t n_bits, int32_t pitch_words,
                       int64_t nnz, void* index, size_t index_bytes, void* workspace, size_t workspace_bytes,
                       r4d_stream_t stream) {
    using namespace r4d;
    const PostingsLayout L = postings_layout(np, n_bits, nnz);
    R4D_REQUIRE(L.ok, "postings_build: unsupported shape (np=%lld, n_bits=%d [max %d], nnz=%lld)", (long long)np, n_bits,
                PJ_MAX_BITS, (long long)nnz);
    const int32_t words = (n_bits + 31) / 32;
    R4D_REQUIRE(pitch_words >= words, "postings_build: pitch_words=%d < words=%d", pitch_words, words);
    R4D_REQUIRE(index && (reinterpret_cast<uintptr_t>(index) & 255) == 0, "postings_build: index must be 256-byte aligned");
    R4D_REQUIRE(np == 0 || (pbits && pcard), "postings_build: null pointer");
    if (index_bytes < L.total) {
        set_error("postings_build: index %zu B < required %zu B", index_bytes, L.total);
        return R4D_E_WORKSPACE;
    }
    const size_t need_ws = r4d_postings_build_workspace_bytes(np, n_bits);
    if (!workspace || workspace_bytes < need_ws) {
        set_error("postings_build: workspace %zu B < required %zu B", workspace_bytes, need_ws);
        return R4D_E_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    uint8_t* base = reinterpret_cast<uint8_t*>(index);
    PostingsHeader* hdr = reinterpret_cast<PostingsHeader*>(base);
    uint32_t* off = reinterpret_cast<uint32_t*>(base + L.off_at);
    uint2* post = reinterpret_cast<uint2*>(base + L.post_at);
    uint32_t* cursor = reinterpret_cast<uint32_t*>(workspace);
    uint32_t* tile_sum = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(workspace) + ((size_t)L.n_buckets * 4 + 255) / 256 * 256);
    PostingsHeader h{};
    h.magic = PJ_MAGIC;
    h.status = 0;
    h.np = np;
    h.nnz_cap = nnz;
    h.n_bits = n_bits;
    h.win_shift = L.win_shift;
    h.n_win = L.n_win;
    R4D_CUDA(cudaMemsetAsync(base, 0, L.post_at, st));   // header + offsets
    postings_header_kernel<<<1, 1, 0, st>>>(hdr, h); note_launch();
    R4D_CUDA(cudaMemsetAsync(cursor, 0, (size_t)L.n_buckets * 4, st));
    if (np == 0) {   // empty pool: every best list is empty
        R4D_CUDA(cudaMemsetAsync(base + L.best_at, 0xff, (size_t)n_bits * PJ_BEST * 8, st));
        return R4D_OK;
    }
    int64_t blocks = (np + 7) / 8;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    postings_scan_rows_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(pbits, pcard, np, words, pitch_words, n_bits, L.win_shift,
                                                                     L.n_win, off, cursor, post, nnz, hdr);
    note_launch();
    const int64_t n_scan = L.n_buckets;   // inclusive scan of off[1 .. n_buckets]
    const int64_t tiles = (n_scan + SCAN_TILE - 1) / SCAN_TILE;
    scan_tile_sums_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(off + 1, n_scan, tile_sum); note_launch();
    scan_sums_kernel<<<1, SCAN_THREADS, 0, st>>>(tile_sum, tiles); note_launch();
    scan_apply_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(off + 1, n_scan, tile_sum); note_launch();
    postings_scan_rows_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(pbits, pcard, np, words, pitch_words, n_bits, L.win_shift,
                                                                    L.n_win, off, cursor, post, nnz, hdr);
    note_launch();
    {
        uint2* best = reinterpret_cast<uint2*>(base + L.best_at);
        int64_t bb = ((int64_t)n_bits + 7) / 8;
        if (bb > cap) bb = cap;
        postings_best_kernel<<<(unsigned)bb, 256, 0, st>>>(off, post, n_bits, L.n_win, best); note_launch();
        postings_max_bucket_kernel<<<(unsigned)cap, 256, 0, st>>>(off, L.n_buckets, hdr); note_launch();
    }
    R4D_CUDA(cudaGetLastError());
    return R4D_OK;
}

size_t r4d_jaccard_topk_postings_workspace_bytes(int64_t nq) {
    return 256 + (size_t)(nq > 0 ? nq : 0) * 8;   // counters + two hand-over lists
}

static size_t relay_done_bytes(int64_t nq) { return (((size_t)((nq + r4d::PJ_RELAY - 1) / r4d::PJ_RELAY) * 4 + 255) / 256) * 256; }
static size_t relay_plane_bytes(int64_t nq, int32_t k) { return (((size_t)nq * (size_t)k * 4 + 255) / 256) * 256; }

size_t r4d_jaccard_topk_postings_relay_bytes(int64_t nq, int32_t k) {
    if (nq <= 0 || k <= 0) return 0;
    return 256 + relay_done_bytes(nq) + 2 * relay_plane_bytes(nq, k) + (((size_t)nq * 4 + 255) / 256) * 256;
}

static int postings_topk_impl(const int32_t* q_ids, const int64_t* q_off, int64_t nq, int64_t q_nnz, const void* index,
                              const uint32_t* pcard, int64_t np, int32_t n_bits, int64_t nnz, int32_t k, int32_t zero_diag,
                              int64_t query_base, int64_t pool_base, uint32_t* top_inter, uint32_t* top_union,
                              int32_t* top_idx, uint32_t* q_card, const r4d::PeerOut& peers, void* workspace,
                              size_t workspace_bytes, r4d_stream_t stream) {
    using namespace r4d;
    const PostingsLayout L = postings_layout(np, n_bits, nnz);
    R4D_REQUIRE(L.ok, "jaccard_topk_postings: unsupported shape (np=%lld, n_bits=%d [max %d])", (long long)np, n_bits, PJ_MAX_BITS);
    R4D_REQUIRE(nq >= 0 && nq < ((int64_t)1 << 31), "jaccard_topk_postings: nq=%lld", (long long)nq);
    R4D_REQUIRE(k >= 1 && k <= R4D_TOPK_MAX, "jaccard_topk_postings: k=%d out of range [1, %d]", k, R4D_TOPK_MAX);
    R4D_REQUIRE(pool_base >= 0 && pool_base + np < (int64_t)R4D_IDX_NONE, "jaccard_topk_postings: pool_base+np exceeds int32");
    if (nq == 0) return R4D_OK;
    R4D_REQUIRE(q_off && index && (np == 0 || pcard), "jaccard_topk_postings: null pointer");
    R4D_REQUIRE(peers.world > 0 || (top_inter && (top_union || q_card) && top_idx), "jaccard_topk_postings: null output");
    const size_t need = r4d_jaccard_topk_postings_workspace_bytes(nq);
    if (!workspace || workspace_bytes < need) {
        set_error("jaccard_topk_postings: workspace %zu B < required %zu B", workspace_bytes, need);
        return R4D_E_WORKSPACE;
    }
    R4D_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "jaccard_topk_postings: workspace must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    const uint8_t* base = reinterpret_cast<const uint8_t*>(index);
    PJParams prm{};
    prm.q_ids = q_ids;
    prm.q_off = q_off;
    prm.nq = nq;
    prm.off = reinterpret_cast<const uint32_t*>(base + L.off_at);
    prm.post = reinterpret_cast<const uint2*>(base + L.post_at);
    prm.hdr = reinterpret_cast<const PostingsHeader*>(base);
    prm.best = options().postings_best ? reinterpret_cast<const uint2*>(base + L.best_at) : nullptr;
    prm.pcard = pcard;
    prm.np = np;
    prm.n_bits = n_bits;
    prm.win_shift = L.win_shift;
    prm.n_win = L.n_win;
    prm.k = k;
    prm.zero_diag = zero_diag;
    prm.n_fill = (int32_t)(np < k ? np : k);
    prm.query_base = query_base;
    prm.pool_base = pool_base;
    prm.out_inter = top_inter;
    prm.out_union = top_union;
    prm.out_idx = top_idx;
    prm.out_qcard = q_card;
    // workspace: counters [64] | list A [nq] | list B [nq]
    //   counters[0] work of the first stage, [1] entries of list A, [2] work of the heavy kernel,
    //   counters[3] work of the hash-table stage, [4] entries of list B, [5] work of the register stage behind the head
    //   kernel, [6] entries of list A when it is filled a second time, [7] long queries listed for the head kernel, [8] its
    //   work counter over them
    uint32_t* counters = reinterpret_cast<uint32_t*>(workspace);
    uint32_t* list_a = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(workspace) + 256);
    uint32_t* list_b = list_a + nq;
    prm.peers = peers;
    prm.q_out_off = 0;
    prm.nq_total = nq;
    R4D_CUDA(cudaMemsetAsync(workspace, 0, 256, st));
    // room behind the lists for the relay of packed lists to pinned host memory (r4d_jaccard_topk_postings_relay_bytes)
    uint8_t* relay_base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(list_b + nq) + 255) & ~(uintptr_t)255);
    const bool relay_room = q_card != nullptr &&
                            relay_base + r4d_jaccard_topk_postings_relay_bytes(nq, k) - 256 <=
                                reinterpret_cast<uint8_t*>(workspace) + workspace_bytes;
    {
        // table size: expected postings per query = (ids per query) x (postings per id), both known on the host
        const double per_query = (double)(q_nnz > 0 ? q_nnz : 0) / (double)nq * ((double)nnz / (double)n_bits);
        const bool large = options().postings_log_t == PJ_LOG_T_LARGE ||
                           (options().postings_log_t != PJ_LOG_T_SMALL && per_query > 4.0 * pj_cap(PJ_LOG_T_SMALL));
        const bool packed = q_card != nullptr;
        // label-like sets (the small-table regime) take the head kernel first (k <= 16, best lists in use) and the
        // register-resident kernel for what it hands over; option "postings_kernel": 1 = hash-table kernel first
        // (comparison point), 2 = register kernel first (no head kernel)
        const bool reg = !large && options().postings_kernel != 1;
        const bool head = reg && options().postings_kernel != 2 && prm.best != nullptr && k <= PH_K;
        static SmemOptIn opt_in[6];
        const size_t smem_reg = sizeof(PRWarpSmem) * PR_WARPS;
        const size_t smem = head ? sizeof(PHWarpSmem) * PH_WARPS
                          : reg ? smem_reg
                                : (large ? sizeof(PJWarpSmem<PJ_LOG_T_LARGE>) : sizeof(PJWarpSmem<PJ_LOG_T_SMALL>)) * PJ_LIGHT_WARPS;
        void (*kern)(const PJParams) = head ? (packed ? postings_head_kernel<true> : postings_head_kernel<false>)
                                     : reg ? (packed ? postings_reg_kernel<true> : postings_reg_kernel<false>)
                                           : (large ? postings_light_kernel<PJ_LOG_T_LARGE> : postings_light_kernel<PJ_LOG_T_SMALL>);
        if (int rc = ensure_dyn_smem(kern, smem, opt_in[head ? (packed ? 5 : 4) : reg ? (packed ? 3 : 2) : (large ? 1 : 0)])) return rc;
        // every warp should find several grabs of work: small calls take fewer queries per grab
        const int64_t cap = (int64_t)num_sms() * (reg ? 6 : (large ? 2 : 4));   // resident CTAs per SM (shared memory / registers)
        int64_t chunk = nq / (cap * PJ_LIGHT_WARPS * 4);
        bool to_host = false;
        if (top_idx != nullptr && peers.world == 0 && k <= PJ_OBUF_K) {
            // results going straight to pinned HOST memory: the largest chunks make the fewest, widest PCIe writes
            cudaPointerAttributes attr{};
            if (cudaPointerGetAttributes(&attr, top_idx) == cudaSuccess) {
                if (attr.type == cudaMemoryTypeHost) {
                    chunk = PJ_CHUNK;
                    to_host = true;
                }
            } else {
                (void)cudaGetLastError();
            }
        }
        // packed lists bound for the host leave the head kernel in whole 64-query blocks (see PJParams::relay_done)
        const bool relay = head && to_host && packed && relay_room && options().postings_relay != 0 &&
                           (reinterpret_cast<uintptr_t>(top_inter) & 127) == 0 && (reinterpret_cast<uintptr_t>(top_idx) & 127) == 0;
        if (options().postings_chunk > 0) chunk = options().postings_chunk;
        chunk = chunk < 1 ? 1 : (chunk > PJ_CHUNK ? PJ_CHUNK : chunk);
        if (relay) {   // a chunk must not straddle two relay blocks: a power of two
            if (options().postings_chunk <= 0) chunk = 2;
            while (chunk & (chunk - 1)) chunk &= chunk - 1;
        }
        prm.chunk = (int32_t)chunk;
        int64_t grid = (nq + PJ_LIGHT_WARPS * chunk - 1) / (PJ_LIGHT_WARPS * chunk);
        if (grid > cap) grid = cap;
        // first stage: every query; what it cannot serve goes to list A
        prm.in_list = nullptr;
        prm.in_count = nullptr;
        prm.work = counters + 0;
        prm.hand_list = list_a;
        prm.hand_count = counters + 1;
        prof_begin(PROF_JACCARD_POSTINGS, st);
        if (head) {
            // the long queries of the call, listed (in list B, which the later stages reuse) for postings_big_kernel
            prm.big_list = list_b;
            prm.big_count = counters + 7;
            prm.big_work = counters + 8;
            prm.big_ids = PH_BIG_IDS;
            int64_t gb = (nq + 255) / 256;
            if (gb > (int64_t)num_sms() * 4) gb = (int64_t)num_sms() * 4;
            postings_big_scan_kernel<<<(unsigned)gb, 256, 0, st>>>(q_off, nq, PH_BIG_IDS, list_b, counters + 7); note_launch();
        }
        {
            PJParams first = prm;
            if (relay) {
                uint32_t* done = reinterpret_cast<uint32_t*>(relay_base);
                uint8_t* at = relay_base + relay_done_bytes(nq);
                R4D_CUDA(cudaMemsetAsync(done, 0, relay_done_bytes(nq), st));
                first.relay_pair = top_inter;
                first.relay_idx = top_idx;
                first.relay_qcard = q_card;
                first.relay_done = done;
                first.out_inter = reinterpret_cast<uint32_t*>(at);
                first.out_idx = reinterpret_cast<int32_t*>(at + relay_plane_bytes(nq, k));
                first.out_qcard = reinterpret_cast<uint32_t*>(at + 2 * relay_plane_bytes(nq, k));
            }
            if (head) {   // the long queries, a CTA each (all CTAs are resident at once; those without a query exit)
                int64_t gbig = (int64_t)num_sms() * 6;
                if (gbig > nq) gbig = nq;
                if (packed) postings_big_kernel<true><<<(unsigned)gbig, PB_THREADS, 0, st>>>(first);
                else postings_big_kernel<false><<<(unsigned)gbig, PB_THREADS, 0, st>>>(first);
                note_launch();
            }
            kern<<<(unsigned)grid, PJ_LIGHT_WARPS * 32, smem, st>>>(first); note_launch();
        }
        prm.big_list = nullptr;   // (the later stages copy prm)
        // the lists ping-pong between the stages: each stage reads the list its predecessor filled and fills the other
        // (whose earlier contents are consumed by then), with its own counter
        const uint32_t* cur_list = list_a;
        const uint32_t* cur_count = counters + 1;
        if (reg) {
            // register-resident kernels: the first stage without the head kernel (the 8-slot body serves pools with hot
            // (id, window) buckets; one of the two launches exits at once), else ONE launch of the 8-slot body for the
            // queries the head kernel handed over.
            PJParams pr = prm;
            if (head) {
                pr.in_list = list_a;
                pr.in_count = counters + 1;
                pr.work = counters + 5;
                pr.hand_list = list_b;
                pr.hand_count = counters + 4;
                cur_list = list_b;
                cur_count = counters + 4;
            }
            static SmemOptIn opt_in_l[2];
            void (*kern_l)(const PJParams) = packed ? postings_reg_large_kernel<true> : postings_reg_large_kernel<false>;
            if (int rc = ensure_dyn_smem(kern_l, smem_reg, opt_in_l[packed ? 1 : 0])) return rc;
            const int64_t cap_l = (int64_t)num_sms() * 4;
            kern_l<<<(unsigned)(grid > cap_l ? cap_l : grid), PR_WARPS * 32, smem_reg, st>>>(pr); note_launch();
        }
        prof_end(PROF_JACCARD_POSTINGS, st);
        if (reg) {
            // next stage: the register kernel's hand-overs (a row window with more postings than its registers hold, more
            // than 32 ids) are served by the hash-table kernel, whose 512-slot tables take several times as many per
            // pass; what even that cannot serve goes on to the heavy kernel.  An empty list costs one idle launch (~3 us).
            static SmemOptIn opt_in2;
            const size_t smem2 = sizeof(PJWarpSmem<PJ_LOG_T_SMALL>) * PJ_LIGHT_WARPS;
            if (int rc = ensure_dyn_smem(postings_light_kernel<PJ_LOG_T_SMALL>, smem2, opt_in2)) return rc;
            const bool from_a = cur_list == list_a;
            PJParams p2 = prm;
            p2.in_list = cur_list;
            p2.in_count = cur_count;
            p2.work = counters + 3;
            p2.hand_list = from_a ? list_b : list_a;
            p2.hand_count = counters + (from_a ? 4 : 6);
            int64_t grid2 = (int64_t)num_sms() * 2;
            if (grid2 > (nq + PJ_LIGHT_WARPS - 1) / PJ_LIGHT_WARPS) grid2 = (nq + PJ_LIGHT_WARPS - 1) / PJ_LIGHT_WARPS;
            postings_light_kernel<PJ_LOG_T_SMALL><<<(unsigned)grid2, PJ_LIGHT_WARPS * 32, smem2, st>>>(p2); note_launch();
            cur_list = p2.hand_list;
            cur_count = p2.hand_count;
        }
        static SmemOptIn opt_in_h;
        const size_t smem_h = sizeof(PJHeavySmem);
        if (int rc = ensure_dyn_smem(postings_heavy_kernel, smem_h, opt_in_h)) return rc;
        PJParams ph = prm;
        ph.in_list = cur_list;
        ph.in_count = cur_count;
        ph.work = counters + 2;
        ph.hand_list = nullptr;
        ph.hand_count = nullptr;
        int64_t grid_h = (int64_t)num_sms() * 2;
        if (grid_h > nq) grid_h = nq;
        postings_heavy_kernel<<<(unsigned)grid_h, PJ_HEAVY_THREADS, smem_h, st>>>(ph); note_launch();
    }
    R4D_CUDA(cudaGetLastError());
    return R4D_OK;
}

int r4d_jaccard_topk_postings(const int32_t* q_ids, const int64_t* q_off, int64_t nq, int64_t q_nnz, const void* index,
                              const uint32_t* pcard, int64_t np, int32_t n_bits, int64_t nnz, int32_t k, int32_t zero_diag,
                              int64_t query_base, int64_t pool_base, uint32_t* top_inter, uint32_t* top_union,
                              int32_t* top_idx, void* workspace, size_t workspace_bytes, r4d_stream_t stream) {
    r4d::PeerOut none{};
    return postings_topk_impl(q_ids, q_off, nq, q_nnz, index, pcard, np, n_bits, nnz, k, zero_diag, query_base, pool_base, top_inter,
                              top_union, top_idx, nullptr, none, workspace, workspace_bytes, stream);
}

int r4d_jaccard_topk_postings_packed(const int32_t* q_ids, const int64_t* q_off, int64_t nq, int64_t q_nnz, const void* index,
                                     const uint32_t* pcard, int64_t np, int32_t n_bits, int64_t nnz, int32_t k,
                                     int32_t zero_diag, int64_t query_base, int64_t pool_base, uint32_t* top_pair,
                                     int32_t* top_idx, uint32_t* q_card, void* workspace, size_t workspace_bytes,
                                     r4d_stream_t stream) {
    r4d::PeerOut none{};
    R4D_REQUIRE(nq <= 0 || q_card != nullptr, "jaccard_topk_postings_packed: null q_card");
    return postings_topk_impl(q_ids, q_off, nq, q_nnz, index, pcard, np, n_bits, nnz, k, zero_diag, query_base, pool_base, top_pair,
                              nullptr, top_idx, q_card, none, workspace, workspace_bytes, stream);
}

int r4d_jaccard_topk_postings_scatter(const int32_t* q_ids, const int64_t* q_off, int64_t nq, int64_t q_nnz, const void* index,
                                      const uint32_t* pcard, int64_t np, int32_t n_bits, int64_t nnz, int32_t k,
                                      int32_t zero_diag, int64_t query_base, int64_t pool_base, void* const* peer_base,
                                      int32_t world, int32_t rank, void* workspace, size_t workspace_bytes,
                                      r4d_stream_t stream) {
    using namespace r4d;
    R4D_REQUIRE(peer_base && world >= 1 && world <= R4D_MAX_PEERS && rank >= 0 && rank < world,
                "fused exchange: world=%d rank=%d (max %d peers)", world, rank, R4D_MAX_PEERS);
    PeerOut po{};
    po.world = world;
    po.rank = rank;
    for (int r = 0; r < world; ++r) {
        R4D_REQUIRE(peer_base[r] != nullptr, "fused exchange: null peer pointer %d", r);
        po.base[r] = peer_base[r];
    }
    return postings_topk_impl(q_ids, q_off, nq, q_nnz, index, pcard, np, n_bits, nnz, k, zero_diag, query_base, pool_base, nullptr,
                              nullptr, nullptr, nullptr, po, workspace, workspace_bytes, stream);
}

}  // extern "C"
