// jaccard_sparse.cu — fused Jaccard top-K with a SPARSE QUERY SIDE (pool side unchanged: bitset tiles streamed by TMA,
// AND + POPC on 8-word spans, exact rational ranking, warp-level top-K lists).
//
// Why: node-id sets are tiny next to the vocabulary (2.2 of 20 000 bits), so a query row has ~2 non-zero 8-word spans
// out of 80.  The dense kernel (jaccard.cu) still walks all 80 spans of every (query slab, pool slab) to find that out
// and re-streams the query tile for every pool tile.  Here the query tile is turned ONCE per launch into a list of its
// non-zero spans, bucketed by 32-word chunk; per pool chunk the entries are dealt round-robin to the 16 warps and each
// is tested against the 128 pool rows of the tile.  Non-zero intersections are rare, so they are accumulated with smem
// atomics in a count tile and remembered in the "touched" list of the warp that owns the query row; the per-tile
// epilogue visits only those cells.
// Zero-score candidates matter only as the lowest-index filler of a short list: every list starts with the first k pool
// rows of its stripe as (score 0) placeholders, which real candidates for the same row replace.
//
// Query tiles with more than SQ_E_MAX non-zero spans (dense data, e.g. history sets) are flagged by the span-list
// kernel and handled by the dense kernel in the same launch sequence; both write the same per-stripe partial lists.
// Results are bit-identical to the dense kernel (tests/test_gpu_jaccard.py).
#include "jaccard_common.cuh"

namespace r4d {

// ---------------------------------------------------------------------------- span lists of a query tile
// grid = n_qtiles CTAs.  Entries are bucketed by 32-word chunk; inside a chunk their order is arbitrary.
__global__ void __launch_bounds__(256)
qspans_kernel(const uint32_t* __restrict__ qbits, int64_t nq, int32_t pitch_words, int32_t n_chunks, SparseQ sq) {
    __shared__ uint32_t cnt[SQ_MAX_CHUNKS];
    __shared__ uint32_t cur[SQ_MAX_CHUNKS];
    __shared__ uint32_t lane_tot[32];
    __shared__ uint32_t total_s;
    const int t = blockIdx.x, tid = threadIdx.x;
    const int nb = n_chunks, n_spans = n_chunks * 4;
    for (int i = tid; i < nb; i += 256) cnt[i] = 0u;
    __syncthreads();

    auto span_words = [&](int r, int s, uint4& a, uint4& b) -> bool {
        a = make_uint4(0, 0, 0, 0);
        b = make_uint4(0, 0, 0, 0);
        const int64_t gq = (int64_t)t * SQ_TQ + r;
        if (gq >= nq || s * 8 >= pitch_words) return false;
        const uint4* row = reinterpret_cast<const uint4*>(qbits + gq * pitch_words + s * 8);
        a = row[0];
        if (s * 8 + 4 < pitch_words) b = row[1];
        return ((a.x | a.y | a.z | a.w) | (b.x | b.y | b.z | b.w)) != 0u;
    };

    for (int u = tid; u < SQ_TQ * n_spans; u += 256) {
        const int r = u / n_spans, s = u - r * n_spans;
        uint4 a, b;
        if (span_words(r, s, a, b)) atomicAdd(&cnt[s >> 2], 1u);
    }
    __syncthreads();
    if (tid < 32) {  // exclusive scan of the bucket counts by one warp
        const int per = (nb + 31) / 32;
        uint32_t sum = 0;
        for (int i = tid * per; i < min(nb, (tid + 1) * per); ++i) sum += cnt[i];
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (tid >= o) incl += v;
        }
        uint32_t run = incl - sum;
        for (int i = tid * per; i < min(nb, (tid + 1) * per); ++i) {
            cur[i] = run;
            run += cnt[i];
        }
        if (tid == 31) total_s = incl;
        (void)lane_tot;
    }
    __syncthreads();
    const uint32_t total = total_s;
    if (total > (uint32_t)SQ_E_MAX) {  // too dense for the sparse path: the dense kernel takes this tile
        if (tid == 0) sq.tile_dense[t] = 1u;
        return;
    }
    uint16_t* off = sq.off + (size_t)t * SQ_OFF_LD;
    for (int i = tid; i < nb; i += 256) off[i] = (uint16_t)cur[i];
    if (tid == 0) {
        off[nb] = (uint16_t)total;
        sq.tile_dense[t] = 0u;
    }
    __syncthreads();
    uint32_t* hdr = sq.hdr + (size_t)t * SQ_E_MAX;
    uint4* words = reinterpret_cast<uint4*>(sq.words + (size_t)t * SQ_E_MAX * 8);
    for (int u = tid; u < SQ_TQ * n_spans; u += 256) {
        const int r = u / n_spans, s = u - r * n_spans;
        uint4 a, b;
        if (span_words(r, s, a, b)) {
            const uint32_t pos = atomicAdd(&cur[s >> 2], 1u);
            hdr[pos] = (uint32_t)r | ((uint32_t)(s & 3) << 8);
            words[pos * 2] = a;
            words[pos * 2 + 1] = b;
        }
    }
}

// ---------------------------------------------------------------------------- main kernel
struct SparseParams {
    const uint32_t* qcard;
    const uint32_t* pcard;
    int64_t nq, np;
    int32_t n_chunks, k, zero_diag, n_stages;
    int64_t query_base, pool_base;
    int32_t n_qtiles, n_ptiles, n_stripes, ptiles_per_stripe;
    uint32_t* part_inter;
    uint32_t* part_union;
    int32_t* part_idx;
    SparseQ sq;
};

__device__ __forceinline__ void lds128s(uint4& v, uint32_t addr) {
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
}

__global__ void __launch_bounds__(SQ_THREADS + 32, 1)
jaccard_sparse_kernel(const __grid_constant__ CUtensorMap tm_p, const SparseParams prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int NST = prm.n_stages;
    uint8_t* stages = smem;                                                      // [NST][128 rows x 128 B]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)NST * SQ_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + 8;
    uint32_t* acc = reinterpret_cast<uint32_t*>(empty_bar + 8);                  // [128][128] intersection counts
    uint4* ent_words = reinterpret_cast<uint4*>(acc + SQ_TQ * SQ_TP);            // [E_MAX][2]
    uint32_t* ent_hdr = reinterpret_cast<uint32_t*>(ent_words + SQ_E_MAX * 2);   // [E_MAX]
    uint16_t* off_s = reinterpret_cast<uint16_t*>(ent_hdr + SQ_E_MAX);           // [SQ_OFF_LD]
    uint16_t* touched = off_s + SQ_OFF_LD;                                       // [16 owners][SQ_TOUCH_CAP]
    uint32_t* t_cnt = reinterpret_cast<uint32_t*>(touched + SQ_WARPS * SQ_TOUCH_CAP);    // [16] cells per owner
    uint32_t* l_inter = t_cnt + SQ_WARPS;                                        // [128][k] x 3
    const int K = prm.k;
    uint32_t* l_union = l_inter + SQ_TQ * K;
    int32_t* l_idx = reinterpret_cast<int32_t*>(l_union + SQ_TQ * K);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_p);
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], SQ_WARPS);
        }
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < SQ_TQ * SQ_TP; i += SQ_THREADS + 32) acc[i] = 0u;  // cells are reset after every use
    if (threadIdx.x < SQ_WARPS) t_cnt[threadIdx.x] = 0u;
    __syncthreads();

    const int n_items = prm.n_qtiles * prm.n_stripes;
    auto item_mine = [&](int item) { return prm.sq.tile_dense[item % prm.n_qtiles] == 0u; };

    // ---- dedicated TMA producer: warp 16 (this kernel needs < 96 registers per thread, so a 17th warp is affordable;
    // an inline producer lane would have to wait for the slowest warp of the previous chunk before its own chunk).
    if (warp == SQ_WARPS) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                if (!item_mine(item)) continue;
                const int stripe = item / prm.n_qtiles;
                const int pt_beg = stripe * prm.ptiles_per_stripe;
                const int pt_end = min(pt_beg + prm.ptiles_per_stripe, prm.n_ptiles);
                for (int pt = pt_beg; pt < pt_end; ++pt)
                    for (int c = 0; c < prm.n_chunks; ++c) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&full_bar[stage], SQ_STAGE_BYTES);
                        tma_load_2d(stages + (size_t)stage * SQ_STAGE_BYTES, &tm_p, &full_bar[stage], c * SQ_CHUNK_WORDS,
                                    pt * SQ_TP);
                        if (++stage == NST) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
            }
        }
        return;
    }

    const uint32_t stages_u32 = smem_u32(stages);
    const uint32_t sw = (uint32_t)(lane & 7);  // (pool row & 7) of every row this lane tests: rows lane + 32 j
    int stage = 0;
    uint32_t phase = 0;
    uint16_t* my_touched = touched + warp * SQ_TOUCH_CAP;  // cells of the rows this warp owns (r & 15 == warp)

    // insert a non-zero candidate into the list of query row r (all 32 lanes call with warp-uniform arguments)
    auto take = [&](int r, int64_t gq, int64_t gp, uint32_t inter) {
        if (gq >= prm.nq || gp >= prm.np || inter == 0u) return;
        if (prm.zero_diag && prm.query_base + gq == prm.pool_base + gp) return;  // score forced to 0: stays a filler
        const uint32_t uni = prm.qcard[gq] + prm.pcard[gp] - inter;
        const JEntry c{inter, uni, (int32_t)(prm.pool_base + gp)};
        JEntry mine = lane < K ? JEntry{l_inter[r * K + lane], l_union[r * K + lane], l_idx[r * K + lane]} : JEntry::worst();
        // a zero-score placeholder of the same pool row is replaced, not duplicated
        const uint32_t same = __ballot_sync(0xffffffffu, lane < K && mine.idx == c.idx);
        if (same) {
            const int pos = __ffs(same) - 1;
            const JEntry dn = mine.shfl_down1();
            if (lane >= pos && lane < K) mine = (lane == K - 1) ? JEntry::worst() : dn;
        }
        WarpTopK<JEntry> tk;
        tk.k = K;
        tk.mine = mine;
        tk.refresh_kth();
        tk.insert(c);
        if (lane < K) {
            l_inter[r * K + lane] = tk.mine.inter;
            l_union[r * K + lane] = tk.mine.uni;
            l_idx[r * K + lane] = tk.mine.idx;
        }
        __syncwarp();
    };

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        if (!item_mine(item)) continue;  // block-uniform
        const int stripe = item / prm.n_qtiles;
        const int qtile = item - stripe * prm.n_qtiles;
        const int pt_beg = stripe * prm.ptiles_per_stripe;
        const int pt_end = min(pt_beg + prm.ptiles_per_stripe, prm.n_ptiles);
        const int nb = prm.n_chunks;

        named_bar_sync(1, SQ_THREADS);  // previous item's lists / entries are no longer in use (consumer warps only)
        {
            const uint32_t* g_off = reinterpret_cast<const uint32_t*>(prm.sq.off + (size_t)qtile * SQ_OFF_LD);
            uint32_t* s_off = reinterpret_cast<uint32_t*>(off_s);
            for (int i = threadIdx.x; i < SQ_OFF_LD / 2; i += SQ_THREADS) s_off[i] = g_off[i];
        }
        named_bar_sync(1, SQ_THREADS);
        const int total = off_s[nb];
        {
            const uint32_t* g_hdr = prm.sq.hdr + (size_t)qtile * SQ_E_MAX;
            const uint4* g_words = reinterpret_cast<const uint4*>(prm.sq.words + (size_t)qtile * SQ_E_MAX * 8);
            for (int i = threadIdx.x; i < total; i += SQ_THREADS) ent_hdr[i] = g_hdr[i];
            for (int i = threadIdx.x; i < total * 2; i += SQ_THREADS) ent_words[i] = g_words[i];
        }
        // lists of the 8 rows this warp owns (r & 15 == warp) start with the stripe's first k pool rows at score 0
        for (int i = 0; i < 8; ++i) {
            const int r = warp + 16 * i;
            const int64_t gq = (int64_t)qtile * SQ_TQ + r;
            if (lane < K) {
                const int64_t gp = (int64_t)pt_beg * SQ_TP + lane;
                const bool ok = gq < prm.nq && gp < prm.np && gp < (int64_t)pt_end * SQ_TP;
                const uint32_t u = ok ? max(prm.qcard[gq] + prm.pcard[gp], 1u) : 1u;
                l_inter[r * K + lane] = 0u;
                l_union[r * K + lane] = u;
                l_idx[r * K + lane] = ok ? (int32_t)(prm.pool_base + gp) : R4D_IDX_NONE;
            }
        }
        named_bar_sync(1, SQ_THREADS);

        for (int pt = pt_beg; pt < pt_end; ++pt) {
            for (int c = 0; c < prm.n_chunks; ++c) {
                mbar_wait(&full_bar[stage], phase);
                const uint32_t sbase = stages_u32 + (uint32_t)stage * SQ_STAGE_BYTES;
                // the chunk's entries are dealt round-robin to the 16 warps: balanced whatever rows they belong to
                const int e_end = off_s[c + 1];
                for (int e = off_s[c] + warp; e < e_end; e += SQ_WARPS) {
                    const uint32_t h = ent_hdr[e];
                    const int r = (int)(h & 0xffu);
                    const uint32_t sub2 = ((h >> 8) & 3u) * 2u;  // first 16-byte unit of the span inside the chunk
                    const uint4 qa = ent_words[e * 2], qb = ent_words[e * 2 + 1];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int p = lane + 32 * j;
                        const uint32_t row_addr = sbase + (uint32_t)p * 128u;
                        uint4 pa, pb;
                        lds128s(pa, row_addr + ((sub2 ^ sw) << 4));
                        lds128s(pb, row_addr + (((sub2 + 1u) ^ sw) << 4));
                        if (((pa.x | pa.y | pa.z | pa.w) | (pb.x | pb.y | pb.z | pb.w)) != 0u) {
                            const uint32_t v = __popc(qa.x & pa.x) + __popc(qa.y & pa.y) + __popc(qa.z & pa.z) +
                                               __popc(qa.w & pa.w) + __popc(qb.x & pb.x) + __popc(qb.y & pb.y) +
                                               __popc(qb.z & pb.z) + __popc(qb.w & pb.w);
                            // another warp may hold a different span of the same query row: atomic (and rare)
                            if (v && atomicAdd(&acc[r * SQ_TP + p], v) == 0u) {
                                const uint32_t pos = atomicAdd(&t_cnt[r & 15], 1u);  // first hit: remember the cell
                                if (pos < (uint32_t)SQ_TOUCH_CAP) touched[(r & 15) * SQ_TOUCH_CAP + pos] = (uint16_t)((r << 7) | p);
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[stage]);
                if (++stage == NST) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            // ---- per-tile epilogue: every warp serves the rows it owns (r & 15 == warp), only non-zero cells
            named_bar_sync(1, SQ_THREADS);  // all counts of this pool tile are final
            const uint32_t tcount = t_cnt[warp];
            if (tcount > (uint32_t)SQ_TOUCH_CAP) {
                // list overflowed (many matches): scan the 8 x 128 cells this warp owns
                for (int i = 0; i < 8; ++i) {
                    const int r = warp + 16 * i;
                    const int64_t gq = (int64_t)qtile * SQ_TQ + r;
#pragma unroll 1
                    for (int j = 0; j < 4; ++j) {
                        const int p = lane + 32 * j;
                        const uint32_t v = acc[r * SQ_TP + p];
                        acc[r * SQ_TP + p] = 0u;
                        uint32_t m = __ballot_sync(0xffffffffu, v != 0u);
                        while (m) {
                            const int src = __ffs(m) - 1;
                            m &= m - 1;
                            take(r, gq, (int64_t)pt * SQ_TP + src + 32 * j, __shfl_sync(0xffffffffu, v, src));
                        }
                    }
                }
            } else {
                for (uint32_t t = 0; t < tcount; ++t) {
                    const uint32_t cell = my_touched[t];
                    const int r = (int)(cell >> 7), p = (int)(cell & 127u);
                    const uint32_t v = acc[r * SQ_TP + p];
                    __syncwarp();
                    if (lane == 0) acc[r * SQ_TP + p] = 0u;
                    take(r, (int64_t)qtile * SQ_TQ + r, (int64_t)pt * SQ_TP + p, v);
                }
            }
            __syncwarp();
            if (lane == 0) t_cnt[warp] = 0u;
            named_bar_sync(1, SQ_THREADS);  // cells and counters are clean before the next pool tile accumulates
        }

        // flush the lists of the rows this warp owns
        for (int i = 0; i < 8; ++i) {
            const int r = warp + 16 * i;
            const int64_t gq = (int64_t)qtile * SQ_TQ + r;
            if (gq < prm.nq && lane < K) {
                const int64_t o = ((int64_t)stripe * prm.nq + gq) * K + lane;
                prm.part_inter[o] = l_inter[r * K + lane];
                prm.part_union[o] = l_union[r * K + lane];
                prm.part_idx[o] = l_idx[r * K + lane];
            }
        }
    }
}

// ---------------------------------------------------------------------------- host side
size_t sparseq_workspace_bytes(int64_t nq) {
    const size_t n_qtiles = (size_t)((nq + SQ_TQ - 1) / SQ_TQ);
    return 256 + n_qtiles * (sizeof(uint32_t) + SQ_OFF_LD * sizeof(uint16_t) + SQ_E_MAX * 4 + (size_t)SQ_E_MAX * 32) + 256;
}

bool sparseq_supported(int32_t words, int32_t k) {
    const int n_chunks = (words + SQ_CHUNK_WORDS - 1) / SQ_CHUNK_WORDS;
    return options().jaccard_sparse_q != 0 && options().jaccard_skip_zero != 0 && n_chunks <= SQ_MAX_CHUNKS && k <= R4D_TOPK_MAX;
}

SparseQ sparseq_carve(void* base, int64_t nq) {
    const size_t n_qtiles = (size_t)((nq + SQ_TQ - 1) / SQ_TQ);
    uint8_t* p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(base) + 255) & ~uintptr_t(255));
    SparseQ sq;
    sq.words = reinterpret_cast<uint32_t*>(p);
    p += n_qtiles * (size_t)SQ_E_MAX * 32;
    sq.hdr = reinterpret_cast<uint32_t*>(p);
    p += n_qtiles * (size_t)SQ_E_MAX * 4;
    sq.off = reinterpret_cast<uint16_t*>(p);
    p += n_qtiles * (size_t)SQ_OFF_LD * 2;
    sq.tile_dense = reinterpret_cast<uint32_t*>(p);
    return sq;
}

int sparseq_build(const uint32_t* qbits, int64_t nq, int32_t words, int32_t pitch_words, const SparseQ& sq, cudaStream_t st) {
    const int n_qtiles = (int)((nq + SQ_TQ - 1) / SQ_TQ);
    const int n_chunks = (words + SQ_CHUNK_WORDS - 1) / SQ_CHUNK_WORDS;
    qspans_kernel<<<n_qtiles, 256, 0, st>>>(qbits, nq, pitch_words, n_chunks, sq);
    R4D_CUDA(cudaGetLastError());
    return R4D_OK;
}

int sparseq_topk_launch(const uint32_t* pbits, const uint32_t* qcard, const uint32_t* pcard, int64_t nq, int64_t np,
                        int32_t words, int32_t pitch_words, int32_t k, int32_t zero_diag, int64_t query_base,
                        int64_t pool_base, int32_t n_qtiles, int32_t n_ptiles, int32_t n_stripes, int32_t ptiles_per_stripe,
                        uint32_t* part_inter, uint32_t* part_union, int32_t* part_idx, const SparseQ& sq, cudaStream_t st) {
    CUtensorMap tm_p;
    int rc = make_tmap_2d(&tm_p, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, pbits, (uint64_t)pitch_words, (uint64_t)np,
                          (uint64_t)pitch_words * 4, SQ_CHUNK_WORDS, SQ_TP, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    SparseParams prm{};
    prm.qcard = qcard;
    prm.pcard = pcard;
    prm.nq = nq;
    prm.np = np;
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
    prm.sq = sq;
    const size_t fixed = 1024 + 16 * sizeof(uint64_t) + (size_t)SQ_TQ * SQ_TP * 4 + (size_t)SQ_E_MAX * 32 + (size_t)SQ_E_MAX * 4 +
                         (size_t)SQ_OFF_LD * 2 + (size_t)SQ_WARPS * SQ_TOUCH_CAP * 2 + SQ_WARPS * 4 + 3 * (size_t)SQ_TQ * (size_t)k * 4;
    int n_stages = (int)((227 * 1024 - fixed) / SQ_STAGE_BYTES);
    if (n_stages > 8) n_stages = 8;
    if (n_stages < 2) {
        set_error("jaccard sparse path: not enough shared memory (k=%d)", k);
        return R4D_E_ARG;
    }
    prm.n_stages = n_stages;
    const size_t smem = fixed + (size_t)n_stages * SQ_STAGE_BYTES;
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        R4D_CUDA(cudaFuncSetAttribute(jaccard_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    const int64_t n_items = (int64_t)n_qtiles * n_stripes;
    int grid = num_sms();
    if (n_items < grid) grid = (int)n_items;
    jaccard_sparse_kernel<<<grid, SQ_THREADS + 32, smem, st>>>(tm_p, prm);
    R4D_CUDA(cudaGetLastError());
    return R4D_OK;
}

}  // namespace r4d
