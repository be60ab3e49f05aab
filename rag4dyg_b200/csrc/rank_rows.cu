// rank_rows.cu — ranking / selection on explicit score matrices, canonical order (score desc, index asc).
//
//  r4d_rank_rows_f64/f32 : order[q] = np.argsort(-scores[q], kind='stable')     (save_index_score,
//                          retrieval_data_annotation.py:89; train/train_retriever.py:358)
//  r4d_topk_rows_f64     : np.argsort(-row, kind='stable')[:k] + the scores      (save_score_file_train, :97-103)
//  r4d_triplet_mine_f64  : positives count + hard/fill negatives per row         (save_train_annotation, :54-71)
//
// The ranking is a per-row LSD radix sort (8-bit digits, stable by construction, so equal scores keep
// ascending pool index without putting the index in the key).  One CTA per row, keys + permutation ping-pong in
// a caller-provided global workspace that stays L2 resident (rows are <= a few 10^4 long).
#include "r4d_common.cuh"

namespace r4d {

constexpr int RK_THREADS = 256;
constexpr int RK_WARPS = RK_THREADS / 32;
constexpr int RK_ITEMS = 8;                             // items per lane per tile
constexpr int RK_TILE = RK_THREADS * RK_ITEMS;          // 2048 items per tile

// order-preserving key: ascending key order == descending score order; all zeros equal; NaN last.
__device__ __forceinline__ uint64_t desc_key(double x) {
    if (x != x) return ~0ull;
    uint64_t b = (x == 0.0) ? 0ull : (uint64_t)__double_as_longlong(x);
    b ^= 0x8000000000000000ull;  // y = -x
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ uint32_t desc_key(float x) {
    if (x != x) return ~0u;
    uint32_t b = (x == 0.0f) ? 0u : (uint32_t)__float_as_int(x);
    b ^= 0x80000000u;
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

template <class T> struct KeyOf;
template <> struct KeyOf<double> { typedef uint64_t type; };
template <> struct KeyOf<float> { typedef uint32_t type; };

template <class T>
__global__ void __launch_bounds__(RK_THREADS)
rank_rows_kernel(const T* __restrict__ scores, int64_t nq, int64_t n, int64_t ld, int32_t* __restrict__ order,
                 uint8_t* workspace, size_t ws_per_cta) {
    typedef typename KeyOf<T>::type K;
    constexpr int NPASS = sizeof(K);
    __shared__ uint32_t hist_all[NPASS][256];
    __shared__ uint32_t whist[RK_WARPS][256];
    __shared__ uint32_t binbase[256];
    __shared__ uint32_t warp_tot[RK_WARPS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* ws = workspace + (size_t)blockIdx.x * ws_per_cta;
    K* keyA = reinterpret_cast<K*>(ws);
    K* keyB = keyA + n;
    int32_t* idxA = reinterpret_cast<int32_t*>(keyB + n);
    int32_t* idxB = idxA + n;

    for (int64_t row = blockIdx.x; row < nq; row += gridDim.x) {
        const T* src = scores + row * ld;
        // ---- sweep 0: keys + all digit histograms
        for (int i = tid; i < NPASS * 256; i += RK_THREADS) (&hist_all[0][0])[i] = 0u;
        __syncthreads();
        for (int64_t i = tid; i < n; i += RK_THREADS) {
            const K key = desc_key(src[i]);
            keyA[i] = key;
            idxA[i] = (int32_t)i;
#pragma unroll
            for (int p = 0; p < NPASS; ++p) atomicAdd(&hist_all[p][(key >> (8 * p)) & 0xff], 1u);
        }
        __syncthreads();

        K* kin = keyA;
        K* kout = keyB;
        int32_t* iin = idxA;
        int32_t* iout = idxB;
        for (int p = 0; p < NPASS; ++p) {
            // skip a pass whose digit is the same for every key (warp-uniform decision via smem)
            const uint32_t mine = hist_all[p][tid];
            const int all_same = __syncthreads_or(mine == (uint32_t)n);
            if (all_same) continue;
            // exclusive scan of the 256 bin counts -> binbase
            uint32_t incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) warp_tot[warp] = incl;
            __syncthreads();
            uint32_t woff = 0;
            for (int w = 0; w < warp; ++w) woff += warp_tot[w];
            binbase[tid] = woff + incl - mine;
            for (int w = 0; w < RK_WARPS; ++w) whist[w][tid] = 0u;
            __syncthreads();

            for (int64_t tile = 0; tile < n; tile += RK_TILE) {
                // each warp owns a contiguous run of 32*RK_ITEMS items; iteration `it` covers 32 consecutive ones
                const int64_t wbase = tile + (int64_t)warp * 32 * RK_ITEMS;
                K key[RK_ITEMS];
                int32_t ix[RK_ITEMS];
#pragma unroll
                for (int it = 0; it < RK_ITEMS; ++it) {
                    const int64_t i = wbase + it * 32 + lane;
                    if (i < n) {
                        key[it] = kin[i];
                        ix[it] = iin[i];
                        atomicAdd(&whist[warp][(key[it] >> (8 * p)) & 0xff], 1u);
                    }
                }
                __syncthreads();
                {  // bin `tid`: turn per-warp counts into per-warp start offsets, advance the running base
                    uint32_t run = binbase[tid];
#pragma unroll
                    for (int w = 0; w < RK_WARPS; ++w) {
                        const uint32_t c = whist[w][tid];
                        whist[w][tid] = run;
                        run += c;
                    }
                    binbase[tid] = run;
                }
                __syncthreads();
#pragma unroll
                for (int it = 0; it < RK_ITEMS; ++it) {
                    const int64_t i = wbase + it * 32 + lane;
                    const bool valid = i < n;
                    const uint32_t d = valid ? (uint32_t)((key[it] >> (8 * p)) & 0xff) : 0x100u + lane;
                    const uint32_t peers = __match_any_sync(0xffffffffu, d);
                    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
                    uint32_t base = 0;
                    if (valid) base = whist[warp][d];
                    __syncwarp();
                    if (valid) {
                        const uint32_t pos = base + rank;
                        kout[pos] = key[it];
                        iout[pos] = ix[it];
                        if (rank == 0) whist[warp][d] = base + __popc(peers);
                    }
                    __syncwarp();
                }
                __syncthreads();
                for (int w = 0; w < RK_WARPS; ++w) whist[w][tid] = 0u;
                __syncthreads();
            }
            K* tk = kin; kin = kout; kout = tk;
            int32_t* ti = iin; iin = iout; iout = ti;
        }
        __syncthreads();
        int32_t* dst = order + row * n;
        for (int64_t i = tid; i < n; i += RK_THREADS) dst[i] = iin[i];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------- top-k of explicit rows
__global__ void __launch_bounds__(256)
topk_rows_f64_kernel(const double* __restrict__ scores, int64_t nq, int64_t n, int64_t ld, int32_t k,
                     double* __restrict__ top_score, int32_t* __restrict__ top_idx) {
    const int lane = threadIdx.x & 31;
    const int64_t wpg = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < nq; q += wpg) {
        const double* row = scores + q * ld;
        WarpTopK<FEntry<double>> tk;
        tk.init(k);
        for (int64_t j0 = 0; j0 < n; j0 += 32) {
            const int64_t j = j0 + lane;
            FEntry<double> c = FEntry<double>::worst();
            if (j < n) c = FEntry<double>{row[j], (int32_t)j};
            uint32_t m = __ballot_sync(0xffffffffu, j < n && FEntry<double>::better(c, tk.kth));
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                tk.insert(c.shfl(src));
            }
        }
        if (lane < k) {
            top_score[q * k + lane] = tk.mine.s;
            top_idx[q * k + lane] = tk.mine.idx;
        }
    }
}

// ---------------------------------------------------------------------------- triplet mining
__global__ void __launch_bounds__(256)
triplet_mine_f64_kernel(const double* __restrict__ out, const double* __restrict__ in, int64_t n, int64_t ld,
                        double thr, int32_t neg_num, int32_t* __restrict__ n_pos, int32_t* __restrict__ neg,
                        int32_t* __restrict__ n_neg) {
    const int lane = threadIdx.x & 31;
    const int64_t wpg = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += wpg) {
        const double* orow = out + i * ld;
        const double* irow = in + i * ld;
        WarpTopK<FEntry<double>> hard, fill;
        hard.init(neg_num);
        fill.init(neg_num);
        int32_t pos_cnt = 0;
        for (int64_t j0 = 0; j0 < n; j0 += 32) {
            const int64_t j = j0 + lane;
            double o = 0.0, s = 0.0;
            if (j < n) {
                o = orow[j];
                s = irow[j];
            }
            const bool is_pos = (j < n) && (o > thr);                 // :54  strict >
            const bool is_hard = (j < n) && !is_pos && (o > 0.0);     // :60
            const bool is_fill = (j < n) && !is_pos && (o == 0.0);    // :67
            pos_cnt += __popc(__ballot_sync(0xffffffffu, is_pos));
            FEntry<double> c{s, (int32_t)j};
            uint32_t m = __ballot_sync(0xffffffffu, is_hard && FEntry<double>::better(c, hard.kth));
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                hard.insert(c.shfl(src));
            }
            m = __ballot_sync(0xffffffffu, is_fill && FEntry<double>::better(c, fill.kth));
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                fill.insert(c.shfl(src));
            }
        }
        const int n_hard = __popc(__ballot_sync(0xffffffffu, lane < neg_num && hard.mine.idx != R4D_IDX_NONE));
        const int n_fill = __popc(__ballot_sync(0xffffffffu, lane < neg_num && fill.mine.idx != R4D_IDX_NONE));
        const int total = min(neg_num, n_hard + n_fill);
        // slot t: hard[t] for t < n_hard, else fill[t - n_hard]
        const int fsrc = lane - n_hard;
        const int32_t from_fill = __shfl_sync(0xffffffffu, fill.mine.idx, fsrc < 0 ? 0 : (fsrc > 31 ? 31 : fsrc));
        if (lane < neg_num) neg[i * neg_num + lane] = lane < n_hard ? hard.mine.idx : (lane < total ? from_fill : -1);
        if (lane == 0) {
            n_pos[i] = pos_cnt;
            n_neg[i] = total;
        }
    }
}

// ---------------------------------------------------------------------------- counter-based negative sampling
// One thread per positive pair t: neg_choice[t] = neg[row][hash(seed, row, rank-of-t-within-row) % n_neg[row]].
// Stateless (splitmix64 of a counter), so the triplet list no longer needs the host's sequential np.random.choice.
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256)
triplet_sample_kernel(const int64_t* __restrict__ pos_row, const int64_t* __restrict__ row_start, int64_t n_pairs,
                      const int32_t* __restrict__ neg, const int32_t* __restrict__ n_neg, int32_t neg_num, uint64_t seed,
                      int32_t* __restrict__ choice) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_pairs; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = pos_row[t];
        const int64_t rank = t - row_start[row];
        const int32_t cnt = n_neg[row];
        int32_t c = -1;
        if (cnt > 0) {
            const uint64_t h = splitmix64(splitmix64(seed ^ (uint64_t)row * 0xd6e8feb86659fd93ull) + (uint64_t)rank);
            c = neg[row * neg_num + (int32_t)(h % (uint64_t)cnt)];
        }
        choice[t] = c;
    }
}

template <class T>
static int rank_rows_impl(const T* scores, int64_t nq, int64_t n, int64_t ld, int32_t* order, void* workspace,
                          size_t workspace_bytes, cudaStream_t st) {
    R4D_REQUIRE(nq >= 0 && n >= 0 && n < (int64_t)R4D_IDX_NONE && ld >= n, "rank_rows: nq=%lld n=%lld ld=%lld",
                (long long)nq, (long long)n, (long long)ld);
    if (nq == 0 || n == 0) return R4D_OK;
    R4D_REQUIRE(scores && order && workspace, "rank_rows: null pointer");
    const size_t per = ((size_t)n * (2 * sizeof(T) + 8) + 255) / 256 * 256;
    int64_t grid = (int64_t)num_sms() * 2;
    if (grid > nq) grid = nq;
    if (workspace_bytes < per * (size_t)grid) {
        set_error("rank_rows: workspace %zu B < required %zu B", workspace_bytes, per * (size_t)grid);
        return R4D_E_WORKSPACE;
    }
    rank_rows_kernel<T><<<(unsigned)grid, RK_THREADS, 0, st>>>(scores, nq, n, ld, order,
                                                              reinterpret_cast<uint8_t*>(workspace), per); note_launch();
    R4D_CUDA(cudaGetLastError());
    return R4D_OK;
}

}  // namespace r4d

extern "C" {

size_t r4d_rank_rows_workspace_bytes(int64_t nq, int64_t n, int32_t elem_bytes) {
    if (nq <= 0 || n <= 0) return 256;
    const size_t per = ((size_t)n * (2 * (size_t)elem_bytes + 8) + 255) / 256 * 256;
    int64_t grid = (int64_t)r4d::num_sms() * 2;
    if (grid > nq) grid = nq;
    return per * (size_t)grid;
}

int r4d_rank_rows_f64(const double* scores, int64_t nq, int64_t n, int64_t ld, int32_t* order, void* workspace,
                      size_t workspace_bytes, r4d_stream_t stream) {
    return r4d::rank_rows_impl<double>(scores, nq, n, ld, order, workspace, workspace_bytes, r4d::as_stream(stream));
}

int r4d_rank_rows_f32(const float* scores, int64_t nq, int64_t n, int64_t ld, int32_t* order, void* workspace,
                      size_t workspace_bytes, r4d_stream_t stream) {
    return r4d::rank_rows_impl<float>(scores, nq, n, ld, order, workspace, workspace_bytes, r4d::as_stream(stream));
}

int r4d_topk_rows_f64(const double* scores, int64_t nq, int64_t n, int64_t ld, int32_t k, double* top_score,
                      int32_t* top_idx, r4d_stream_t stream) {
    using namespace r4d;
    R4D_REQUIRE(nq >= 0 && n >= 0 && ld >= n && k >= 1 && k <= R4D_TOPK_MAX, "topk_rows: nq=%lld n=%lld k=%d",
                (long long)nq, (long long)n, k);
    if (nq == 0) return R4D_OK;
    R4D_REQUIRE((scores || n == 0) && top_score && top_idx, "topk_rows: null pointer");
    int64_t blocks = (nq + 7) / 8;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    topk_rows_f64_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(scores, nq, n, ld, k, top_score, top_idx); note_launch();
    R4D_CUDA(cudaGetLastError());
    return R4D_OK;
}

int r4d_triplet_sample(const int64_t* pos_row, const int64_t* row_start, int64_t n_pairs, const int32_t* neg,
                       const int32_t* n_neg, int32_t neg_num, uint64_t seed, int32_t* choice, r4d_stream_t stream) {
    using namespace r4d;
    R4D_REQUIRE(n_pairs >= 0 && neg_num >= 1, "triplet_sample: n_pairs=%lld neg_num=%d", (long long)n_pairs, neg_num);
    if (n_pairs == 0) return R4D_OK;
    R4D_REQUIRE(pos_row && row_start && neg && n_neg && choice, "triplet_sample: null pointer");
    int64_t blocks = (n_pairs + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    triplet_sample_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(pos_row, row_start, n_pairs, neg, n_neg, neg_num,
                                                                         seed, choice); note_launch();
    R4D_CUDA(cudaGetLastError());
    return R4D_OK;
}

int r4d_triplet_mine_f64(const double* out, const double* in, int64_t n, int64_t ld, double thr, int32_t neg_num,
                         int32_t* n_pos, int32_t* neg, int32_t* n_neg, r4d_stream_t stream) {
    using namespace r4d;
    R4D_REQUIRE(n >= 0 && ld >= n && neg_num >= 1 && neg_num <= R4D_TOPK_MAX, "triplet_mine: n=%lld neg_num=%d",
                (long long)n, neg_num);
    if (n == 0) return R4D_OK;
    R4D_REQUIRE(out && in && n_pos && neg && n_neg, "triplet_mine: null pointer");
    int64_t blocks = (n + 7) / 8;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    triplet_mine_f64_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(out, in, n, ld, thr, neg_num, n_pos, neg,
                                                                          n_neg); note_launch();
    R4D_CUDA(cudaGetLastError());
    return R4D_OK;
}

}  // extern "C"
