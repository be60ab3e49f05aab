// triplet_mine.cu — save_train_annotation's device half straight from the OUT (label) and IN (history) BITSETS
// (retrieval_data_annotation.py:43-71 with both train matrices' diagonals zeroed as :172-173).  Nothing [N, N] is
// ever written to HBM: a CTA owns 16 anchor rows i, walks the pool in 128-row tiles, forms the intersection counts of
// BOTH set families for every (i, j) of the tile in registers (AND + POPC over 64-word chunks staged in shared
// memory, all-zero anchor words skipped) and consumes them at once:
//   positives   out[i,j] > thr                 -> appended to a global list (key = i << 32 | j, counts = inter << 32 | union)
//   hard negs   out[i,j] <= thr, out[i,j] > 0  -> per-row top-neg_num by in[i,j] (exact rational compare, index asc on ties)
//   fill negs   out[i,j] == 0                  -> same ranking, used while a row has fewer than neg_num hard ones
// which is exactly the reference's walk over np.argsort(-in[i], kind='stable') (:56-71).  The negatives come out with
// their OUT counts, so the score file needs no matrix either.
//
// Also here: r4d_mt19937_choice_replay, the host-side (CPU) replay of the reference's np.random.choice(neg) calls (:79)
// on numpy's legacy MT19937 state — the one sequential piece of the stage, done natively instead of one Python call
// per triplet.
#include "r4d_common.cuh"

namespace r4d {

constexpr int TM_THREADS = 256;
constexpr int TM_WARPS = TM_THREADS / 32;
constexpr int TM_RI = 2;                      // anchor rows per warp
constexpr int TM_TI = TM_WARPS * TM_RI;       // 16 anchor rows per CTA
constexpr int TM_SJ = 4;                      // pool rows per lane and tile
constexpr int TM_TJ = 32 * TM_SJ;             // 128 pool rows per tile
constexpr int TM_KC = 64;                     // words per staged chunk
constexpr int TM_PITCH = TM_KC + 4;           // odd multiple of 16 B: the LDS.128 of 8 consecutive rows hit distinct banks

struct TMParams {
    const uint32_t* bits[2];   // [0] OUT (label) sets, [1] IN (history) sets
    const uint32_t* card[2];
    int32_t words[2], pitch[2];
    int64_t n;
    double thr;
    int32_t neg_num, zero_diag;
    int32_t* n_pos;
    long long* pos_key;
    long long* pos_counts;
    int64_t pos_cap;
    unsigned long long* pos_total;
    int32_t* neg;
    uint32_t* neg_inter;
    uint32_t* neg_union;
    int32_t* n_neg;
};

// negative candidate: ranked by the IN score si/su (exact), ties by ascending index; carries its OUT counts
struct NEntry {
    uint32_t si, su;
    int32_t idx;
    uint32_t oi, ou;
    __device__ __forceinline__ static NEntry worst() { return NEntry{0u, 1u, R4D_IDX_NONE, 0u, 1u}; }
    __device__ __forceinline__ static bool better(const NEntry& a, const NEntry& b) {
        const uint64_t l = (uint64_t)a.si * (uint64_t)b.su;
        const uint64_t r = (uint64_t)b.si * (uint64_t)a.su;
        return (l > r) || (l == r && a.idx < b.idx);
    }
    __device__ __forceinline__ NEntry shfl(int src) const {
        return NEntry{__shfl_sync(0xffffffffu, si, src), __shfl_sync(0xffffffffu, su, src), __shfl_sync(0xffffffffu, idx, src),
                      __shfl_sync(0xffffffffu, oi, src), __shfl_sync(0xffffffffu, ou, src)};
    }
    __device__ __forceinline__ NEntry shfl_up1() const {
        return NEntry{__shfl_up_sync(0xffffffffu, si, 1), __shfl_up_sync(0xffffffffu, su, 1), __shfl_up_sync(0xffffffffu, idx, 1),
                      __shfl_up_sync(0xffffffffu, oi, 1), __shfl_up_sync(0xffffffffu, ou, 1)};
    }
};

__device__ __forceinline__ uint32_t popc4(const uint4& a, const uint4& b) {
    return __popc(a.x & b.x) + __popc(a.y & b.y) + __popc(a.z & b.z) + __popc(a.w & b.w);
}

__global__ void __launch_bounds__(TM_THREADS) triplet_mine_bits_kernel(const TMParams p) {
    __shared__ __align__(16) uint32_t s_i[TM_TI * TM_PITCH];
    __shared__ __align__(16) uint32_t s_j[TM_TJ * TM_PITCH];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n = p.n;
    for (int64_t i0 = (int64_t)blockIdx.x * TM_TI; i0 < n; i0 += (int64_t)gridDim.x * TM_TI) {
        WarpTopK<NEntry> hard[TM_RI], fill[TM_RI];
        int32_t pos_cnt[TM_RI];
        int64_t ia[TM_RI];
        uint32_t ci[2][TM_RI];
#pragma unroll
        for (int r = 0; r < TM_RI; ++r) {
            hard[r].init(p.neg_num);
            fill[r].init(p.neg_num);
            pos_cnt[r] = 0;
            ia[r] = i0 + warp * TM_RI + r;
            ci[0][r] = ia[r] < n ? p.card[0][ia[r]] : 0u;
            ci[1][r] = ia[r] < n ? p.card[1][ia[r]] : 0u;
        }
        for (int64_t j0 = 0; j0 < n; j0 += TM_TJ) {
            uint32_t acc[2][TM_RI][TM_SJ];
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int r = 0; r < TM_RI; ++r)
#pragma unroll
                    for (int s = 0; s < TM_SJ; ++s) acc[m][r][s] = 0u;
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const uint32_t* __restrict__ bits = p.bits[m];
                const int32_t pitch = p.pitch[m];
                const int32_t words = p.words[m];
                for (int32_t c0 = 0; c0 < words; c0 += TM_KC) {
                    __syncthreads();   // the previous chunk has been consumed
                    for (int u = tid; u < (TM_TI + TM_TJ) * (TM_KC / 4); u += TM_THREADS) {
                        const int row = u / (TM_KC / 4), w4 = u % (TM_KC / 4);
                        const int64_t g = row < TM_TI ? i0 + row : j0 + (row - TM_TI);
                        const int32_t w = c0 + w4 * 4;
                        uint4 v = make_uint4(0u, 0u, 0u, 0u);
                        if (g < n && w < pitch) v = __ldg(reinterpret_cast<const uint4*>(bits + g * pitch + w));
                        uint32_t* dst = row < TM_TI ? s_i + row * TM_PITCH + w4 * 4 : s_j + (row - TM_TI) * TM_PITCH + w4 * 4;
                        *reinterpret_cast<uint4*>(dst) = v;
                    }
                    __syncthreads();
                    const uint32_t* a0 = s_i + (warp * TM_RI) * TM_PITCH;
#pragma unroll 4
                    for (int w4 = 0; w4 < TM_KC / 4; ++w4) {
                        const uint4 x0 = *reinterpret_cast<const uint4*>(a0 + w4 * 4);
                        const uint4 x1 = *reinterpret_cast<const uint4*>(a0 + TM_PITCH + w4 * 4);
                        // node-id sets are sparse: most 128-bit groups of the two anchor rows are empty (warp-uniform skip)
                        if (((x0.x | x0.y | x0.z | x0.w) | (x1.x | x1.y | x1.z | x1.w)) == 0u) continue;
#pragma unroll
                        for (int s = 0; s < TM_SJ; ++s) {
                            const uint4 y = *reinterpret_cast<const uint4*>(s_j + (s * 32 + lane) * TM_PITCH + w4 * 4);
                            acc[m][0][s] += popc4(x0, y);
                            acc[m][1][s] += popc4(x1, y);
                        }
                    }
                }
            }
            // ---- the tile's counts are complete for both set families: consume them
#pragma unroll
            for (int s = 0; s < TM_SJ; ++s) {
                const int64_t j = j0 + s * 32 + lane;
                const uint32_t cjo = j < n ? p.card[0][j] : 0u, cji = j < n ? p.card[1][j] : 0u;
#pragma unroll
                for (int r = 0; r < TM_RI; ++r) {
                    const bool valid = ia[r] < n && j < n;
                    uint32_t oi = acc[0][r][s], ii = acc[1][r][s];
                    if (p.zero_diag && ia[r] == j) oi = ii = 0u;                      // :172-173
                    const uint32_t ou = ci[0][r] + cjo - acc[0][r][s], iu = ci[1][r] + cji - acc[1][r][s];
                    const double o = ou ? (double)oi / (double)ou : 0.0;              // == Python's len(a & b) / len(a | b)
                    const bool is_pos = valid && o > p.thr;                           // :54, strict >
                    const bool is_hard = valid && !is_pos && o > 0.0;                 // :60
                    const bool is_fill = valid && !is_pos && o == 0.0;                // :67
                    const uint32_t pm = __ballot_sync(0xffffffffu, is_pos);
                    if (pm) {
                        unsigned long long base = 0;
                        if (lane == 0) base = atomicAdd(p.pos_total, (unsigned long long)__popc(pm));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        if (is_pos) {
                            const unsigned long long at = base + __popc(pm & ((1u << lane) - 1u));
                            if ((int64_t)at < p.pos_cap) {
                                p.pos_key[at] = (long long)(((unsigned long long)ia[r] << 32) | (unsigned long long)j);
                                p.pos_counts[at] = (long long)(((unsigned long long)oi << 32) | (unsigned long long)ou);
                            }
                        }
                        pos_cnt[r] += __popc(pm);
                    }
                    const NEntry c{ii, iu ? iu : 1u, (int32_t)j, oi, ou ? ou : 1u};
                    uint32_t hm = __ballot_sync(0xffffffffu, is_hard && NEntry::better(c, hard[r].kth));
                    while (hm) {
                        const int src = __ffs(hm) - 1;
                        hm &= hm - 1;
                        hard[r].insert(c.shfl(src));
                    }
                    uint32_t fm = __ballot_sync(0xffffffffu, is_fill && NEntry::better(c, fill[r].kth));
                    while (fm) {
                        const int src = __ffs(fm) - 1;
                        fm &= fm - 1;
                        fill[r].insert(c.shfl(src));
                    }
                }
            }
        }
        // ---- negatives of the two anchor rows: hard ones first, then fillers (:57-71)
#pragma unroll
        for (int r = 0; r < TM_RI; ++r) {
            const int n_hard = __popc(__ballot_sync(0xffffffffu, lane < p.neg_num && hard[r].mine.idx != R4D_IDX_NONE));
            const int n_fill = __popc(__ballot_sync(0xffffffffu, lane < p.neg_num && fill[r].mine.idx != R4D_IDX_NONE));
            const int total = min(p.neg_num, n_hard + n_fill);
            const int fsrc = lane - n_hard;
            const NEntry ff = fill[r].mine.shfl(fsrc < 0 ? 0 : (fsrc > 31 ? 31 : fsrc));
            const NEntry e = lane < n_hard ? hard[r].mine : ff;
            if (ia[r] < n) {
                if (lane < p.neg_num) {
                    const int64_t at = ia[r] * p.neg_num + lane;
                    const bool have = lane < total;
                    p.neg[at] = have ? e.idx : -1;
                    p.neg_inter[at] = have ? e.oi : 0u;
                    p.neg_union[at] = have ? e.ou : 1u;
                }
                if (lane == 0) {
                    p.n_pos[ia[r]] = pos_cnt[r];
                    p.n_neg[ia[r]] = total;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------- numpy legacy RNG replay (host)
// np.random.choice(a) of the legacy global RandomState == a[randint(0, len(a))]; randint draws 32-bit outputs of
// MT19937, masks them with the smallest 2^b - 1 >= len(a) - 1 and rejects values above len(a) - 1; a one-element
// array consumes no randomness.  (numpy/random/mtrand.pyx choice -> randint -> _bounded_integers ->
// random_bounded_uint64_fill / buffered_bounded_masked_uint32 with use_masked = true.)
struct MT19937 {
    uint32_t* key;
    int pos;
    void regenerate() {
        constexpr int N = 624, M = 397;
        constexpr uint32_t MATRIX_A = 0x9908b0dfu, UPPER = 0x80000000u, LOWER = 0x7fffffffu;
        uint32_t y;
        int kk = 0;
        for (; kk < N - M; ++kk) {
            y = (key[kk] & UPPER) | (key[kk + 1] & LOWER);
            key[kk] = key[kk + M] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
        }
        for (; kk < N - 1; ++kk) {
            y = (key[kk] & UPPER) | (key[kk + 1] & LOWER);
            key[kk] = key[kk + (M - N)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
        }
        y = (key[N - 1] & UPPER) | (key[0] & LOWER);
        key[N - 1] = key[M - 1] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
        pos = 0;
    }
    uint32_t next32() {
        if (pos >= 624) regenerate();
        uint32_t y = key[pos++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
};

}  // namespace r4d

extern "C" {

int r4d_triplet_mine(const uint32_t* obits, const uint32_t* ocard, int32_t owords, int32_t opitch, const uint32_t* ibits,
                     const uint32_t* icard, int32_t iwords, int32_t ipitch, int64_t n, double thr, int32_t neg_num,
                     int32_t zero_diag, int32_t* n_pos, int64_t* pos_key, int64_t* pos_counts, int64_t pos_cap,
                     int64_t* pos_total, int32_t* neg, uint32_t* neg_inter, uint32_t* neg_union, int32_t* n_neg,
                     r4d_stream_t stream) {
    using namespace r4d;
    R4D_REQUIRE(n >= 0 && n < ((int64_t)1 << 31) && neg_num >= 1 && neg_num <= R4D_TOPK_MAX, "triplet_mine: n=%lld neg_num=%d",
                (long long)n, neg_num);
    R4D_REQUIRE(owords > 0 && iwords > 0 && opitch >= owords && ipitch >= iwords && opitch % 4 == 0 && ipitch % 4 == 0,
                "triplet_mine: words/pitch (out %d/%d, in %d/%d; pitches must be multiples of 4 words)", owords, opitch, iwords,
                ipitch);
    R4D_REQUIRE(pos_total && pos_cap >= 0 && (pos_cap == 0 || (pos_key && pos_counts)), "triplet_mine: positives buffers");
    cudaStream_t st = as_stream(stream);
    R4D_CUDA(cudaMemsetAsync(pos_total, 0, sizeof(int64_t), st));
    if (n == 0) return R4D_OK;
    R4D_REQUIRE(obits && ocard && ibits && icard && n_pos && neg && neg_inter && neg_union && n_neg, "triplet_mine: null pointer");
    R4D_REQUIRE(((reinterpret_cast<uintptr_t>(obits) | reinterpret_cast<uintptr_t>(ibits)) & 15) == 0,
                "triplet_mine: bitsets must be 16-byte aligned");
    TMParams p{};
    p.bits[0] = obits;
    p.bits[1] = ibits;
    p.card[0] = ocard;
    p.card[1] = icard;
    p.words[0] = owords;
    p.words[1] = iwords;
    p.pitch[0] = opitch;
    p.pitch[1] = ipitch;
    p.n = n;
    p.thr = thr;
    p.neg_num = neg_num;
    p.zero_diag = zero_diag;
    p.n_pos = n_pos;
    p.pos_key = reinterpret_cast<long long*>(pos_key);
    p.pos_counts = reinterpret_cast<long long*>(pos_counts);
    p.pos_cap = pos_cap;
    p.pos_total = reinterpret_cast<unsigned long long*>(pos_total);
    p.neg = neg;
    p.neg_inter = neg_inter;
    p.neg_union = neg_union;
    p.n_neg = n_neg;
    int64_t blocks = (n + TM_TI - 1) / TM_TI;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    triplet_mine_bits_kernel<<<(unsigned)blocks, TM_THREADS, 0, st>>>(p); note_launch();
    R4D_CUDA(cudaGetLastError());
    return R4D_OK;
}

int r4d_mt19937_choice_replay(uint32_t* key, int32_t* pos, const int32_t* sizes, int64_t n_calls, int32_t* choice) {
    using namespace r4d;
    R4D_REQUIRE(key && pos && (n_calls == 0 || (sizes && choice)) && n_calls >= 0, "mt19937_choice_replay: null pointer");
    R4D_REQUIRE(*pos >= 0 && *pos <= 624, "mt19937_choice_replay: state position %d", *pos);
    MT19937 mt{key, *pos};
    for (int64_t t = 0; t < n_calls; ++t) {
        const int32_t m = sizes[t];
        if (m <= 0) {
            set_error("mt19937_choice_replay: call %lld draws from an empty array", (long long)t);
            *pos = mt.pos;
            return R4D_E_ARG;
        }
        const uint32_t rng = (uint32_t)(m - 1);
        if (rng == 0) {                  // randint(0, 1): no randomness consumed
            choice[t] = 0;
            continue;
        }
        uint32_t mask = rng;
        mask |= mask >> 1;
        mask |= mask >> 2;
        mask |= mask >> 4;
        mask |= mask >> 8;
        mask |= mask >> 16;
        uint32_t v;
        do {
            v = mt.next32() & mask;
        } while (v > rng);
        choice[t] = (int32_t)v;
    }
    *pos = mt.pos;
    return R4D_OK;
}

}  // extern "C"
