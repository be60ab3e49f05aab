// bitset_encode.cu — subsystem 1: CSR of bit positions -> fixed-width uint32 bitsets + cardinalities.
// Replaces the per-pair set(seq) construction of co_occurrence_ratio (retrieval_data_annotation.py:12-13):
// set semantics (duplicates collapse) come from OR-ing, |set| from counting first-time bit sets.
//
// HBM traffic: one memset of the bitset matrix (rows * pitch * 4 B written) + CSR read + one atomic per element.
#include "r4d_common.cuh"

namespace r4d {

// One warp per row.  atomicOr returns the previous word, so "bit was clear before" counts distinct members
// without a second pass over the bitsets.
__global__ void __launch_bounds__(256) bitset_scatter_kernel(const int32_t* __restrict__ bit_pos,
                                                              const int64_t* __restrict__ row_off, int64_t n_rows,
                                                              int32_t n_bits, int32_t pitch_words,
                                                              uint32_t* __restrict__ bits, uint32_t* __restrict__ card) {
    const int64_t warps_per_grid = (int64_t)gridDim.x * (blockDim.x >> 5);
    const uint32_t lane = threadIdx.x & 31;
    for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n_rows;
         row += warps_per_grid) {
        const int64_t beg = row_off[row], end = row_off[row + 1];
        uint32_t* rbits = bits + row * (int64_t)pitch_words;
        uint32_t cnt = 0;
        for (int64_t e = beg + lane; e < end; e += 32) {
            const int32_t b = bit_pos[e];
            if (b >= 0 && b < n_bits) {
                const uint32_t m = 1u << (b & 31);
                const uint32_t old = atomicOr(rbits + (b >> 5), m);
                cnt += (old & m) ? 0u : 1u;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) card[row] = cnt;
    }
}

}  // namespace r4d

extern "C" {

int32_t r4d_bitset_words(int32_t n_bits) { return n_bits <= 0 ? 0 : (n_bits + 31) / 32; }

int32_t r4d_bitset_pitch_words(int32_t n_bits) {
    int32_t w = r4d_bitset_words(n_bits);
    if (w == 0) w = 1;
    return (w + 31) / 32 * 32;  // 128-byte rows: one TMA swizzle atom per 32-word chunk
}

int r4d_bitset_encode(const int32_t* bit_pos, const int64_t* row_off, int64_t n_rows, int32_t n_bits,
                      int32_t pitch_words, uint32_t* bits, uint32_t* card, r4d_stream_t stream) {
    using namespace r4d;
    R4D_REQUIRE(n_rows >= 0 && n_bits > 0, "bitset_encode: n_rows=%lld n_bits=%d", (long long)n_rows, n_bits);
    R4D_REQUIRE(pitch_words >= r4d_bitset_words(n_bits) && pitch_words % 4 == 0,
                "bitset_encode: pitch_words=%d too small or not a multiple of 4 for n_bits=%d", pitch_words, n_bits);
    if (n_rows == 0) return R4D_OK;
    R4D_REQUIRE(row_off && bits && card, "bitset_encode: null pointer");
    cudaStream_t st = as_stream(stream);
    R4D_CUDA(cudaMemsetAsync(bits, 0, (size_t)n_rows * (size_t)pitch_words * sizeof(uint32_t), st));
    const int warps_per_block = 8;
    int64_t blocks = (n_rows + warps_per_block - 1) / warps_per_block;
    const int64_t cap = (int64_t)num_sms() * 32;
    if (blocks > cap) blocks = cap;
    bitset_scatter_kernel<<<(unsigned)blocks, warps_per_block * 32, 0, st>>>(bit_pos, row_off, n_rows, n_bits,
                                                                           pitch_words, bits, card); note_launch();
    R4D_CUDA(cudaGetLastError());
    return R4D_OK;
}

}  // extern "C"
