// text_format_device.cu — the reference's text files assembled ON THE DEVICE (producer half).
// After GPU scoring, writing `' '.join(str(x) for x in row) + '\n'` for every row (retrieval_data_annotation.py:92-93,
// train/train_retriever.py:362-363) is what is left of the wall time: dialog's val/test rankings are 19 M integers and
// 19 M floats, 170 MB of text.  Here the rankings / score codes never leave HBM as numbers: two passes turn them into
// the final bytes, the host copies the text out and writes it.
//   r4d_format_rows_device_sizes : row_off[q] = byte offset of row q in the file, row_off[nq] = file size
//   r4d_format_rows_device       : the bytes.  Decimal int32 fields, or (lut != NULL) fields looked up in a string table
//                                  the caller built with the reference's own float formatter — one entry per DISTINCT
//                                  value — so float text is byte-identical by construction.
// One CTA per row; a thread formats 4 consecutive fields, a block-wide exclusive scan of the field lengths places them.
#include "r4d_common.cuh"

namespace r4d {

constexpr int TF_THREADS = 256;
constexpr int TF_ITEMS = 4;
constexpr int TF_TILE = TF_THREADS * TF_ITEMS;

struct TFSrc {
    const int32_t* vals;
    int64_t nq, n, ld;
    const char* blob;          // LUT mode: concatenated strings (device)
    const int64_t* lut_off;    // LUT mode: n_codes + 1 offsets into blob (device); NULL = decimal integers
    int32_t n_codes;
    int32_t* status;           // set to 1 when a code is out of range (the field is then formatted as code 0)
};

__device__ __forceinline__ int dec_len(int32_t v) {
    const uint32_t a = v < 0 ? 0u - (uint32_t)v : (uint32_t)v;
    return 1 + (a >= 10u) + (a >= 100u) + (a >= 1000u) + (a >= 10000u) + (a >= 100000u) + (a >= 1000000u) + (a >= 10000000u) +
           (a >= 100000000u) + (a >= 1000000000u) + (v < 0);
}

template <bool LUT>
__device__ __forceinline__ int field_len(const TFSrc& s, int32_t& v) {
    if (!LUT) return dec_len(v);
    if (v < 0 || v >= s.n_codes) {
        if (s.status) *s.status = 1;
        v = 0;
    }
    return (int)(s.lut_off[v + 1] - s.lut_off[v]);
}

// exclusive scan of one value per thread over the CTA; `total` = the block's sum
__device__ __forceinline__ uint32_t tf_block_scan(uint32_t v, uint32_t* warp_tot /*[TF_THREADS / 32]*/, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();   // warp_tot of the previous call has been read
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < TF_THREADS / 32; ++w) {
        const uint32_t t = warp_tot[w];
        if (w < warp) base += t;
        tot += t;
    }
    total = tot;
    return base + inc - v;
}

template <bool LUT>
__global__ void __launch_bounds__(TF_THREADS) tf_row_len_kernel(const TFSrc s, int64_t* __restrict__ row_len) {
    __shared__ uint32_t warp_tot[TF_THREADS / 32];
    for (int64_t q = blockIdx.x; q < s.nq; q += gridDim.x) {
        const int32_t* r = s.vals + q * s.ld;
        uint64_t sum = 0;
        for (int64_t j = threadIdx.x; j < s.n; j += TF_THREADS) {
            int32_t v = r[j];
            sum += (uint64_t)field_len<LUT>(s, v) + 1u;          // the field and its separator (' ' or the final '\n')
        }
        // block sum (two 32-bit halves keep the scan helper simple; a row is far below 2^32 bytes per thread)
        uint32_t total;
        tf_block_scan((uint32_t)sum, warp_tot, total);
        if (threadIdx.x == 0) row_len[q] = s.n > 0 ? (int64_t)total : 1;   // an empty row is a lone '\n'
        __syncthreads();
    }
}

// row_off[0] = 0, row_off[q + 1] = row_off[q] + len[q]; in place is fine (len == row_off + 1).  Single CTA.
__global__ void __launch_bounds__(1024) tf_scan_rows_kernel(const int64_t* len, int64_t nq, int64_t* off) {
    __shared__ int64_t warp_tot[32];
    __shared__ int64_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        carry_s = 0;
        off[0] = 0;
    }
    __syncthreads();
    for (int64_t base = 0; base < nq; base += 1024) {
        const int64_t q = base + threadIdx.x;
        int64_t inc = q < nq ? len[q] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        int64_t pre = carry_s;
        for (int w = 0; w < warp; ++w) pre += warp_tot[w];
        if (q < nq) off[q + 1] = pre + inc;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = pre + inc;
        __syncthreads();
    }
}

template <bool LUT>
__global__ void __launch_bounds__(TF_THREADS) tf_write_kernel(const TFSrc s, const int64_t* __restrict__ row_off,
                                                             char* __restrict__ text) {
    __shared__ uint32_t warp_tot[TF_THREADS / 32];
    for (int64_t q = blockIdx.x; q < s.nq; q += gridDim.x) {
        const int32_t* r = s.vals + q * s.ld;
        char* out = text + row_off[q];
        if (s.n == 0) {
            if (threadIdx.x == 0) out[0] = '\n';
            continue;
        }
        uint64_t run = 0;                                  // bytes of the row written by earlier tiles
        for (int64_t t0 = 0; t0 < s.n; t0 += TF_TILE) {
            const int64_t j0 = t0 + (int64_t)threadIdx.x * TF_ITEMS;
            int32_t v[TF_ITEMS];
            int len[TF_ITEMS];
            uint32_t mine = 0;
#pragma unroll
            for (int e = 0; e < TF_ITEMS; ++e) {
                len[e] = 0;
                if (j0 + e < s.n) {
                    v[e] = r[j0 + e];
                    len[e] = field_len<LUT>(s, v[e]);
                    mine += (uint32_t)len[e] + 1u;
                }
            }
            uint32_t total;
            const uint32_t at = tf_block_scan(mine, warp_tot, total);
            char* p = out + run + at;
#pragma unroll
            for (int e = 0; e < TF_ITEMS; ++e) {
                if (j0 + e < s.n) {
                    if (LUT) {
                        const char* src = s.blob + s.lut_off[v[e]];
                        for (int c = 0; c < len[e]; ++c) p[c] = src[c];
                    } else {
                        uint32_t a = v[e] < 0 ? 0u - (uint32_t)v[e] : (uint32_t)v[e];
                        for (int c = len[e] - 1; c >= (v[e] < 0 ? 1 : 0); --c) {
                            p[c] = (char)('0' + a % 10u);
                            a /= 10u;
                        }
                        if (v[e] < 0) p[0] = '-';
                    }
                    p[len[e]] = (j0 + e == s.n - 1) ? '\n' : ' ';
                    p += len[e] + 1;
                }
            }
            run += total;
        }
        __syncthreads();
    }
}

static int tf_check(const int32_t* vals, int64_t nq, int64_t n, int64_t ld, const int64_t* lut_off, int32_t n_codes) {
    R4D_REQUIRE(nq >= 0 && n >= 0 && ld >= n, "format_rows_device: nq=%lld n=%lld ld=%lld", (long long)nq, (long long)n, (long long)ld);
    R4D_REQUIRE(vals || nq * n == 0, "format_rows_device: null values");
    R4D_REQUIRE(!lut_off || n_codes > 0, "format_rows_device: empty string table");
    return R4D_OK;
}

static unsigned tf_grid(int64_t nq) {
    const int64_t cap = (int64_t)num_sms() * 8;
    return (unsigned)(nq < cap ? (nq > 0 ? nq : 1) : cap);
}

}  // namespace r4d

extern "C" {

int r4d_format_rows_device_sizes(const int32_t* vals, int64_t nq, int64_t n, int64_t ld, const int64_t* lut_off,
                                 int32_t n_codes, int64_t* row_off, int32_t* status, r4d_stream_t stream) {
    using namespace r4d;
    if (int rc = tf_check(vals, nq, n, ld, lut_off, n_codes)) return rc;
    R4D_REQUIRE(row_off, "format_rows_device_sizes: null row_off");
    cudaStream_t st = as_stream(stream);
    if (status) R4D_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    TFSrc s{vals, nq, n, ld, nullptr, lut_off, n_codes, status};
    if (nq > 0) {
        if (lut_off) tf_row_len_kernel<true><<<tf_grid(nq), TF_THREADS, 0, st>>>(s, row_off + 1);
        else tf_row_len_kernel<false><<<tf_grid(nq), TF_THREADS, 0, st>>>(s, row_off + 1);
        note_launch();
    }
    tf_scan_rows_kernel<<<1, 1024, 0, st>>>(row_off + 1, nq, row_off); note_launch();
    R4D_CUDA(cudaGetLastError());
    return R4D_OK;
}

int r4d_format_rows_device(const int32_t* vals, int64_t nq, int64_t n, int64_t ld, const char* lut_blob, const int64_t* lut_off,
                           int32_t n_codes, const int64_t* row_off, char* text, r4d_stream_t stream) {
    using namespace r4d;
    if (int rc = tf_check(vals, nq, n, ld, lut_off, n_codes)) return rc;
    R4D_REQUIRE((lut_blob != nullptr) == (lut_off != nullptr), "format_rows_device: lut_blob and lut_off go together");
    if (nq == 0) return R4D_OK;
    R4D_REQUIRE(row_off && text, "format_rows_device: null pointer");
    cudaStream_t st = as_stream(stream);
    TFSrc s{vals, nq, n, ld, lut_blob, lut_off, n_codes, nullptr};
    if (lut_off) tf_write_kernel<true><<<tf_grid(nq), TF_THREADS, 0, st>>>(s, row_off, text);
    else tf_write_kernel<false><<<tf_grid(nq), TF_THREADS, 0, st>>>(s, row_off, text);
    note_launch();
    R4D_CUDA(cudaGetLastError());
    return R4D_OK;
}

}  // extern "C"
