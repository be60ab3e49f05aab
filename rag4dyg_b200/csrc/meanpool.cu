// meanpool.cu — pool-embedding producer feeding the dense scorer.
// Replaces `h_egos = torch.mean(h, dim=1)` (train/train_retriever.py:420, :432) followed by the per-batch
// re-normalisation (:433, :436): hidden states [B, L, D] fp32 -> mean over the PADDED length L (pads included, like the
// reference) -> L2-normalise -> bf16 hi (/lo) planes in the scorer's layout, without the [B, D] round trips.
// HBM bound: reads B*L*D*4 bytes once; two deterministic passes (fixed summation order, no float atomics).
#include "dense_common.cuh"

namespace r4d {

constexpr int MP_LSPLIT = 8;  // CTAs along L per batch row

// grid (B, MP_LSPLIT): partial[b][s][d] = sum over this CTA's slice of L
__global__ void __launch_bounds__(256)
meanpool_partial_kernel(const float* __restrict__ h, int64_t B, int32_t L, int32_t D, float* __restrict__ partial) {
    const int64_t b = blockIdx.x;
    const int s = blockIdx.y;
    const int l0 = (int)((int64_t)L * s / MP_LSPLIT), l1 = (int)((int64_t)L * (s + 1) / MP_LSPLIT);
    const float* base = h + (b * L) * (int64_t)D;
    if ((D & 3) == 0) {
        for (int d = threadIdx.x * 4; d < D; d += blockDim.x * 4) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int l = l0; l < l1; ++l) {
                const float4 v = *reinterpret_cast<const float4*>(base + (int64_t)l * D + d);
                acc.x += v.x;
                acc.y += v.y;
                acc.z += v.z;
                acc.w += v.w;
            }
            *reinterpret_cast<float4*>(partial + ((b * MP_LSPLIT + s) * (int64_t)D + d)) = acc;
        }
    } else {
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            float acc = 0.f;
            for (int l = l0; l < l1; ++l) acc += base[(int64_t)l * D + d];
            partial[(b * MP_LSPLIT + s) * (int64_t)D + d] = acc;
        }
    }
}

// one warp per row: mean = (sum of partials in fixed order) / L; optional fp32 copy; normalise; bf16 split
__global__ void __launch_bounds__(256)
meanpool_finish_kernel(const float* __restrict__ partial, int64_t B, int32_t L, int32_t D, int32_t d_pad,
                       int32_t want_lo, float* __restrict__ mean_out, __nv_bfloat16* __restrict__ hi,
                       __nv_bfloat16* __restrict__ lo) {
    const int lane = threadIdx.x & 31;
    const int64_t wpg = (int64_t)gridDim.x * (blockDim.x >> 5);
    const float inv_l = 1.0f / (float)L;
    for (int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < B; b += wpg) {
        float ss = 0.f;
        for (int d = lane; d < D; d += 32) {
            float acc = 0.f;
#pragma unroll
            for (int s = 0; s < MP_LSPLIT; ++s) acc += partial[(b * MP_LSPLIT + s) * (int64_t)D + d];
            const float m = acc * inv_l;
            ss = fmaf(m, m, ss);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float nrm = sqrtf(ss);
        for (int d = lane; d < d_pad; d += 32) {
            float y = 0.f;
            if (d < D) {
                float acc = 0.f;
#pragma unroll
                for (int s = 0; s < MP_LSPLIT; ++s) acc += partial[(b * MP_LSPLIT + s) * (int64_t)D + d];
                const float m = acc * inv_l;
                if (mean_out) mean_out[b * (int64_t)D + d] = m;
                y = m / nrm;
            }
            const __nv_bfloat16 hv = __float2bfloat16_rn(y);
            hi[b * (int64_t)d_pad + d] = hv;
            if (want_lo) lo[b * (int64_t)d_pad + d] = __float2bfloat16_rn(y - __bfloat162float(hv));
        }
    }
}

}  // namespace r4d

extern "C" {

size_t r4d_meanpool_workspace_bytes(int64_t batch, int32_t d) {
    if (batch <= 0 || d <= 0) return 256;
    return (size_t)batch * r4d::MP_LSPLIT * (size_t)d * sizeof(float) + 256;
}

int r4d_meanpool_prepare(const float* hidden, int64_t batch, int32_t len, int32_t d, int32_t prec, float* mean_out,
                         void* hi, void* lo, void* workspace, size_t workspace_bytes, r4d_stream_t stream) {
    using namespace r4d;
    R4D_REQUIRE(batch >= 0 && len > 0 && d > 0, "meanpool: batch=%lld len=%d d=%d", (long long)batch, len, d);
    R4D_REQUIRE(prec == R4D_PREC_BF16 || prec == R4D_PREC_BF16X3, "meanpool: unknown precision %d", prec);
    if (batch == 0) return R4D_OK;
    R4D_REQUIRE(hidden && hi && (prec == R4D_PREC_BF16 || lo) && workspace, "meanpool: null pointer");
    R4D_REQUIRE(batch < 65536 * 32768ll, "meanpool: batch too large");
    if (workspace_bytes < r4d_meanpool_workspace_bytes(batch, d)) {
        set_error("meanpool: workspace %zu B < required %zu B", workspace_bytes, r4d_meanpool_workspace_bytes(batch, d));
        return R4D_E_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    float* partial = reinterpret_cast<float*>(workspace);
    for (int64_t b0 = 0; b0 < batch; b0 += 2147483647ll) {  // gridDim.x limit (never hit in practice)
        const int64_t nb = batch - b0 < 2147483647ll ? batch - b0 : 2147483647ll;
        meanpool_partial_kernel<<<dim3((unsigned)nb, MP_LSPLIT), 256, 0, st>>>(hidden + b0 * len * (int64_t)d, nb, len, d,
                                                                              partial + b0 * MP_LSPLIT * (int64_t)d); note_launch();
    }
    R4D_CUDA(cudaGetLastError());
    int64_t blocks = (batch + 7) / 8;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    meanpool_finish_kernel<<<(unsigned)blocks, 256, 0, st>>>(partial, batch, len, d, r4d_dense_dpad(d),
                                                             prec == R4D_PREC_BF16X3 ? 1 : 0, mean_out,
                                                             reinterpret_cast<__nv_bfloat16*>(hi),
                                                             reinterpret_cast<__nv_bfloat16*>(lo)); note_launch();
    R4D_CUDA(cudaGetLastError());
    return R4D_OK;
}

}  // extern "C"
