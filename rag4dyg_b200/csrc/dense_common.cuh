// dense_common.cuh — tcgen05 / TMEM PTX wrappers and tile constants shared by the dense scorer kernels.
#pragma once
#include <cuda_bf16.h>

#include "r4d_common.cuh"

namespace r4d {

constexpr int DQ = 128;          // query rows per CTA tile (UMMA M per CTA)
constexpr int DKB = 64;          // bf16 elements per k-block (128 B rows: one swizzle atom)
constexpr int Q_TILE_BYTES = DQ * DKB * 2;  // 16 KB

// ---- tcgen05 wrappers (cta_group::1)
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, SWIZZLE_128B operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (=1), version 1.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
        "%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}


// epilogue score from the raw cosine x (operands are pre-normalised): mode 0 (x+1)/2, 1 x*decay, 2 ((x+1)/2)*decay
__device__ __forceinline__ float dense_score(float x, int mode, float tq, float tp, float neg_lambda_log2e) {
    if (mode == R4D_DENSE_HALF_COS) return (x + 1.0f) * 0.5f;
    const float dec = ex2_approx(neg_lambda_log2e * fabsf(tq - tp));
    return (mode == R4D_DENSE_COS_DECAY) ? x * dec : (x + 1.0f) * 0.5f * dec;
}

constexpr int D_EPI_THREADS = 256;  // 8 epilogue warps in both dense kernels

// sorted insertion into this thread's list (column-major in smem: slot t of thread `tid` at [t*D_EPI_THREADS+tid]).
// Returns the new k-th best score.
static __device__ __noinline__ float list_insert(float* ls, int32_t* li, int k, float s, int32_t idx) {
    int t = k - 1;
    while (t > 0) {
        const float prev = ls[(t - 1) * D_EPI_THREADS];
        if (!(prev < s)) break;  // strict: an equal earlier (smaller index) entry stays ahead
        ls[t * D_EPI_THREADS] = prev;
        li[t * D_EPI_THREADS] = li[(t - 1) * D_EPI_THREADS];
        --t;
    }
    ls[t * D_EPI_THREADS] = s;
    li[t * D_EPI_THREADS] = idx;
    return ls[(k - 1) * D_EPI_THREADS];
}

// dense2.cu (CTA-pair kernel) entry points used by the C ABI in dense.cu
bool dense2_supported(int64_t nq, int64_t np, int32_t d_pad, int32_t prec, int32_t k);
size_t dense2_workspace_bytes(int64_t nq, int64_t np, int32_t d_pad, int32_t k, bool x3);
// q_lo / p_lo non-null selects the split-precision (BF16X3) contraction
int dense2_topk(const void* q_hi, const void* q_lo, int64_t nq, const void* p_hi, const void* p_lo, int64_t np, int32_t d_pad,
                const float* q_time, const float* p_time, float lambda, int32_t mode, int32_t k, int64_t pool_base,
                void* workspace, float** part_score_out, int32_t** part_idx_out, int32_t* n_lists_out, cudaStream_t st);

}  // namespace r4d
