// r4d_common.cuh — shared host/device helpers for libr4d.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/r4d.h"

namespace r4d {

// ---------------------------------------------------------------- host side
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
int num_sms();  // SM count of the current device (cached)
void note_launch();  // counts the kernels this library has enqueued (r4d_kernel_launches)

struct Options {
    int jaccard_skip_zero = 1;
    int jaccard_sparse_q = 1;   // fused top-K: sparse query-side span lists for sparse query tiles (jaccard_sparse.cu)
    int jaccard_debug = 0;      // query-index kernel experiments (stage bypass); results are wrong unless 0
    int jaccard_warps = 16;
    int dense_pair_kernel = 1;
    int dense_pair_qres = -1;
    int jaccard_stripes = 0;    // 0: automatic; > 0: force the number of pool stripes (experiments)
    int dense_stripes = 0;
    int stripe_interleave = 0;  // dense pair kernel: 1 = stripe s owns pool tiles s, s+S, ...; 0 = contiguous stripes
    int kernel_timing = 0;      // bracket the dominant kernels with CUDA events (r4d_profile_read)
    int dense_x3_combined = 1;      // dense pair kernel, split precision: hi + lo planes of a k-block in ONE 64 KB stage
    int dense_walker_window = 4;    // dense pair kernel: tiles a walker may lead the slowest walker of its stripe (0 = off)
    int postings_best = 1;      // postings path: single-id queries read their top-K from the per-id best lists
    int postings_kernel = 0;    // postings path, label-like sets, first stage: 0 = head kernel, 1 = hash-table, 2 = register kernel
    int postings_chunk = 0;     // postings path: queries per grab of the work counter (0 = automatic, <= 8)
    int postings_relay = 1;     // postings path: packed lists bound for pinned host memory leave in whole 64-query blocks
    int postings_log_t = 0;     // postings path: 0 = automatic, 9 / 10 = force 512 / 1 024-slot hash tables
};
Options& options();  // process-wide knobs (r4d_set_option); environment variables R4D_* give the initial values

#define R4D_CUDA(expr)                                                        \
    do {                                                                      \
        cudaError_t e__ = (expr);                                             \
        if (e__ != cudaSuccess) return r4d::cuda_fail(e__, #expr, __FILE__, __LINE__); \
    } while (0)

#define R4D_REQUIRE(cond, ...)              \
    do {                                    \
        if (!(cond)) {                      \
            r4d::set_error(__VA_ARGS__);    \
            return R4D_E_ARG;               \
        }                                   \
    } while (0)

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda at link time).
// 2-D row-major tensor: dim0 = contiguous elements per row, dim1 = rows.
int make_tmap_2d(CUtensorMap* map, CUtensorMapDataType dtype, int elem_bytes, const void* base, uint64_t dim0,
                 uint64_t dim1, uint64_t row_stride_bytes, uint32_t box0, uint32_t box1, CUtensorMapSwizzle swz);

// Peer scatter target of the fused exchange: every rank's merge kernel stores its [nq][k] result planes into slot
// `rank` of EVERY peer's gather buffer (layout [n_planes][world][nq][k], 4-byte elements) through NVLink peer pointers.
struct PeerOut {
    void* base[R4D_MAX_PEERS];
    int32_t world;  // 0: no scatter (plain local output)
    int32_t rank;
};

static inline cudaStream_t as_stream(r4d_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: one flag per (call site, device),
// so a process that drives several GPUs opts every one of them in before its first launch there.
struct SmemOptIn {
    bool done[64] = {};
};
template <class K>
static inline int ensure_dyn_smem(K kern, size_t bytes, SmemOptIn& once) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = -1;
    if (dev >= 0 && dev < 64 && once.done[dev]) return R4D_OK;
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)", __FILE__, __LINE__);
    if (dev >= 0 && dev < 64) once.done[dev] = true;
    return R4D_OK;
}

// Measurement aid: with the option "kernel_timing" set, a dominant kernel's launch is bracketed by CUDA events on its
// own stream (prof_begin / prof_end around the <<<>>>); r4d_profile_read sums the elapsed times.
enum ProfKernel { PROF_JACCARD_QINDEX = 0, PROF_DENSE_PAIR = 1, PROF_JACCARD_POSTINGS = 2, PROF_KERNELS = 3 };
void prof_begin(ProfKernel k, cudaStream_t st);
void prof_end(ProfKernel k, cudaStream_t st);

// ---------------------------------------------------------------- device side
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded spin: a pipeline bug must surface as a launch failure (trap), never as a hung GPU.
#ifndef R4D_SPIN_TIMEOUT_CYCLES
#define R4D_SPIN_TIMEOUT_CYCLES (40ll * 1000 * 1000 * 1000)  // ~20 s at 2 GHz
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    uint32_t polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++polls & 0x3ffu) == 0 && clock64() - t0 > R4D_SPIN_TIMEOUT_CYCLES) __trap();
    }
}
// TMA 2-D tile load: global (tensor map) -> shared, completion on an mbarrier (SASS: UTMALDG).
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* map, uint64_t* bar, int32_t c0,
                                            int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
            "r"(smem_u32(dst_smem)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------- warp-level sorted top-K list
// One list per warp, one entry per lane (lane t holds the t-th best, t < k <= 32).  `Entry` provides
//   static bool better(const Entry& a, const Entry& b)   -- strict total order "a ranks before b"
//   Entry shfl(int src_lane) const                         -- warp broadcast
//   Entry shfl_up1() const                                 -- value of lane-1
//   static Entry worst()
template <class Entry>
struct WarpTopK {
    Entry mine;  // lane t: t-th best entry (lanes >= k hold worst())
    Entry kth;   // warp-uniform copy of lane k-1
    int k;

    __device__ __forceinline__ void init(int k_) {
        k = k_;
        mine = Entry::worst();
        kth = Entry::worst();
    }
    __device__ __forceinline__ void refresh_kth() { kth = mine.shfl(k - 1); }
    // warp-uniform candidate `c` (all lanes pass the same value); all 32 lanes must call.
    __device__ __forceinline__ void insert(const Entry& c) {
        if (!Entry::better(c, kth)) return;  // warp-uniform branch
        const uint32_t lane = lane_id();
        // entries that rank before c keep their slot; the rest shift down by one
        const bool keep = (lane < (uint32_t)k) && Entry::better(mine, c);
        const uint32_t pos = __popc(__ballot_sync(0xffffffffu, keep));
        Entry up = mine.shfl_up1();
        if (lane == pos)
            mine = c;
        else if (lane > pos && lane < (uint32_t)k)
            mine = up;
        refresh_kth();
    }
};

// Jaccard candidate: score = inter/uni compared exactly by cross-multiplication (no rounding), ties by
// ascending pool index.  uni == 0 only when both sets are empty (inter == 0): it then compares as score 0.
struct JEntry {
    uint32_t inter, uni;
    int32_t idx;
    __device__ __forceinline__ static JEntry worst() { return JEntry{0u, 1u, R4D_IDX_NONE}; }
    __device__ __forceinline__ static bool better(const JEntry& a, const JEntry& b) {
        const uint64_t l = (uint64_t)a.inter * (uint64_t)b.uni;
        const uint64_t r = (uint64_t)b.inter * (uint64_t)a.uni;
        return (l > r) || (l == r && a.idx < b.idx);
    }
    __device__ __forceinline__ JEntry shfl(int src) const {
        return JEntry{__shfl_sync(0xffffffffu, inter, src), __shfl_sync(0xffffffffu, uni, src),
                      __shfl_sync(0xffffffffu, idx, src)};
    }
    __device__ __forceinline__ JEntry shfl_up1() const {
        return JEntry{__shfl_up_sync(0xffffffffu, inter, 1), __shfl_up_sync(0xffffffffu, uni, 1),
                      __shfl_up_sync(0xffffffffu, idx, 1)};
    }
    __device__ __forceinline__ JEntry shfl_down1() const {
        return JEntry{__shfl_down_sync(0xffffffffu, inter, 1), __shfl_down_sync(0xffffffffu, uni, 1),
                      __shfl_down_sync(0xffffffffu, idx, 1)};
    }
};

// Float candidate (dense scorer / explicit matrices).  NaN ranks after every number (numpy puts NaN last).
template <class T>
struct FEntry {
    T s;
    int32_t idx;
    __device__ __forceinline__ static FEntry worst() {
        // a NaN with the largest index: nothing ranks after it
        T nan_v = (sizeof(T) == 8) ? (T)__longlong_as_double(0x7ff8000000000000LL) : (T)__int_as_float(0x7fc00000);
        return FEntry{nan_v, R4D_IDX_NONE};
    }
    __device__ __forceinline__ static bool better(const FEntry& a, const FEntry& b) {
        const bool an = (a.s != a.s), bn = (b.s != b.s);
        if (an || bn) return (an == bn) ? (a.idx < b.idx) : bn;
        return (a.s > b.s) || (a.s == b.s && a.idx < b.idx);
    }
    __device__ __forceinline__ FEntry shfl(int src) const {
        return FEntry{__shfl_sync(0xffffffffu, s, src), __shfl_sync(0xffffffffu, idx, src)};
    }
    __device__ __forceinline__ FEntry shfl_up1() const {
        return FEntry{__shfl_up_sync(0xffffffffu, s, 1), __shfl_up_sync(0xffffffffu, idx, 1)};
    }
};

#endif  // __CUDACC__
}  // namespace r4d
