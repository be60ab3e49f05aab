// r4d_host.cu — error plumbing, device probe, TMA descriptor factory.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>

#include "r4d_common.cuh"

namespace r4d {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return R4D_E_CUDA;
}

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches_so_far() { return g_launches.load(std::memory_order_relaxed); }

Options& options() {
    static Options o = [] {
        Options v;
        if (const char* e = getenv("R4D_JACCARD_NOSKIP")) v.jaccard_skip_zero = atoi(e) ? 0 : 1;
        if (const char* e = getenv("R4D_JACCARD_NOSPARSE")) v.jaccard_sparse_q = atoi(e) ? 0 : 1;
        if (const char* e = getenv("R4D_JACCARD_WARPS")) v.jaccard_warps = atoi(e) == 8 ? 8 : 16;
        if (const char* e = getenv("R4D_DENSE_V1")) v.dense_pair_kernel = atoi(e) ? 0 : 1;
        if (const char* e = getenv("R4D_DENSE2_QRES")) v.dense_pair_qres = atoi(e);
        if (const char* e = getenv("R4D_JACCARD_STRIPES")) v.jaccard_stripes = atoi(e);
        if (const char* e = getenv("R4D_DENSE_STRIPES")) v.dense_stripes = atoi(e);
        if (const char* e = getenv("R4D_DENSE_WALKER_WINDOW")) v.dense_walker_window = atoi(e);
        if (const char* e = getenv("R4D_POSTINGS_CHUNK")) v.postings_chunk = atoi(e);
        if (const char* e = getenv("R4D_POSTINGS_KERNEL")) v.postings_kernel = atoi(e);
        return v;
    }();
    return o;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode_fn() {
    static encode_tiled_fn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<encode_tiled_fn>(p);
    });
    return fn;
}

int make_tmap_2d(CUtensorMap* map, CUtensorMapDataType dtype, int elem_bytes, const void* base, uint64_t dim0,
                 uint64_t dim1, uint64_t row_stride_bytes, uint32_t box0, uint32_t box1, CUtensorMapSwizzle swz) {
    encode_tiled_fn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled not available from the CUDA driver");
        return R4D_E_CUDA;
    }
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (row_stride_bytes & 15) != 0) {
        set_error("TMA operand must be 16-byte aligned (base %p, row stride %llu B)", base,
                  (unsigned long long)row_stride_bytes);
        return R4D_E_ARG;
    }
    (void)elem_bytes;
    cuuint64_t dims[2] = {dim0, dim1};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, dtype, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (dims %llu x %llu, stride %llu, box %u x %u)", (int)r,
                  (unsigned long long)dim0, (unsigned long long)dim1, (unsigned long long)row_stride_bytes, box0, box1);
        return R4D_E_CUDA;
    }
    return R4D_OK;
}

// ---------------------------------------------------------------- kernel timing (measurement aid)
struct ProfState {
    static constexpr int CAP = 4096;
    cudaEvent_t beg[CAP], end[CAP];
    int created = 0, used = 0;
    bool open = false;
};
static ProfState g_prof[PROF_KERNELS];
static std::mutex g_prof_mu;

void prof_begin(ProfKernel k, cudaStream_t st) {
    if (!options().kernel_timing) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfState& p = g_prof[k];
    if (p.used >= ProfState::CAP) return;
    if (p.used >= p.created) {
        if (cudaEventCreate(&p.beg[p.created]) != cudaSuccess || cudaEventCreate(&p.end[p.created]) != cudaSuccess) return;
        ++p.created;
    }
    p.open = cudaEventRecord(p.beg[p.used], st) == cudaSuccess;
}

void prof_end(ProfKernel k, cudaStream_t st) {
    if (!options().kernel_timing) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfState& p = g_prof[k];
    if (!p.open) return;
    p.open = false;
    if (cudaEventRecord(p.end[p.used], st) == cudaSuccess) ++p.used;
}

}  // namespace r4d

extern "C" {

int r4d_version(void) { return 100; }

int64_t r4d_kernel_launches(void) { return (int64_t)r4d::launches_so_far(); }

int r4d_profile_read(const char* kernel, double* total_ms, int64_t* launches) {
    using namespace r4d;
    R4D_REQUIRE(kernel && total_ms && launches, "r4d_profile_read: null argument");
    int k = -1;
    if (!strcmp(kernel, "jaccard_qindex")) k = PROF_JACCARD_QINDEX;
    else if (!strcmp(kernel, "dense_pair")) k = PROF_DENSE_PAIR;
    else if (!strcmp(kernel, "jaccard_postings")) k = PROF_JACCARD_POSTINGS;
    R4D_REQUIRE(k >= 0, "r4d_profile_read: unknown kernel '%s'", kernel);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfState& p = g_prof[k];
    double sum = 0.0;
    for (int i = 0; i < p.used; ++i) {
        R4D_CUDA(cudaEventSynchronize(p.end[i]));
        float ms = 0.f;
        R4D_CUDA(cudaEventElapsedTime(&ms, p.beg[i], p.end[i]));
        sum += ms;
    }
    *total_ms = sum;
    *launches = p.used;
    p.used = 0;
    return R4D_OK;
}

const char* r4d_last_error(void) { return r4d::g_err; }

int r4d_set_option(const char* key, int value) {
    if (!key) return R4D_E_ARG;
    r4d::Options& o = r4d::options();
    int* slot = nullptr;
    if (!strcmp(key, "jaccard_skip_zero")) slot = &o.jaccard_skip_zero;
    else if (!strcmp(key, "jaccard_sparse_q")) slot = &o.jaccard_sparse_q;
    else if (!strcmp(key, "jaccard_debug")) slot = &o.jaccard_debug;
    else if (!strcmp(key, "jaccard_warps")) slot = &o.jaccard_warps;
    else if (!strcmp(key, "dense_pair_kernel")) slot = &o.dense_pair_kernel;
    else if (!strcmp(key, "dense_pair_qres")) slot = &o.dense_pair_qres;
    else if (!strcmp(key, "stripe_interleave")) slot = &o.stripe_interleave;
    else if (!strcmp(key, "jaccard_stripes")) slot = &o.jaccard_stripes;
    else if (!strcmp(key, "dense_stripes")) slot = &o.dense_stripes;
    else if (!strcmp(key, "kernel_timing")) slot = &o.kernel_timing;
    else if (!strcmp(key, "postings_log_t")) slot = &o.postings_log_t;
    else if (!strcmp(key, "dense_walker_window")) slot = &o.dense_walker_window;
    else if (!strcmp(key, "dense_x3_combined")) slot = &o.dense_x3_combined;
    else if (!strcmp(key, "postings_chunk")) slot = &o.postings_chunk;
    else if (!strcmp(key, "postings_kernel")) slot = &o.postings_kernel;
    else if (!strcmp(key, "postings_best")) slot = &o.postings_best;
    else if (!strcmp(key, "postings_relay")) slot = &o.postings_relay;
    if (!slot || (slot == &o.jaccard_warps && value != 8 && value != 16)) {
        r4d::set_error("r4d_set_option: unknown key or bad value (%s = %d)", key, value);
        return R4D_E_ARG;
    }
    const int prev = *slot;
    *slot = value;
    return prev;
}

int r4d_device_ok(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        r4d::set_error("no CUDA device visible");
        return 0;
    }
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    if (major != 10) {
        r4d::set_error("device %d has compute capability %d.x; libr4d is built for sm_100a only", dev, major);
        return 0;
    }
    return 1;
}

}  // extern "C"
