// probe_host_store.cu — what rate can a kernel store results straight into pinned HOST memory at, by store shape?
// (HostTopK's direct mode: the top-K kernel writes its [Q,K] lists over PCIe as posted writes.)
// Patterns, all writing the same number of bytes:
//   0  warp-wide 128-byte stores, 128-byte aligned (one full line per instruction)
//   1  warp-wide 128-byte stores, 64 bytes off the line grid (every instruction spans two lines)
//   2  320-byte pieces per warp (8 rows x 10 words), three planes, line-oblivious loop (the kernel's shape before)
//   3  320-byte pieces per warp, three planes, loop aligned to the line grid
//   4  warp-wide 512-byte stores (st.v4 per lane), aligned
// and the copy engine (cudaMemcpyAsync D2H) for the same bytes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_host_store tools/probe_host_store.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void __launch_bounds__(256) store_kernel(uint32_t* __restrict__ dst, int64_t n_words, int pattern, int delay) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    if (pattern == 0 || pattern == 1) {
        const int64_t shift = pattern == 1 ? 16 : 0;
        for (int64_t w = warp * 32; w + 32 + shift <= n_words; w += n_warps * 32) {
            if (delay) __nanosleep(delay);
            dst[w + shift + lane] = (uint32_t)w;
        }
    } else if (pattern == 2 || pattern == 3) {
        const int64_t plane = n_words / 3;
        const int64_t n_pieces = plane / 80;
        for (int64_t pc = warp; pc < n_pieces; pc += n_warps) {
            if (delay) __nanosleep(delay);
            const int64_t at = pc * 80;
            const int mis = pattern == 3 ? (int)(at & 31) : 0;
            for (int pl = 0; pl < 3; ++pl)
                for (int i = lane - mis; i < 80; i += 32)
                    if (i >= 0) dst[pl * plane + at + i] = (uint32_t)pc;
        }
    } else {
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        const int64_t n4 = n_words / 4;
        for (int64_t w = warp * 32; w + 32 <= n4; w += n_warps * 32) {
            if (delay) __nanosleep(delay);
            d4[w + lane] = make_uint4((uint32_t)w, 1u, 2u, 3u);
        }
    }
}

int main() {
    const size_t sizes[2] = {12000000, 8400000};
    uint32_t *host, *dev;
    cudaHostAlloc(&host, 16 << 20, cudaHostAllocDefault);
    cudaMalloc(&dev, 16 << 20);
    cudaMemset(dev, 1, 16 << 20);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (size_t bytes : sizes) {
        const int64_t n_words = (int64_t)(bytes / 4 / 960 * 960);
        for (int pattern = 0; pattern < 5; ++pattern)
            for (int grid_mul : {1, 4}) {
                float best = 1e9f;
                for (int rep = 0; rep < 6; ++rep) {
                    cudaEventRecord(e0);
                    store_kernel<<<sms * grid_mul, 256>>>(host, n_words, pattern, 0);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    if (rep && ms < best) best = ms;
                }
                printf("{\"bytes\": %lld, \"pattern\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"GBps\": %.1f}\n",
                       (long long)n_words * 4, pattern, grid_mul, best, n_words * 4 / best * 1e-6);
            }
        // stores spread over ~0.25 ms (the scoring kernel's run time): does the transfer hide behind it?
        for (int pattern : {0, 3}) {
            float best = 1e9f;
            for (int rep = 0; rep < 6; ++rep) {
                cudaEventRecord(e0);
                store_kernel<<<sms * 4, 256>>>(host, n_words, pattern, 8000);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep && ms < best) best = ms;
            }
            printf("{\"bytes\": %lld, \"pattern\": %d, \"paced\": true, \"ms\": %.4f}\n", (long long)n_words * 4, pattern, best);
        }
        float best = 1e9f;
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(e0);
            cudaMemcpyAsync(host, dev, n_words * 4, cudaMemcpyDeviceToHost, 0);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        printf("{\"bytes\": %lld, \"pattern\": \"copy engine\", \"ms\": %.4f, \"GBps\": %.1f}\n", (long long)n_words * 4, best,
               n_words * 4 / best * 1e-6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
