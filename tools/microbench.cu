// microbench.cu — per-SM issue rates of the integer ops the Jaccard kernel is made of (POPC, LOP3, IADD3) and of
// the combined AND+POPC+ADD body, measured with clock64 on one 1024-thread CTA per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int ITERS = 4096;
constexpr int ILP = 8;

template <int KIND>
__global__ void __launch_bounds__(1024) k(uint32_t seed, uint32_t* out, long long* cycles) {
    uint32_t a[ILP], acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        a[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u;
        acc[i] = i;
    }
    uint32_t b = seed ^ 0x5bd1e995u;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (KIND == 0) {        // POPC chain-free: acc = popc(acc ^ const)   (1 POPC + 1 LOP3 dependent pair)
                asm volatile("popc.b32 %0, %1;" : "=r"(acc[i]) : "r"(acc[i] ^ a[i]));
            } else if (KIND == 1) { // LOP3 only
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(acc[i]) : "r"(acc[i]), "r"(a[i]), "r"(b));
            } else if (KIND == 2) { // IADD3
                asm volatile("add.u32 %0, %1, %2;" : "=r"(acc[i]) : "r"(acc[i]), "r"(a[i]));
            } else {                // the kernel's body: acc += popc(a & b')
                uint32_t t;
                asm volatile("and.b32 %0, %1, %2;" : "=r"(t) : "r"(a[i]), "r"(b));
                uint32_t p;
                asm volatile("popc.b32 %0, %1;" : "=r"(p) : "r"(t));
                acc[i] += p;
                b = b * 3 + 1;
            }
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND>
double run(int sms, const char* name) {
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(uint32_t) * sms * 1024);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    k<KIND><<<sms, 1024>>>(12345u, out, cyc);
    k<KIND><<<sms, 1024>>>(12345u, out, cyc);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < sms; ++i) mean += (double)h[i];
    mean /= sms;
    const double ops = 1024.0 * ITERS * ILP;
    printf("\"%s_per_clk_per_sm\": %.2f, ", name, ops / mean);
    cudaFree(out);
    cudaFree(cyc);
    return ops / mean;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("{\"sms\": %d, ", sms);
    run<0>(sms, "popc_xor");
    run<1>(sms, "lop3");
    run<2>(sms, "iadd");
    run<3>(sms, "and_popc_add");
    printf("\"note\": \"thread-ops per SM clock, 1024 threads/SM, ILP 8\"}\n");
    return 0;
}
