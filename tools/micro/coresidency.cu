// Can blocks of a small kernel be scheduled on SMs that are full of a persistent, shared-memory-heavy kernel?
// A: 2 CTAs/SM x 107 KB dynamic smem, 192 threads, spins for ~1 ms.  B: 128 threads, 4 KB static smem, trivial.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(192, 2) big(long long cycles, int *sink)
{
    extern __shared__ char smem[];
    smem[threadIdx.x] = 1;
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
    if (smem[threadIdx.x] == 2) *sink = 1;
}
__global__ void __launch_bounds__(128, 4) small_k(long long *stamp, int *sink, int regs_dummy)
{
    __shared__ char s[4096];
    s[threadIdx.x] = 1;
    if (threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        stamp[blockIdx.x] = (long long)t;
    }
    if (s[threadIdx.x] == 2) *sink = regs_dummy;
}
int main(int argc, char **argv)
{
    const int big_smem = argc > 1 ? atoi(argv[1]) : 107 * 1024;
    const int carve_small = argc > 2 ? atoi(argv[2]) : -1;
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(big, cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem);
    if (carve_small >= 0) cudaFuncSetAttribute(small_k, cudaFuncAttributePreferredSharedMemoryCarveout, carve_small);
    cudaStream_t s1, s2; cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    long long *stamp; int *sink; cudaMalloc(&stamp, 8 * 1024); cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    for (int rep = 0; rep < 3; ++rep) {
        cudaDeviceSynchronize();
        cudaEventRecord(e0, s1);
        big<<<2 * sms, 192, big_smem, s1>>>(2000000LL, sink);      // ~1 ms at ~2 GHz
        cudaEventRecord(e1, s1);
        cudaEventRecord(e2, s2);
        small_k<<<256, 128, 0, s2>>>(stamp, sink, rep);
        cudaEvent_t e3; cudaEventCreate(&e3); cudaEventRecord(e3, s2);
        cudaDeviceSynchronize();
        float big_ms, small_ms; cudaEventElapsedTime(&big_ms, e0, e1); cudaEventElapsedTime(&small_ms, e2, e3);
        printf("big_smem=%d carve_small=%d: big kernel %.3f ms, small kernel (launched right after, other stream) finished after %.3f ms -> %s\n",
               big_smem, carve_small, big_ms, small_ms, small_ms < 0.5f * big_ms ? "CO-RESIDENT" : "waited for the big kernel");
    }
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
