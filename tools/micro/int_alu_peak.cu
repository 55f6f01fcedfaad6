// Integer-ALU peak of the SMs (SURVEY 8d asks for one next to the perft numbers): dependent-free LOP3 / IADD3 / SHF
// streams, 8 independent chains per thread, enough warps to fill every scheduler.
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND>
__global__ void __launch_bounds__(256) k(unsigned *out, int iters)
{
    unsigned a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 2654435761u + i;
    const unsigned b = blockIdx.x | 1u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (KIND == 0) a[i] = (a[i] & b) ^ (a[(i + 1) & 7] | 0x5bd1e995u);            // LOP3
                else if (KIND == 1) a[i] = a[i] + a[(i + 1) & 7] + b;                          // IADD3
                else a[i] = __funnelshift_l(a[i], a[(i + 1) & 7], 7) ;                         // SHF
            }
    }
    unsigned s = 0;
    for (int i = 0; i < 8; ++i) s ^= a[i];
    if (s == 0x12345678u) out[0] = s;
}
template <int KIND>
void run(const char *name, int sms)
{
    unsigned *out; cudaMalloc(&out, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000, blocks = sms * 8;
    k<KIND><<<blocks, 256>>>(out, 10);
    cudaEventRecord(e0);
    k<KIND><<<blocks, 256>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * 256 * iters * 16 * 8;
    printf("%s: %.2f Tops/s (thread-level int32 ops), %.1f ops/clk/SM at 1.965 GHz\n", name, ops / ms / 1e9, ops / (ms * 1e-3) / sms / 1.965e9);
}
int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<0>("LOP3 ", sms); run<1>("IADD3", sms); run<2>("SHF  ", sms);
    return 0;
}
