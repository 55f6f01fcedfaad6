// fp64 add on the SMs as the tree kernel's softmax uses it: latency of a dependent DADD chain in one warp,
// and what W warps per SM (each on its own chain) pay per DADD -- with all 32 lanes active and with one lane active --
// i.e. whether the strictly sequential 833-term sums are bound by the pipe's width or by its latency.
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void k(double *out, const double *in, int iters, int active_lanes, long long *cycles)
{
    double a[CHAINS];
    for (int i = 0; i < CHAINS; ++i) a[i] = in[i];
    const double b = in[8 + (threadIdx.x & 1)];
    __syncthreads();
    const long long t0 = clock64();
    if ((int)(threadIdx.x & 31) < active_lanes) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 32; ++r)
#pragma unroll
                for (int i = 0; i < CHAINS; ++i) a[i] = __dadd_rn(a[i], b);
        }
    }
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < CHAINS; ++i) s += a[i];
    if (s == 0.123) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}
template <int CHAINS>
void run(int warps, int lanes)
{
    double *out, *in; long long *cyc;
    cudaMalloc(&out, 8); cudaMalloc(&in, 128); cudaMalloc(&cyc, 8);
    double h[16]; for (int i = 0; i < 16; ++i) h[i] = 1.0 + i * 1e-9;
    cudaMemcpy(in, h, 128, cudaMemcpyHostToDevice);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 2000;
    k<CHAINS><<<sms, warps * 32>>>(out, in, 10, lanes, cyc);
    k<CHAINS><<<sms, warps * 32>>>(out, in, iters, lanes, cyc);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)c / ((double)iters * 32 * CHAINS);
    printf("warps/SM %2d  chains/warp %d  active lanes %2d : %.2f clk per DADD per warp, %.2f warp-DADD/clk/SM\n", warps, CHAINS, lanes, per,
           warps / per);
}
int main()
{
    run<1>(1, 32); run<1>(1, 1);
    run<1>(4, 32); run<1>(8, 32); run<1>(16, 32); run<1>(16, 1); run<1>(16, 4); run<1>(32, 32);
    run<4>(4, 32); run<4>(16, 32); run<4>(16, 1);
    return 0;
}
