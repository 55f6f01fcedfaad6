// az_common.h -- shared declarations of libataxxzero (internal).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/ataxxzero.h"

#define AZ_HD __host__ __device__ __forceinline__
#define AZ_D __device__ __forceinline__

// error plumbing ----------------------------------------------------------
int az_fail(int code, const char *fmt, ...);            // sets thread-local message, returns code
#define AZ_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return az_fail(AZ_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                           __FILE__, __LINE__);                                               \
    } while (0)
#define AZ_REQUIRE(cond, code, ...)                     \
    do {                                                \
        if (!(cond)) return az_fail(code, __VA_ARGS__); \
    } while (0)

// device scratch buffer that only ever grows
struct AzBuffer {
    void *ptr = nullptr;
    size_t bytes = 0;
    int reserve(size_t need);
    void release();
    template <typename T> T *as() const { return reinterpret_cast<T *>(ptr); }
};

struct AzNet;        // az_net.cu
struct AzPerftState; // az_perft.cu

struct az_context {
    int device = 0;
    int sm_count = 0;
    uint64_t seed = 0;
    cudaStream_t stream = nullptr;
    AzBuffer scratch[8];        // general-purpose staging for the *_batch entry points
    AzPerftState *perft = nullptr;
    AzNet *net = nullptr;
    unsigned long long launches = 0;   // kernels launched by this library on this context
};
