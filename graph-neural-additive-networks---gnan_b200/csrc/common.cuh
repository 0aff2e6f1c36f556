// Shared helpers for the gnan_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gnan_b200.h"

void gnan_set_error(const char *fmt, ...);

#define GNAN_REQUIRE(cond, ...)          \
    do {                                 \
        if (!(cond)) {                   \
            gnan_set_error(__VA_ARGS__); \
            return GNAN_ERR_INVALID;     \
        }                                \
    } while (0)

#define GNAN_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            gnan_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return GNAN_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

void gnan_count_launch();
#define GNAN_LAUNCH_OK()               \
    do {                               \
        gnan_count_launch();           \
        GNAN_CUDA(cudaGetLastError()); \
    } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// B200: 148 SMs. Queried once; used to size persistent grids.
int gnan_sm_count();
int gnan_reduce_chunks(const float *part, int nchunk, size_t n, size_t stride, float *out, cudaStream_t st);
struct GnanReduceSegs {
    const float *src[6];
    float *dst[6];
    size_t n[6], total;
    int count;
    void add(const float *s, float *d, size_t len)
    {
        if (!d || len == 0) return;
        src[count] = s; dst[count] = d; n[count] = len; total += len; ++count;
    }
};
int gnan_reduce_chunks_multi(const GnanReduceSegs &sg, int nchunk, size_t stride, cudaStream_t st);

// Counter-based dropout mask: splitmix64 finaliser over (seed, flat element index). keep-probability 1-p.
__device__ __forceinline__ uint32_t gnan_hash32(uint64_t seed, uint64_t idx)
{
    uint64_t z = seed + idx * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 32);
}

// element key = (((layer * G + g) * R + row) * H + unit); returns the multiplier (0 or 1/(1-p)).
__device__ __forceinline__ float gnan_dropout_mul(uint64_t seed, uint64_t key, uint32_t thresh, float scale)
{
    return gnan_hash32(seed, key) >= thresh ? scale : 0.0f;
}

__host__ __device__ inline uint32_t gnan_dropout_thresh(float p)
{
    double t = (double)p * 4294967296.0;
    if (t < 0) t = 0;
    if (t > 4294967295.0) t = 4294967295.0;
    return (uint32_t)t;
}
