// Error reporting, version and device queries of the C ABI (include/gnan_b200.h).
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

static thread_local char g_err[512] = "";

void gnan_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int gnan_sm_count()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
            n = v;
        else
            n = 148;
    }
    return n;
}

static unsigned long long g_launches = 0;
void gnan_count_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }
extern "C" uint64_t gnan_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

extern "C" const char *gnan_last_error(void) { return g_err; }
extern "C" int gnan_version(void) { return GNAN_B200_VERSION; }

// out[i] = sum_c part[c*stride + i] in fixed order (deterministic); shared by the partial-sum paths
__global__ void gnan_reduce_chunks_kernel(const float *__restrict__ part, int nchunk, size_t n, size_t stride, float *__restrict__ out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t step = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += step) {
        float s = 0.f;
        for (int c = 0; c < nchunk; ++c) s += part[(size_t)c * stride + i];
        out[i] = s;
    }
}

int gnan_reduce_chunks(const float *part, int nchunk, size_t n, size_t stride, float *out, cudaStream_t st)
{
    if (n == 0) return GNAN_OK;
    size_t blocks = (n + 255) / 256;
    if (blocks > (size_t)gnan_sm_count() * 8) blocks = (size_t)gnan_sm_count() * 8;
    gnan_reduce_chunks_kernel<<<(unsigned)blocks, 256, 0, st>>>(part, nchunk, n, stride, out);
    GNAN_LAUNCH_OK();
    return GNAN_OK;
}

// the same for up to 6 (src, dst, n) segments that share chunk count and stride (the gradient arrays of one MLP backward): ONE launch
__global__ void gnan_reduce_chunks_multi_kernel(GnanReduceSegs sg, int nchunk, size_t stride)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t step = (size_t)gridDim.x * blockDim.x;
    for (; i < sg.total; i += step) {
        int k = 0;
        size_t off = i;
        while (k + 1 < sg.count && off >= sg.n[k]) { off -= sg.n[k]; ++k; }
        const float *src = sg.src[k] + off;
        float s = 0.f;
        for (int c = 0; c < nchunk; ++c) s += src[(size_t)c * stride];
        sg.dst[k][off] = s;
    }
}

int gnan_reduce_chunks_multi(const GnanReduceSegs &sg, int nchunk, size_t stride, cudaStream_t st)
{
    if (sg.count == 0 || sg.total == 0) return GNAN_OK;
    size_t blocks = (sg.total + 255) / 256;
    if (blocks > (size_t)gnan_sm_count() * 8) blocks = (size_t)gnan_sm_count() * 8;
    gnan_reduce_chunks_multi_kernel<<<(unsigned)blocks, 256, 0, st>>>(sg, nchunk, stride);
    GNAN_LAUNCH_OK();
    return GNAN_OK;
}
