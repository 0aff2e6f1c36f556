// Glue kernels of the compressed-feature ("entries") path: entry values <-> node rows. Deterministic (fixed summation
// order, no atomics). See gnan_b200.h (gnan_mlp_entries_*) and gnan_b200/sparse.py.
//
// Layout: group g owns entries [grp_ptr[g], grp_ptr[g+1]); its FIRST entry is the feature's baseline value (shared by
// every row that is not listed), the others are exceptions (row, value). S[r,:] = sum_g f_g(x[r,g]) becomes
//   S[r,:] = sum_g Y[base_g,:] + sum_{e in exceptions of row r} (Y[e,:] - Y[base_{g(e)},:]).
#include <algorithm>

#include "common.cuh"

namespace {

// out[c] = sum_i src[idx(i), c] for i < n, fixed-order tree; one block per channel
template <bool BASELINES>
__global__ void colsum_kernel(const float *__restrict__ src, const int64_t *__restrict__ grp_ptr, int64_t n, int C,
                              float *__restrict__ out)
{
    __shared__ float red[512];
    const int c = blockIdx.x;
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += src[(BASELINES ? grp_ptr[i] : i) * C + c];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = blockDim.x / 2; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[c] = red[0];
}

// two-stage deterministic column sum of a tall [n,C] matrix: partial[blk][c] over row slabs, then the fixed-order final sum
constexpr int COLSUM_SLABS = 64;
__global__ void colsum_partial_kernel(const float *__restrict__ src, int64_t n, int C, float *__restrict__ partial)
{
    __shared__ float red[256];
    const int c = blockIdx.x, blk = blockIdx.y;
    const int64_t per = (n + gridDim.y - 1) / gridDim.y, r0 = blk * per, r1 = min(n, r0 + per);
    float s = 0.f;
    for (int64_t i = r0 + threadIdx.x; i < r1; i += blockDim.x) s += src[i * C + c];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = blockDim.x / 2; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blk * C + c] = red[0];
}

__global__ void colsum_final_kernel(const float *__restrict__ partial, int nblk, int C, float *__restrict__ out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    for (int b = 0; b < nblk; ++b) s += partial[b * C + c];
    out[c] = s;
}

__global__ void entries_to_rows_kernel(const float *__restrict__ Y, int64_t N, int C, const int64_t *__restrict__ grp_ptr,
                                       const int64_t *__restrict__ csr_ptr, const int64_t *__restrict__ csr_eid,
                                       const int32_t *__restrict__ ent_grp, const float *__restrict__ S0, float *__restrict__ S)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * C) return;
    const int64_t r = t / C;
    const int c = (int)(t % C);
    float s = S0[c];
    for (int64_t i = csr_ptr[r]; i < csr_ptr[r + 1]; ++i) {
        const int64_t e = csr_eid[i];
        s += Y[e * C + c] - Y[grp_ptr[ent_grp[e]] * C + c];
    }
    S[t] = s;
}

// dY of the exceptions: a gather of the owning row's dS
__global__ void rows_to_exceptions_kernel(const float *__restrict__ dS, int64_t E, int C, const int64_t *__restrict__ ent_row,
                                          float *__restrict__ dY)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= E * C) return;
    const int64_t e = t / C;
    const int64_t r = ent_row[e];
    if (r >= 0) dY[t] = dS[r * C + (t % C)];
}

// dY of the baseline of group g = (sum of dS over all rows) - (sum over the rows listed as exceptions of g); one block per
// (g, c), fixed-order strided partial sums + tree (deterministic)
__global__ void rows_to_baselines_kernel(const float *__restrict__ dS, int G, int C, const int64_t *__restrict__ grp_ptr,
                                         const int64_t *__restrict__ ent_row, const float *__restrict__ dStot,
                                         float *__restrict__ dY)
{
    __shared__ float red[512];
    const int g = blockIdx.x / C, c = blockIdx.x % C;
    const int64_t e0 = grp_ptr[g], e1 = grp_ptr[g + 1];
    float s = 0.f;
    for (int64_t e = e0 + 1 + threadIdx.x; e < e1; e += blockDim.x) s += dS[ent_row[e] * C + c];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = blockDim.x / 2; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) dY[e0 * C + c] = dStot[c] - red[0];
}

// long groups (one-hot columns: a feature with ~1e5 exceptions): partial[slab][g*C+c] = sum of dS over the exceptions of slab
// `slab` of group g (fixed thread-strided order + tree), then dY[base] = dStot - sum of the slabs in order. Deterministic.
constexpr int BASE_SLAB = 8192;
__global__ void __launch_bounds__(256)
rows_to_baselines_partial_kernel(const float *__restrict__ dS, int C, const int64_t *__restrict__ grp_ptr,
                                 const int64_t *__restrict__ ent_row, float *__restrict__ partial)
{
    __shared__ float red[256];
    const int g = blockIdx.x / C, c = blockIdx.x % C;
    const int64_t e0 = grp_ptr[g] + 1 + (int64_t)blockIdx.y * BASE_SLAB, e1 = min(grp_ptr[g + 1], e0 + BASE_SLAB);
    float s0 = 0.f, s1 = 0.f;
    int64_t e = e0 + threadIdx.x;
    for (; e + 256 < e1; e += 512) {
        s0 += dS[ent_row[e] * C + c];
        s1 += dS[ent_row[e + 256] * C + c];
    }
    if (e < e1) s0 += dS[ent_row[e] * C + c];
    red[threadIdx.x] = s0 + s1;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = red[0];
}

__global__ void rows_to_baselines_final_kernel(int GC, int C, int nslab, const int64_t *__restrict__ grp_ptr,
                                               const float *__restrict__ partial, const float *__restrict__ dStot, float *__restrict__ dY)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= GC) return;
    const int g = t / C, c = t % C;
    float s = 0.f;
    for (int k = 0; k < nslab; ++k) s += partial[(size_t)k * GC + t];
    dY[grp_ptr[g] * C + c] = dStot[c] - s;
}

// out[s,:] = sum over k in [seg_ptr[s], seg_ptr[s+1]) of src[order[k],:]; one warp per (segment, channel block of 32 lanes is
// not needed: lanes stride over the segment's elements, one channel at a time), fixed order -> deterministic
__global__ void gather_segment_sum_kernel(const float *__restrict__ src, const int64_t *__restrict__ order,
                                          const int64_t *__restrict__ seg_ptr, int64_t nseg, int C, float *__restrict__ out)
{
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;     // one warp per (segment, channel)
    const int lane = threadIdx.x & 31;
    if (w >= nseg * C) return;
    const int64_t sg = w / C;
    const int c = (int)(w % C);
    const int64_t k0 = seg_ptr[sg], k1 = seg_ptr[sg + 1];
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;                              // four independent gathers in flight per lane
    int64_t k = k0 + lane;
    if (order) {
        for (; k + 96 < k1; k += 128) {
            s0 += src[order[k] * C + c];
            s1 += src[order[k + 32] * C + c];
            s2 += src[order[k + 64] * C + c];
            s3 += src[order[k + 96] * C + c];
        }
        for (; k < k1; k += 32) s0 += src[order[k] * C + c];
    } else {                                                                     // rows already in segment order
        for (; k + 96 < k1; k += 128) {
            s0 += src[k * C + c];
            s1 += src[(k + 32) * C + c];
            s2 += src[(k + 64) * C + c];
            s3 += src[(k + 96) * C + c];
        }
        for (; k < k1; k += 32) s0 += src[k * C + c];
    }
    float s = (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[w] = s;
}

// few, long segments (one-hot columns: a handful of distinct values shared by millions of entries): one CTA per (segment,
// channel), fixed thread-strided partial sums + a fixed tree -> still deterministic
__global__ void __launch_bounds__(512)
gather_segment_sum_block_kernel(const float *__restrict__ src, const int64_t *__restrict__ order, const int64_t *__restrict__ seg_ptr,
                                int C, float *__restrict__ out)
{
    __shared__ float red[512];
    const int64_t sg = blockIdx.x / C;
    const int c = (int)(blockIdx.x % C);
    const int64_t k0 = seg_ptr[sg], k1 = seg_ptr[sg + 1];
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int64_t k = k0 + threadIdx.x;
    for (; k + 3 * 512 < k1; k += 4 * 512) {
        s0 += src[(order ? order[k] : k) * C + c];
        s1 += src[(order ? order[k + 512] : k + 512) * C + c];
        s2 += src[(order ? order[k + 1024] : k + 1024) * C + c];
        s3 += src[(order ? order[k + 1536] : k + 1536) * C + c];
    }
    for (; k < k1; k += 512) s0 += src[(order ? order[k] : k) * C + c];
    red[threadIdx.x] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    for (int w = 256; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = red[0];
}

}  // namespace

extern "C" int gnan_gather_segment_sum(const float *src, const int64_t *order, const int64_t *seg_ptr, int64_t nseg, int32_t C,
                                       float *out, gnan_stream_t stream)
{
    GNAN_REQUIRE(nseg >= 0 && C >= 1, "gather_segment_sum: bad sizes");
    if (nseg == 0) return GNAN_OK;
    GNAN_REQUIRE(src && seg_ptr && out, "gather_segment_sum: NULL pointer");
    // a warp per (segment, channel) starves the GPU when there are only a few (then necessarily long) segments
    if (nseg * C <= 2 * gnan_sm_count())
        gather_segment_sum_block_kernel<<<(unsigned)(nseg * C), 512, 0, (cudaStream_t)stream>>>(src, order, seg_ptr, C, out);
    else
        gather_segment_sum_kernel<<<(unsigned)ceil_div64(nseg * C * 32, 256), 256, 0, (cudaStream_t)stream>>>(src, order, seg_ptr, nseg, C, out);
    GNAN_LAUNCH_OK();
    return GNAN_OK;
}

extern "C" int gnan_entries_to_rows(const float *Y, int64_t N, int32_t G, int32_t C, const int64_t *grp_ptr,
                                    const int64_t *csr_ptr, const int64_t *csr_eid, const int32_t *ent_grp, float *S0,
                                    float *S, gnan_stream_t stream)
{
    GNAN_REQUIRE(N >= 0 && G >= 1 && C >= 1, "entries_to_rows: bad sizes");
    GNAN_REQUIRE(Y && grp_ptr && csr_ptr && ent_grp && S0 && (N == 0 || S), "entries_to_rows: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    colsum_kernel<true><<<C, 512, 0, st>>>(Y, grp_ptr, G, C, S0);
    GNAN_LAUNCH_OK();
    if (N > 0) {
        entries_to_rows_kernel<<<(unsigned)ceil_div64(N * C, 256), 256, 0, st>>>(Y, N, C, grp_ptr, csr_ptr, csr_eid, ent_grp, S0, S);
        GNAN_LAUNCH_OK();
    }
    return GNAN_OK;
}

// floats of scratch behind dStot: 65*C for the column sums + the slab partial sums of long groups
extern "C" size_t gnan_rows_to_entries_scratch_floats(int32_t G, int32_t C, int64_t E)
{
    return 65 * (size_t)C + (size_t)ceil_div64(E, BASE_SLAB) * (size_t)G * (size_t)C;
}

extern "C" int gnan_rows_to_entries(const float *dS, int64_t N, int32_t G, int32_t C, const int64_t *grp_ptr, int64_t E,
                                    const int64_t *ent_row, float *dStot, float *dY, gnan_stream_t stream)
{
    GNAN_REQUIRE(N >= 0 && G >= 1 && C >= 1 && E >= G, "rows_to_entries: bad sizes");
    GNAN_REQUIRE(grp_ptr && ent_row && dStot && dY && (N == 0 || dS), "rows_to_entries: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = (int)std::max<int64_t>(1, std::min<int64_t>(COLSUM_SLABS, N / 2048));
    colsum_partial_kernel<<<dim3((unsigned)C, (unsigned)nblk), 256, 0, st>>>(dS, N, C, dStot + C);
    GNAN_LAUNCH_OK();
    colsum_final_kernel<<<(unsigned)ceil_div64(C, 64), 64, 0, st>>>(dStot + C, nblk, C, dStot);
    GNAN_LAUNCH_OK();
    rows_to_exceptions_kernel<<<(unsigned)ceil_div64(E * C, 256), 256, 0, st>>>(dS, E, C, ent_row, dY);
    GNAN_LAUNCH_OK();
    if ((E / G) > 2048) {                               // long groups (few features, many exceptions): slabs of 8192 entries in parallel
        const int nslab = (int)ceil_div64(E, BASE_SLAB);
        float *partial = dStot + 65 * (size_t)C;       // [nslab][G*C]
        rows_to_baselines_partial_kernel<<<dim3((unsigned)(G * C), (unsigned)nslab), 256, 0, st>>>(dS, C, grp_ptr, ent_row, partial);
        GNAN_LAUNCH_OK();
        rows_to_baselines_final_kernel<<<(unsigned)ceil_div64((int64_t)G * C, 128), 128, 0, st>>>(G * C, C, nslab, grp_ptr, partial, dStot, dY);
        GNAN_LAUNCH_OK();
        return GNAN_OK;
    }
    rows_to_baselines_kernel<<<(unsigned)(G * C), 64, 0, st>>>(dS, G, C, grp_ptr, ent_row, dStot, dY);
    GNAN_LAUNCH_OK();
    return GNAN_OK;
}
