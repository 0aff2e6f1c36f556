// Deep graphs: hop distances beyond 254 (long chains, grids, road-like graphs) do not fit the uint8 hop matrix. The reference has
// no depth limit (scipy Dijkstra returns floats: pre_process_datasets.py:109-121), so this file carries the same path with int16
// hops (-1 = unreachable, levels 0..32766): a warp-per-source BFS, the level histogram, and the aggregation with its backward.
// It is the general, slower form (2 bytes per pair, direct table lookups instead of bin sums / tensor cores): the uint8 kernels
// stay the path for every graph whose diameter fits them, preprocess.apsp switches over only when the uint8 BFS overflows.
//
// Reference lines replaced: pre_process_datasets.py:104-142 (distances + normaliser) and GNAN.py:64-79 / models.py:366-384.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int W16_MAX_LEVEL = 32766;

// one warp per source (same scheme as apsp_bfs_kernel: frontier queue + claimed bitmap in the workspace)
__global__ void __launch_bounds__(256)
apsp_bfs16_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col, int N, int src_begin, int src_end,
                  int16_t *__restrict__ hop, int64_t ld, int32_t *__restrict__ overflow, int32_t *__restrict__ max_level,
                  int32_t *__restrict__ queues, uint32_t *__restrict__ bitmaps, int bm_words, int64_t nwarps)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= nwarps) return;
    int32_t *q = queues + warp * (int64_t)N;
    uint32_t *bm = bitmaps + warp * (int64_t)bm_words;
    int deepest = 0;
    for (int64_t s = src_begin + warp; s < src_end; s += nwarps) {
        int16_t *row = hop + (s - src_begin) * ld;
        for (int64_t v = lane; v < ld; v += 32) row[v] = -1;
        for (int t = lane; t < bm_words; t += 32) bm[t] = 0u;
        __syncwarp();
        if (lane == 0) {
            q[0] = (int32_t)s;
            bm[s >> 5] = 1u << (s & 31);
            row[s] = 0;
        }
        __syncwarp();
        int head = 0, tail = 1, level = 0;
        while (head < tail) {
            ++level;
            const int16_t lv = (int16_t)min(level, W16_MAX_LEVEL);
            int new_tail = tail;
            for (int base = head; base < tail; base += 32) {
                const int idx = base + lane;
                int e0 = 0, e1 = 0;
                if (idx < tail) {
                    const int v = q[idx];
                    e0 = rowptr[v];
                    e1 = rowptr[v + 1];
                }
                int maxdeg = e1 - e0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) maxdeg = max(maxdeg, __shfl_xor_sync(0xffffffffu, maxdeg, o));
                for (int k = 0; k < maxdeg; ++k) {
                    bool won = false;
                    int t = -1;
                    if (e0 + k < e1) {
                        t = col[e0 + k];
                        const uint32_t bit = 1u << (t & 31);
                        won = !(atomicOr(bm + (t >> 5), bit) & bit);
                    }
                    const uint32_t mask = __ballot_sync(0xffffffffu, won);
                    if (won) {
                        q[new_tail + __popc(mask & ((1u << lane) - 1))] = t;
                        row[t] = lv;
                    }
                    new_tail += __popc(mask);
                }
            }
            __syncwarp();
            if (new_tail > tail) {
                if (level > W16_MAX_LEVEL && lane == 0) atomicExch(overflow, 1);
                deepest = max(deepest, min(level, W16_MAX_LEVEL));
            }
            head = tail;
            tail = new_tail;
        }
        __syncwarp();
    }
    if (lane == 0 && deepest > 0) atomicMax(max_level, deepest);
}

// cnt[i, d] = #{j < N: hop[i,j] = d} (d < nbins-1), cnt[i, nbins-1] = #unreachable (or deeper than the table); one warp per row
__global__ void __launch_bounds__(256)
level_counts16_kernel(const int16_t *__restrict__ hop, int64_t R, int64_t N, int64_t ld, int32_t *__restrict__ cnt, int nbins)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= R) return;
    int32_t *c = cnt + i * nbins;
    for (int d = lane; d < nbins; d += 32) c[d] = 0;
    __syncwarp();
    const int16_t *row = hop + i * ld;
    for (int64_t j = lane; j < N; j += 32) {
        const int h = row[j];
        atomicAdd(c + ((h < 0 || h >= nbins - 1) ? nbins - 1 : h), 1);
    }
}

struct WideArgs {
    const int16_t *hop;
    int64_t R, N, ld;
    const float *T;
    int per_row, nbins, Cr;
    const float *rscale;
    const float *S;
    int C;
};

__device__ __forceinline__ int wide_bin(int h, int nbins) { return (h < 0 || h >= nbins - 1) ? nbins - 1 : h; }

__device__ __forceinline__ float wide_w(const WideArgs &a, int64_t i, int d, int c)
{
    const int cr = a.Cr == 1 ? 0 : c;
    const float t = a.per_row ? a.T[(i * a.nbins + d) * a.Cr + cr] : a.T[d * a.Cr + cr];
    return a.rscale ? t * a.rscale[i * a.nbins + d] : t;
}

// out[i,c] = sum_j W[i, b(hop_ij), c] S[j,c]: one warp per row, 4 channels per sweep
__global__ void __launch_bounds__(256)
agg_rows16_fwd_kernel(WideArgs a, float *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= a.R) return;
    const int16_t *row = a.hop + i * a.ld;
    for (int c0 = 0; c0 < a.C; c0 += 4) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int64_t j = lane; j < a.N; j += 32) {
            const int d = wide_bin(row[j], a.nbins);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
                if (c0 + cc < a.C) acc[cc] = fmaf(wide_w(a, i, d, c0 + cc), a.S[j * a.C + c0 + cc], acc[cc]);
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            float v = acc[cc];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && c0 + cc < a.C) out[i * a.C + c0 + cc] = v;
        }
    }
}

// dS[j,c] = sum_i W[i, b(hop_ij), c] g[i,c]: one thread per column (coalesced hop reads), rows without a loss skipped
__global__ void __launch_bounds__(256)
agg_rows16_ds_kernel(WideArgs a, const float *__restrict__ g, const uint8_t *__restrict__ row_active, float *__restrict__ dS)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.N) return;
    for (int c0 = 0; c0 < a.C; c0 += 4) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int64_t i = 0; i < a.R; ++i) {
            if (!row_active[i]) continue;
            const int d = wide_bin(a.hop[i * a.ld + j], a.nbins);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
                if (c0 + cc < a.C) acc[cc] = fmaf(wide_w(a, i, d, c0 + cc), g[i * a.C + c0 + cc], acc[cc]);
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
            if (c0 + cc < a.C) dS[j * a.C + c0 + cc] = acc[cc];
    }
}

// dT[ti, d, c'] += rscale[i,d] * sum_{j: b(hop_ij) = d} sum_{c (all c if Cr == 1, else c')} g[i,c] S[j,c]: one warp per active row,
// float atomics into dT (zero-initialised by the host side); the summation order is not fixed on this fallback path
__global__ void __launch_bounds__(256)
agg_rows16_dt_kernel(WideArgs a, const float *__restrict__ g, const uint8_t *__restrict__ row_active, float *__restrict__ dT)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= a.R || !row_active[i]) return;
    const int16_t *row = a.hop + i * a.ld;
    float *dTi = dT + (a.per_row ? i * a.nbins * a.Cr : 0);
    for (int64_t j = lane; j < a.N; j += 32) {
        const int d = wide_bin(row[j], a.nbins);
        const float rs = a.rscale ? a.rscale[i * a.nbins + d] : 1.f;
        if (a.Cr == 1) {
            float v = 0.f;
            for (int c = 0; c < a.C; ++c) v = fmaf(g[i * a.C + c], a.S[j * a.C + c], v);
            atomicAdd(dTi + d, rs * v);
        } else {
            for (int c = 0; c < a.C; ++c) atomicAdd(dTi + d * a.Cr + c, rs * g[i * a.C + c] * a.S[j * a.C + c]);
        }
    }
}

__global__ void row_active_kernel(const float *__restrict__ g, int64_t R, int C, uint8_t *__restrict__ flags)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    bool nz = false;
    for (int c = 0; c < C; ++c) nz |= g[i * C + c] != 0.f;
    flags[i] = nz ? 1 : 0;
}

int check_wide(const char *who, const int16_t *hop, int64_t R, int64_t N, int64_t ld, const float *T, int nbins, int Cr, const float *S, int C)
{
    GNAN_REQUIRE(R >= 0 && N >= 0 && ld >= N, "%s: bad shape", who);
    GNAN_REQUIRE(R * N == 0 || (hop && T && S), "%s: NULL hop/T/S", who);
    GNAN_REQUIRE(nbins >= 2 && nbins <= 32768, "%s: nbins %d out of [2,32768]", who, nbins);
    GNAN_REQUIRE(C >= 1 && (Cr == 1 || Cr == C), "%s: Cr must be 1 or C (Cr=%d C=%d)", who, Cr, C);
    return GNAN_OK;
}

}  // namespace

extern "C" size_t gnan_apsp_bfs16_workspace_bytes(int32_t N, int32_t n_sources)
{
    const int64_t nwarps = std::max<int64_t>(1, std::min<int64_t>(n_sources, (int64_t)gnan_sm_count() * 16));
    return (size_t)nwarps * ((size_t)N * 4 + (size_t)((N + 31) / 32) * 4);
}

extern "C" int gnan_apsp_bfs16(const int32_t *rowptr, const int32_t *col, int32_t N, int32_t src_begin, int32_t src_end, int16_t *hop,
                               int64_t ld_hop, int32_t *overflow_flag, int32_t *max_level, void *workspace, size_t workspace_bytes,
                               gnan_stream_t stream)
{
    GNAN_REQUIRE(N >= 0 && src_begin >= 0 && src_end >= src_begin && src_end <= N && ld_hop >= N, "apsp_bfs16: bad arguments");
    const int R = src_end - src_begin;
    if (R == 0) return GNAN_OK;
    GNAN_REQUIRE(rowptr && hop && overflow_flag && max_level, "apsp_bfs16: NULL pointer");
    const size_t need = gnan_apsp_bfs16_workspace_bytes(N, R);
    if (!workspace || workspace_bytes < need) {
        gnan_set_error("apsp_bfs16: workspace %zu < %zu bytes", workspace_bytes, need);
        return GNAN_ERR_WORKSPACE;
    }
    const int64_t nwarps = std::max<int64_t>(1, std::min<int64_t>(R, (int64_t)gnan_sm_count() * 16));
    const int bm_words = (N + 31) / 32;
    int32_t *queues = (int32_t *)workspace;
    uint32_t *bitmaps = (uint32_t *)(queues + nwarps * (int64_t)N);
    apsp_bfs16_kernel<<<(unsigned)ceil_div64(nwarps * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        rowptr, col, N, src_begin, src_end, hop, ld_hop, overflow_flag, max_level, queues, bitmaps, bm_words, nwarps);
    GNAN_LAUNCH_OK();
    return GNAN_OK;
}

extern "C" int gnan_level_counts16(const int16_t *hop, int64_t R, int64_t N, int64_t ld_hop, int32_t *cnt, int32_t nbins, gnan_stream_t stream)
{
    GNAN_REQUIRE(R >= 0 && N >= 0 && ld_hop >= N && nbins >= 2, "level_counts16: bad arguments");
    if (R == 0) return GNAN_OK;
    GNAN_REQUIRE(hop && cnt, "level_counts16: NULL pointer");
    level_counts16_kernel<<<(unsigned)ceil_div64(R * 32, 256), 256, 0, (cudaStream_t)stream>>>(hop, R, N, ld_hop, cnt, nbins);
    GNAN_LAUNCH_OK();
    return GNAN_OK;
}

extern "C" int gnan_aggregate_rows16_fwd(const int16_t *hop, int64_t R, int64_t N, int64_t ld_hop, const float *T, int table_per_row,
                                         int32_t nbins, int32_t Cr, const float *rscale, const float *S, int32_t C, float *out,
                                         gnan_stream_t stream)
{
    int rc = check_wide("aggregate_rows16_fwd", hop, R, N, ld_hop, T, nbins, Cr, S, C);
    if (rc) return rc;
    if (R == 0) return GNAN_OK;
    GNAN_REQUIRE(out != nullptr, "aggregate_rows16_fwd: NULL out");
    WideArgs a{hop, R, N, ld_hop, T, table_per_row, nbins, Cr, rscale, S, C};
    agg_rows16_fwd_kernel<<<(unsigned)ceil_div64(R * 32, 256), 256, 0, (cudaStream_t)stream>>>(a, out);
    GNAN_LAUNCH_OK();
    return GNAN_OK;
}

extern "C" int gnan_aggregate_rows16_bwd(const int16_t *hop, int64_t R, int64_t N, int64_t ld_hop, const float *T, int table_per_row,
                                         int32_t nbins, int32_t Cr, const float *rscale, const float *S, int32_t C, const float *g,
                                         float *dS, float *dT, uint8_t *row_flags_ws /* [R] */, gnan_stream_t stream)
{
    int rc = check_wide("aggregate_rows16_bwd", hop, R, N, ld_hop, T, nbins, Cr, S, C);
    if (rc) return rc;
    GNAN_REQUIRE(dS && dT && (R == 0 || (g && row_flags_ws)), "aggregate_rows16_bwd: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nT = (size_t)(table_per_row ? R : 1) * nbins * Cr;
    GNAN_CUDA(cudaMemsetAsync(dT, 0, sizeof(float) * nT, st));
    if (R == 0) {
        if (N > 0) GNAN_CUDA(cudaMemsetAsync(dS, 0, sizeof(float) * (size_t)N * C, st));
        return GNAN_OK;
    }
    WideArgs a{hop, R, N, ld_hop, T, table_per_row, nbins, Cr, rscale, S, C};
    row_active_kernel<<<(unsigned)ceil_div64(R, 256), 256, 0, st>>>(g, R, C, row_flags_ws);
    GNAN_LAUNCH_OK();
    if (N > 0) {
        agg_rows16_ds_kernel<<<(unsigned)ceil_div64(N, 256), 256, 0, st>>>(a, g, row_flags_ws, dS);
        GNAN_LAUNCH_OK();
    }
    agg_rows16_dt_kernel<<<(unsigned)ceil_div64(R * 32, 256), 256, 0, st>>>(a, g, row_flags_ws, dT);
    GNAN_LAUNCH_OK();
    return GNAN_OK;
}
