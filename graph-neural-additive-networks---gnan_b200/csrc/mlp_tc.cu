// tcgen05 / TMEM path of the grouped shape-function MLPs (H = 64, 3 layers): the 64x64 hidden contraction runs on the
// 5th-generation tensor cores as a 3-term TF32 split (hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM), which keeps
// fp32-level accuracy (measured 1e-6 norm-wise by scratch/tc_probe.cu) at tensor-core speed.
//
// Reference lines replaced: GNAN.py:57-62,157 (forward), autograd through them (backward).
//
// Mapping (forward): a CTA owns two 128-row tiles ("groups"); a row thread owns one row = one TMEM lane.
//   layer 1: the row thread computes a[i] = relu(x*w1[i] + b1[i]) in registers, splits it into tf32 hi/lo and stores both
//            straight into TMEM with tcgen05.st (the A operand never touches shared memory);
//   layer 2: one elected thread issues tcgen05.mma kind::tf32 with A from TMEM and B = W2 (hi/lo, K-major core-matrix
//            layout) from shared memory, 8 K-steps x 3 terms, accumulating z[128x64] in TMEM; tcgen05.commit -> mbarrier;
//   layer 3: the row thread reads its z row back with tcgen05.ld, adds b2, ReLU, and contracts with wo in registers,
//            accumulating S[row, c] over the features of the chunk.
// Two producer warps stage the next feature's W2 split + small vectors into a double-buffered shared-memory slot while the
// current feature is consumed; the two row groups alternate so one group's CUDA-core work overlaps the other's MMAs.
#include <algorithm>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int HID = 64;
constexpr int ROWS = 128;                 // rows per group (= TMEM lanes)
constexpr int ROW_THREADS = 256;          // two groups
constexpr int PROD_THREADS = 64;
constexpr int FWD_THREADS = ROW_THREADS + PROD_THREADS;
constexpr int CT_MAX = 8;                 // output channels held in registers

// shared-memory slot of one feature: B hi/lo in UMMA K-major no-swizzle core-matrix layout + small vectors
struct __align__(16) FeatSlot {
    float bhi[HID * HID];
    float blo[HID * HID];
    float w1[HID], b1[HID], b2[HID];
    float wo[CT_MAX][HID];
    float bo[CT_MAX];
};

struct FwdSmem {
    FeatSlot slot[2];
    uint64_t b_full[2], b_empty[2], d_full[2];
    uint32_t tmem_base;
};

struct TcArgs {
    const float *u;
    int64_t R, ldu;
    int G, C;
    const float *w1, *b1, *wh, *bh, *wo, *bo;
    uint32_t drop_thresh;
    float drop_scale;
    uint64_t seed;
    int single_pass;  // 1 = plain tf32 (no lo terms)
};

__device__ __forceinline__ uint64_t tc_drop_key(const TcArgs &a, int layer, int g, int64_t row, int unit)
{
    return ((((uint64_t)layer * a.G + g) * (uint64_t)a.R + (uint64_t)row) * HID) + unit;   // same keys as the fp32 path
}

// element (n,k) of a 64x64 K-major operand -> float offset in the core-matrix layout (LBO = 128 B, SBO = 2048 B)
__device__ __forceinline__ int bidx(int n, int k) { return (n >> 3) * 512 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3); }
constexpr uint32_t B_LBO = 128, B_SBO = 2048, B_KSTEP = 256;   // bytes; one K-step = 8 tf32 = two 16-byte chunks

__device__ __forceinline__ void produce_slot(FeatSlot &sl, const TcArgs &a, int g, int pt)
{
    const float *W = a.wh + (size_t)g * HID * HID;     // [j][i] = [n][k]
#pragma unroll 4
    for (int t = 0; t < 16; ++t) {
        const int idx = pt + PROD_THREADS * t;          // 1024 float4 chunks: n fastest -> conflict-free 128-bit stores
        const int n = idx & 63, kc = idx >> 6;
        const float4 v = __ldg(reinterpret_cast<const float4 *>(W + n * HID + kc * 4));
        uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
        split_tf32(v.x, h0, l0); split_tf32(v.y, h1, l1); split_tf32(v.z, h2, l2); split_tf32(v.w, h3, l3);
        const int o = (n >> 3) * 512 + kc * 32 + (n & 7) * 4;
        *reinterpret_cast<uint4 *>(sl.bhi + o) = make_uint4(h0, h1, h2, h3);
        *reinterpret_cast<uint4 *>(sl.blo + o) = make_uint4(l0, l1, l2, l3);
    }
    sl.w1[pt] = __ldg(a.w1 + (size_t)g * HID + pt);
    sl.b1[pt] = a.b1 ? __ldg(a.b1 + (size_t)g * HID + pt) : 0.f;
    sl.b2[pt] = a.bh ? __ldg(a.bh + (size_t)g * HID + pt) : 0.f;
    for (int c = 0; c < a.C; ++c) sl.wo[c][pt] = __ldg(a.wo + ((size_t)g * a.C + c) * HID + pt);
    if (pt < CT_MAX) sl.bo[pt] = (a.bo && pt < a.C) ? __ldg(a.bo + (size_t)g * a.C + pt) : 0.f;
}

// grid (row-pair tiles, feature chunks); Spart[chunk][R][C]
template <int CT>
__global__ void __launch_bounds__(FWD_THREADS, 1)
mlp_tc_fwd_kernel(TcArgs a, int KC, float *__restrict__ Spart)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    FwdSmem &sm = *reinterpret_cast<FwdSmem *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g0 = blockIdx.y * KC;
    const int ng = min(KC, a.G - g0);

    if (warp == 0) tmem_alloc(smem_u32(&sm.tmem_base), 512);
    if (tid == 32) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&sm.b_full[s]), PROD_THREADS);
            mbar_init(smem_u32(&sm.b_empty[s]), ROW_THREADS / 32);
            mbar_init(smem_u32(&sm.d_full[s]), 1);
        }
        mbar_init_fence();
    }
    if (tid >= ROW_THREADS) {   // zero the unused output-channel rows once
        const int pt = tid - ROW_THREADS;
        for (int s = 0; s < 2; ++s)
            for (int c = a.C; c < CT_MAX; ++c) sm.slot[s].wo[c][pt] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    if (tid >= ROW_THREADS) {
        // ===== producers =====
        const int pt = tid - ROW_THREADS;
        for (int kk = 0; kk < ng; ++kk) {
            const int s = kk & 1, n = kk >> 1;
            mbar_wait(smem_u32(&sm.b_empty[s]), (n & 1) ^ 1);
            produce_slot(sm.slot[s], a, g0 + kk, pt);
            fence_async_smem();                      // make the generic-proxy writes visible to the tensor core
            mbar_arrive(smem_u32(&sm.b_full[s]));
        }
    } else {
        // ===== row threads =====
        const int grp = tid >> 7, rt = tid & 127;
        const int64_t row = ((int64_t)blockIdx.x * 2 + grp) * ROWS + rt;
        const bool row_ok = row < a.R;
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)grp * 256;
        const uint32_t colA_hi = 0, colA_lo = 64, colD = 128;
        const uint32_t idesc = umma_idesc_tf32(128, 64);
        float Sacc[CT];
#pragma unroll
        for (int c = 0; c < CT; ++c) Sacc[c] = 0.f;
        float x_next = row_ok ? __ldg(a.u + row * a.ldu + g0) : 0.f;

        for (int kk = 0; kk < ng; ++kk) {
            const int s = kk & 1, n = kk >> 1, g = g0 + kk;
            const FeatSlot &sl = sm.slot[s];
            const float x = x_next;
            if (kk + 1 < ng) x_next = row_ok ? __ldg(a.u + row * a.ldu + g + 1) : 0.f;
            mbar_wait(smem_u32(&sm.b_full[s]), n & 1);
            // ---- layer 1 -> TMEM (A operand, hi | lo)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const float4 w = *reinterpret_cast<const float4 *>(sl.w1 + q * 32 + i4 * 4);
                    const float4 b = *reinterpret_cast<const float4 *>(sl.b1 + q * 32 + i4 * 4);
                    float v[4] = {fmaxf(fmaf(x, w.x, b.x), 0.f), fmaxf(fmaf(x, w.y, b.y), 0.f),
                                  fmaxf(fmaf(x, w.z, b.z), 0.f), fmaxf(fmaf(x, w.w, b.w), 0.f)};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (a.drop_thresh)
                            v[e] *= gnan_dropout_mul(a.seed, tc_drop_key(a, 0, g, row, q * 32 + i4 * 4 + e), a.drop_thresh, a.drop_scale);
                        split_tf32(v[e], hi[i4 * 4 + e], lo[i4 * 4 + e]);
                    }
                }
                tmem_st32(lane_base + colA_hi + q * 32, hi);
                if (!a.single_pass) tmem_st32(lane_base + colA_lo + q * 32, lo);
            }
            tmem_wait_st();
            tc_fence_before();
            named_bar_sync(1 + grp, ROWS);
            // ---- layer 2 on the tensor core
            if (rt == 0) {
                tc_fence_after();
                const uint32_t d_t = tmem + (uint32_t)grp * 256 + colD;
                const uint32_t a_hi = tmem + (uint32_t)grp * 256 + colA_hi, a_lo = tmem + (uint32_t)grp * 256 + colA_lo;
                const uint32_t bh = smem_u32(sl.bhi), bl = smem_u32(sl.blo);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    umma_tf32_ts(d_t, a_hi + ks * 8, umma_desc_kmajor(bh + ks * B_KSTEP, B_LBO, B_SBO), idesc, ks > 0);
                if (!a.single_pass) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        umma_tf32_ts(d_t, a_lo + ks * 8, umma_desc_kmajor(bh + ks * B_KSTEP, B_LBO, B_SBO), idesc, 1);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        umma_tf32_ts(d_t, a_hi + ks * 8, umma_desc_kmajor(bl + ks * B_KSTEP, B_LBO, B_SBO), idesc, 1);
                }
                umma_commit(smem_u32(&sm.d_full[grp]));
            }
            __syncwarp();
            mbar_wait(smem_u32(&sm.d_full[grp]), kk & 1);
            tc_fence_after();
            // ---- layer 3 in registers
            float y[CT];
#pragma unroll
            for (int c = 0; c < CT; ++c) y[c] = 0.f;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                uint32_t d[32];
                tmem_ld32(lane_base + colD + q * 32, d);
                tmem_wait_ld();
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 b = *reinterpret_cast<const float4 *>(sl.b2 + q * 32 + j4 * 4);
                    float h[4] = {fmaxf(__uint_as_float(d[j4 * 4 + 0]) + b.x, 0.f), fmaxf(__uint_as_float(d[j4 * 4 + 1]) + b.y, 0.f),
                                  fmaxf(__uint_as_float(d[j4 * 4 + 2]) + b.z, 0.f), fmaxf(__uint_as_float(d[j4 * 4 + 3]) + b.w, 0.f)};
                    if (a.drop_thresh) {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            h[e] *= gnan_dropout_mul(a.seed, tc_drop_key(a, 1, g, row, q * 32 + j4 * 4 + e), a.drop_thresh, a.drop_scale);
                    }
#pragma unroll
                    for (int c = 0; c < CT; ++c) {
                        const float4 w = *reinterpret_cast<const float4 *>(sl.wo[c] + q * 32 + j4 * 4);
                        y[c] = fmaf(h[0], w.x, y[c]);
                        y[c] = fmaf(h[1], w.y, y[c]);
                        y[c] = fmaf(h[2], w.z, y[c]);
                        y[c] = fmaf(h[3], w.w, y[c]);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < CT; ++c) Sacc[c] += y[c] + sl.bo[c];
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&sm.b_empty[s]));
        }
        if (row_ok) {
            float *out = Spart + ((size_t)blockIdx.y * a.R + row) * a.C;
#pragma unroll
            for (int c = 0; c < CT; ++c)
                if (c < a.C) out[c] = Sacc[c];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

struct TcFwdPlan { int KC, nchunk; int64_t ntile; };

TcFwdPlan plan_tc_fwd(int64_t R, const gnan_mlp_params *p)
{
    TcFwdPlan pl;
    pl.ntile = ceil_div64(R, 2 * ROWS);
    const int target = 2 * gnan_sm_count();                    // ~2 waves of one CTA per SM
    int nchunk = (int)std::max<int64_t>(1, std::min<int64_t>(p->G, ceil_div64(target, pl.ntile)));
    pl.KC = (int)ceil_div64(p->G, nchunk);
    if (pl.KC < 8) pl.KC = std::min(8, (int)p->G);             // amortise the CTA prologue (TMEM allocation, barriers)
    pl.nchunk = (int)ceil_div64(p->G, pl.KC);
    return pl;
}

TcArgs make_tc_args(const float *u, int64_t R, int64_t ldu, const gnan_mlp_params *p, float dropout_p, uint64_t seed, int precision)
{
    TcArgs a;
    a.u = u; a.R = R; a.ldu = ldu; a.G = p->G; a.C = p->C;
    a.w1 = p->w1; a.b1 = p->b1; a.wh = p->wh; a.bh = p->bh; a.wo = p->wo; a.bo = p->bo;
    a.drop_thresh = dropout_p > 0.f ? gnan_dropout_thresh(dropout_p) : 0u;
    a.drop_scale = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
    a.seed = seed;
    a.single_pass = precision == GNAN_PREC_TF32;
    return a;
}

template <int CT>
int launch_tc_fwd(const TcArgs &a, const TcFwdPlan &pl, float *Spart, cudaStream_t st)
{
    const size_t smem = sizeof(FwdSmem) + 1024;
    GNAN_CUDA(cudaFuncSetAttribute(mlp_tc_fwd_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)pl.ntile, (unsigned)pl.nchunk);
    mlp_tc_fwd_kernel<CT><<<grid, FWD_THREADS, smem, st>>>(a, pl.KC, Spart);
    GNAN_LAUNCH_OK();
    return GNAN_OK;
}

}  // namespace

int gnan_mlp_tc_supported(const gnan_mlp_params *p, int precision)
{
    return (precision == GNAN_PREC_TF32X3 || precision == GNAN_PREC_TF32) && p->H == HID && p->n_layers == 3 && p->C <= CT_MAX;
}

// forward only for now: the backward of the tensor-core path still runs the fp32 kernel (same dropout keys)
int gnan_mlp_tc_bwd_supported(const gnan_mlp_params *, int) { return 0; }

size_t gnan_mlp_tc_workspace_bytes(int64_t R, const gnan_mlp_params *p, int backward, int precision)
{
    (void)precision;
    if (backward) return 0;
    const TcFwdPlan pl = plan_tc_fwd(R, p);
    return pl.nchunk > 1 ? sizeof(float) * (size_t)pl.nchunk * R * p->C : 0;
}

int gnan_mlp_tc_fwd(const float *u, int64_t R, int64_t ldu, const gnan_mlp_params *p, float dropout_p, uint64_t seed,
                    int precision, float *S, void *ws, size_t ws_bytes, cudaStream_t st)
{
    const TcFwdPlan pl = plan_tc_fwd(R, p);
    float *Spart = S;
    if (pl.nchunk > 1) {
        const size_t need = sizeof(float) * (size_t)pl.nchunk * R * p->C;
        if (!ws || ws_bytes < need) {
            gnan_set_error("mlp_fwd(tc): workspace %zu < %zu bytes", ws_bytes, need);
            return GNAN_ERR_WORKSPACE;
        }
        Spart = (float *)ws;
    }
    const TcArgs a = make_tc_args(u, R, ldu, p, dropout_p, seed, precision);
    int rc;
    if (p->C == 1) rc = launch_tc_fwd<1>(a, pl, Spart, st);
    else if (p->C == 2) rc = launch_tc_fwd<2>(a, pl, Spart, st);
    else if (p->C <= 4) rc = launch_tc_fwd<4>(a, pl, Spart, st);
    else rc = launch_tc_fwd<8>(a, pl, Spart, st);
    if (rc) return rc;
    if (pl.nchunk > 1) return gnan_reduce_chunks(Spart, pl.nchunk, (size_t)R * p->C, (size_t)R * p->C, S, st);
    return GNAN_OK;
}

int gnan_mlp_tc_bwd(const float *, int64_t, int64_t, const gnan_mlp_params *, float, uint64_t, int, const float *,
                    const gnan_mlp_grads *, void *, size_t, cudaStream_t)
{
    gnan_set_error("tcgen05 mlp backward not built");
    return GNAN_ERR_UNSUPPORTED;
}
