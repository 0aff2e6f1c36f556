// The tail of the training step on the device: masked loss + its gradient in one pass, and one multi-tensor Adam update.
//
// Reference lines replaced: trainer.py:52-67 (mask the outputs, CrossEntropyLoss / BCEWithLogitsLoss, backward of both) and
// main.py:141 + trainer.py:67 (torch.optim.Adam.step). The reference launches ~10 small kernels for the loss and its backward
// and one fused-optimizer pass; here the loss kernel writes the loss AND d(loss)/d(out) for the whole [N,C] output (zeros on
// the rows without a loss, which is what the aggregation backward uses to skip them), and Adam is one launch over every
// parameter tensor of the model.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace {

constexpr int LOSS_BLOCKS = 128;

// loss_part[blk] = sum over this block's rows of CE(logits[row], label); dlogits rows = scale * (softmax - onehot)
// one warp per listed row; rows == NULL: row m = m. dlogits must be zero-filled when rows != NULL (done by the caller).
__global__ void __launch_bounds__(256)
ce_rows_kernel(const float *__restrict__ logits, const int64_t *__restrict__ rows, const int64_t *__restrict__ labels, int64_t M,
               int64_t N, int C, float scale, float *__restrict__ part, float *__restrict__ dlogits, int32_t *__restrict__ bad)
{
    __shared__ float red[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float acc = 0.f;
    for (int64_t m = (int64_t)blockIdx.x * 8 + w; m < M; m += (int64_t)gridDim.x * 8) {
        const int64_t r = rows ? rows[m] : m;
        const int64_t y = labels[m];
        if (r < 0 || r >= N || y < 0 || y >= C) {
            if (lane == 0) atomicExch(bad, 1);
            continue;
        }
        const float *x = logits + r * C;
        float mx = -INFINITY;
        for (int c = lane; c < C; c += 32) mx = fmaxf(mx, x[c]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float se = 0.f;
        for (int c = lane; c < C; c += 32) se += expf(x[c] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
        const float lse = mx + logf(se);
        if (lane == 0) acc += lse - x[y];
        if (dlogits) {
            float *d = dlogits + r * C;
            for (int c = lane; c < C; c += 32) {
                const float g = scale * (expf(x[c] - lse) - (c == y ? 1.f : 0.f));
                if (rows) atomicAdd(d + c, g);          // a row listed twice accumulates, like index_select's backward
                else d[c] = g;
            }
        }
    }
    if (lane == 0) red[w] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int k = 0; k < 8; ++k) s += red[k];
        part[blockIdx.x] = s;
    }
}

// loss_part[blk] = sum of BCE-with-logits terms (torch's stable form); dlogits = scale * (sigmoid(x) - t)
__global__ void __launch_bounds__(256)
bce_logits_kernel(const float *__restrict__ logits, const float *__restrict__ targets, int64_t M, float scale,
                  float *__restrict__ part, float *__restrict__ dlogits)
{
    __shared__ float red[256];
    float acc = 0.f;
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
        const float x = logits[m], t = targets[m];
        acc += fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
        if (dlogits) dlogits[m] = scale * (1.f / (1.f + expf(-x)) - t);
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}

__global__ void loss_final_kernel(const float *__restrict__ part, int nblk, float scale, float *__restrict__ loss)
{
    float s = 0.f;
    for (int b = 0; b < nblk; ++b) s += part[b];      // fixed order
    *loss = scale * s;
}

// ---- Adam ---------------------------------------------------------------------------------------------------------------
constexpr int ADAM_MAX_TENSORS = 24;
struct AdamTensors {
    float *p[ADAM_MAX_TENSORS];
    const float *g[ADAM_MAX_TENSORS];
    float *m[ADAM_MAX_TENSORS], *v[ADAM_MAX_TENSORS];
    int64_t n[ADAM_MAX_TENSORS];
    int32_t blk0[ADAM_MAX_TENSORS + 1];     // first block of each tensor (blocks of 1024 elements)
    int count;
};

// state[0] = step count (as float, like torch's capturable Adam), state[1] = 1 - beta1^step, state[2] = sqrt(1 - beta2^step)
__global__ void adam_tick_kernel(float *__restrict__ state, float beta1, float beta2)
{
    const float step = state[0] + 1.f;
    state[0] = step;
    state[1] = 1.f - powf(beta1, step);
    state[2] = sqrtf(1.f - powf(beta2, step));
}

__global__ void __launch_bounds__(256)
adam_update_kernel(AdamTensors t, const float *__restrict__ state, float lr, float beta1, float beta2, float eps, float weight_decay)
{
    int ti = 0;
    while (ti + 1 < t.count && (int)blockIdx.x >= t.blk0[ti + 1]) ++ti;
    const int64_t base = (int64_t)((int)blockIdx.x - t.blk0[ti]) * 1024;
    const float bc1 = state[1], bc2s = state[2];
    const float step_size = lr / bc1;
    float *p = t.p[ti], *m = t.m[ti], *v = t.v[ti];
    const float *g = t.g[ti];
    const int64_t n = t.n[ti];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t i = base + k * 256 + threadIdx.x;
        if (i < n) {
            float gi = g[i];
            const float pi = p[i];
            if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
            const float mi = m[i] + (gi - m[i]) * (1.f - beta1);                    // lerp, as torch's fused kernel
            const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
            m[i] = mi;
            v[i] = vi;
            p[i] = pi - step_size * (mi / (sqrtf(vi) / bc2s + eps));
        }
    }
}

}  // namespace

extern "C" size_t gnan_loss_workspace_bytes(void) { return sizeof(float) * LOSS_BLOCKS + sizeof(int32_t) * 4; }

extern "C" int gnan_cross_entropy_rows(const float *logits, int64_t N, int32_t C, const int64_t *rows, const int64_t *labels, int64_t M,
                                       float scale, float *loss, float *dlogits, int32_t *bad_index_flag, void *workspace,
                                       size_t workspace_bytes, gnan_stream_t stream)
{
    GNAN_REQUIRE(N >= 0 && M >= 0 && C >= 1, "cross_entropy_rows: bad sizes");
    GNAN_REQUIRE(loss && bad_index_flag && (M == 0 || (logits && labels)), "cross_entropy_rows: NULL pointer");
    if (!workspace || workspace_bytes < gnan_loss_workspace_bytes()) {
        gnan_set_error("cross_entropy_rows: workspace %zu < %zu bytes", workspace_bytes, gnan_loss_workspace_bytes());
        return GNAN_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float *part = (float *)workspace;
    if (dlogits && rows && N > 0) GNAN_CUDA(cudaMemsetAsync(dlogits, 0, sizeof(float) * (size_t)N * C, st));
    const int nblk = (int)std::max<int64_t>(1, std::min<int64_t>(LOSS_BLOCKS, ceil_div64(M, 8)));
    ce_rows_kernel<<<nblk, 256, 0, st>>>(logits, rows, labels, M, N, C, scale, part, dlogits, bad_index_flag);
    GNAN_LAUNCH_OK();
    loss_final_kernel<<<1, 1, 0, st>>>(part, nblk, scale, loss);
    GNAN_LAUNCH_OK();
    return GNAN_OK;
}

extern "C" int gnan_bce_with_logits(const float *logits, const float *targets, int64_t M, float scale, float *loss, float *dlogits,
                                    void *workspace, size_t workspace_bytes, gnan_stream_t stream)
{
    GNAN_REQUIRE(M >= 0 && loss && (M == 0 || (logits && targets)), "bce_with_logits: bad arguments");
    if (!workspace || workspace_bytes < gnan_loss_workspace_bytes()) {
        gnan_set_error("bce_with_logits: workspace %zu < %zu bytes", workspace_bytes, gnan_loss_workspace_bytes());
        return GNAN_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float *part = (float *)workspace;
    const int nblk = (int)std::max<int64_t>(1, std::min<int64_t>(LOSS_BLOCKS, ceil_div64(M, 256)));
    bce_logits_kernel<<<nblk, 256, 0, st>>>(logits, targets, M, scale, part, dlogits);
    GNAN_LAUNCH_OK();
    loss_final_kernel<<<1, 1, 0, st>>>(part, nblk, scale, loss);
    GNAN_LAUNCH_OK();
    return GNAN_OK;
}

extern "C" int gnan_adam_step(int32_t n_tensors, float *const *params, const float *const *grads, float *const *exp_avg,
                              float *const *exp_avg_sq, const int64_t *numel, float *state, float lr, float beta1, float beta2,
                              float eps, float weight_decay, gnan_stream_t stream)
{
    GNAN_REQUIRE(n_tensors >= 0 && state, "adam_step: bad arguments");
    GNAN_REQUIRE(n_tensors == 0 || (params && grads && exp_avg && exp_avg_sq && numel), "adam_step: NULL pointer table");
    cudaStream_t st = (cudaStream_t)stream;
    adam_tick_kernel<<<1, 1, 0, st>>>(state, beta1, beta2);
    GNAN_LAUNCH_OK();
    for (int t0 = 0; t0 < n_tensors; t0 += ADAM_MAX_TENSORS) {
        AdamTensors t;
        t.count = std::min(ADAM_MAX_TENSORS, n_tensors - t0);
        int blk = 0;
        for (int k = 0; k < t.count; ++k) {
            GNAN_REQUIRE(params[t0 + k] && grads[t0 + k] && exp_avg[t0 + k] && exp_avg_sq[t0 + k] && numel[t0 + k] >= 0,
                         "adam_step: NULL tensor %d", t0 + k);
            t.p[k] = params[t0 + k]; t.g[k] = grads[t0 + k]; t.m[k] = exp_avg[t0 + k]; t.v[k] = exp_avg_sq[t0 + k];
            t.n[k] = numel[t0 + k];
            t.blk0[k] = blk;
            blk += (int)ceil_div64(numel[t0 + k], 1024);
        }
        t.blk0[t.count] = blk;
        if (blk == 0) continue;
        adam_update_kernel<<<blk, 256, 0, st>>>(t, state, lr, beta1, beta2, eps, weight_decay);
        GNAN_LAUNCH_OK();
    }
    return GNAN_OK;
}
