// Probe 2: SS-mode tcgen05.mma kind::tf32 with BOTH operands MN-major (no swizzle):
//   D[m][n] = sum_r A[r][m] * B[r][n],  r = 0..127 (K), m = 0..127 (M), n = 0..63 (N)
// and: the same K-major buffer of W2 reused as an MN-major B operand (W2^T) with swapped LBO/SBO in TS mode.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "../graph-neural-additive-networks---gnan_b200/csrc/tc_ptx.cuh"

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) { return umma_desc_kmajor(saddr, lbo, sbo); }

// test 1: SS, A MN-major [r][m] M=128, B MN-major [r][n] N=64, K=128
__global__ void __launch_bounds__(128) probe_ss(const float *A, const float *B, float *D)
{
    extern __shared__ __align__(1024) float sm[];
    float *sA = sm;                 // 128 m x 128 r  : chunk c4 = m/4 (32 chunks) each 2048 B : [c4][r/8][r%8][4]
    float *sB = sm + 128 * 128;     // 64 n x 128 r   : 16 chunks
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tb_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(smem_u32(&tb_s), 64);
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_init_fence(); }
    // thread r writes row r of A and B
    const int r = tid;
    for (int c4 = 0; c4 < 32; ++c4)
        *reinterpret_cast<float4 *>(sA + c4 * 512 + (r >> 3) * 32 + (r & 7) * 4) = *reinterpret_cast<const float4 *>(A + r * 128 + c4 * 4);
    for (int c4 = 0; c4 < 16; ++c4)
        *reinterpret_cast<float4 *>(sB + c4 * 512 + (r >> 3) * 32 + (r & 7) * 4) = *reinterpret_cast<const float4 *>(B + r * 64 + c4 * 4);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tb_s;
    if (tid == 0) {
        const uint32_t idesc = umma_idesc_tf32(128, 64, 1, 1);
        for (int ks = 0; ks < 16; ++ks)   // K-step = 8 rows = one k-block (128 B apart); SBO (between 4-wide mn blocks) = 2048
            umma_tf32_ss(tb, umma_desc(smem_u32(sA) + ks * 128, 128, 2048), umma_desc(smem_u32(sB) + ks * 128, 128, 2048), idesc, ks > 0);
        umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    uint32_t d[64];
    tmem_ld64(tb + ((uint32_t)(warp * 32) << 16), d);
    tmem_wait_ld();
    for (int j = 0; j < 64; ++j) D[tid * 64 + j] = __uint_as_float(d[j]);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 64);
}

// test 2: TS, A[128 rows][64 k=j] from TMEM, B = W (stored K-major as [n=j][k=i]) used as MN-major B'[k=j][n=i] => D[r][i] = sum_j A[r][j] W[j][i]
__global__ void __launch_bounds__(128) probe_ts_t(const float *A, const float *W, float *D)
{
    __shared__ __align__(1024) float sW[64 * 64];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tb_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(smem_u32(&tb_s), 128);
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_init_fence(); }
    for (int i = tid; i < 64 * 64; i += 128) {
        const int n = i / 64, k = i % 64;   // W[n=j][k=i]
        sW[(n >> 3) * 512 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3)] = W[i];
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tb_s;
    const uint32_t lane_base = tb + ((uint32_t)(warp * 32) << 16);
    uint32_t a[64];
    for (int k = 0; k < 64; ++k) a[k] = __float_as_uint(A[tid * 64 + k]);
    tmem_st64(lane_base, a);
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        const uint32_t idesc = umma_idesc_tf32(128, 64, 0, 1);   // B MN-major
        // as MN-major B': element (n=i, k=j) at (i%4)*4 + (j%8)*16 + (i/4)*128 + (j/8)*2048  => SBO(mn blocks)=128, LBO(k blocks)=2048
        for (int ks = 0; ks < 8; ++ks)
            umma_tf32_ts(tb + 64, tb + ks * 8, umma_desc(smem_u32(sW) + ks * 2048, 2048, 128), idesc, ks > 0);
        umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    uint32_t d[64];
    tmem_ld64(lane_base + 64, d);
    tmem_wait_ld();
    for (int j = 0; j < 64; ++j) D[tid * 64 + j] = __uint_as_float(d[j]);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 128);
}

int main()
{
    srand(2);
    auto rnd = [](std::vector<float> &v) { for (auto &x : v) x = (float)rand() / RAND_MAX * 2 - 1; };
    {
        std::vector<float> A(128 * 128), B(128 * 64), D(128 * 64);
        rnd(A); rnd(B);
        float *dA, *dB, *dD;
        cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
        cudaFuncSetAttribute(probe_ss, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 192 * 4);
        probe_ss<<<1, 128, 128 * 192 * 4>>>(dA, dB, dD);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("probe_ss CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double num = 0, den = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
            double s = 0; for (int r = 0; r < 128; ++r) s += (double)A[r * 128 + m] * B[r * 64 + n];
            num += (D[m * 64 + n] - s) * (D[m * 64 + n] - s); den += s * s;
        }
        printf("probe_ss (A,B MN-major, M=128,N=64,K=128, single tf32): rel err %.3e  D[0][0]=%f D[77][13]=%f\n", sqrt(num / den), D[0], D[77 * 64 + 13]);
    }
    {
        std::vector<float> A(128 * 64), W(64 * 64), D(128 * 64);
        rnd(A); rnd(W);
        float *dA, *dW, *dD;
        cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dD, D.size() * 4);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
        probe_ts_t<<<1, 128>>>(dA, dW, dD);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("probe_ts_t CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double num = 0, den = 0;
        for (int r = 0; r < 128; ++r) for (int i = 0; i < 64; ++i) {
            double s = 0; for (int j = 0; j < 64; ++j) s += (double)A[r * 64 + j] * W[j * 64 + i];
            num += (D[r * 64 + i] - s) * (D[r * 64 + i] - s); den += s * s;
        }
        printf("probe_ts_t (B = K-major W buffer read as MN-major W^T): rel err %.3e  D[0][0]=%f\n", sqrt(num / den), D[0]);
    }
    return 0;
}
