// Probe 3: sweep smem-descriptor interpretations for SS-mode tcgen05.mma kind::tf32 with K-major / MN-major operands.
// Host pre-formats the shared-memory images; the kernel copies them verbatim and issues K/8 MMAs.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <vector>
#include <functional>
#include <cuda_runtime.h>
#include "../graph-neural-additive-networks---gnan_b200/csrc/tc_ptx.cuh"

struct Cfg { uint32_t idesc, a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep; int M, N, K; };

__global__ void __launch_bounds__(128) probe(const float *imgA, int nA, const float *imgB, int nB, float *D, Cfg c)
{
    extern __shared__ __align__(1024) float sm[];
    float *sA = sm, *sB = sm + nA;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tb_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(smem_u32(&tb_s), 128);
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_init_fence(); }
    for (int i = tid; i < nA; i += 128) sA[i] = imgA[i];
    for (int i = tid; i < nB; i += 128) sB[i] = imgB[i];
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tb_s;
    {   // poison D so that "MMA did nothing" is distinguishable from zeros
        uint32_t p[64];
        for (int j = 0; j < 64; ++j) p[j] = __float_as_uint(-777.0f);
        tmem_st64(tb + ((uint32_t)(warp * 32) << 16), p);
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        for (int ks = 0; ks < c.K / 8; ++ks)
            umma_tf32_ss(tb, umma_desc_kmajor(smem_u32(sA) + ks * c.a_kstep, c.a_lbo, c.a_sbo),
                         umma_desc_kmajor(smem_u32(sB) + ks * c.b_kstep, c.b_lbo, c.b_sbo), c.idesc, ks > 0);
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
    if (warp == 0) tmem_dealloc(tb, 128);
}

int main()
{
    const int M = 128, N = 64, K = 64;
    srand(3);
    std::vector<float> A(M * K), B(N * K);     // logical A[m][k], B[n][k];  D[m][n] = sum_k A[m][k] B[n][k]
    for (auto &x : A) x = (float)rand() / RAND_MAX * 2 - 1;
    for (auto &x : B) x = (float)rand() / RAND_MAX * 2 - 1;
    std::vector<double> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k]; ref[m * N + n] = s; }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, M * K * 4); cudaMalloc(&dB, N * K * 4); cudaMalloc(&dD, M * N * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (M + N) * K * 4);

    // layouts: offsets in floats
    auto kmaj = [&](int R) { return [=](int mn, int k) { return (mn / 8) * (K / 4) * 32 + (k / 4) * 32 + (mn % 8) * 4 + (k % 4); }; };   // LBO=128B (k chunks), SBO=(K/4)*128B
    auto mn_L1 = [&](int R) { return [=](int mn, int k) { return (mn / 4) * (K / 8) * 32 + (k / 8) * 32 + (k % 8) * 4 + (mn % 4); }; };  // k-blocks contiguous (128B), mn-blocks (K/8)*128B
    auto mn_L2 = [&](int R) { return [=](int mn, int k) { return (k / 8) * (R / 4) * 32 + (mn / 4) * 32 + (k % 8) * 4 + (mn % 4); }; };  // mn-blocks contiguous (128B), k-blocks (R/4)*128B
    struct Case { const char *name; int a_mn, b_mn; std::function<int(int,int)> la, lb; uint32_t a_lbo, a_sbo, a_ks, b_lbo, b_sbo, b_ks; };
    const uint32_t KK = (K / 4) * 128, K8 = (K / 8) * 128, MM = (M / 4) * 128, NN = (N / 4) * 128;
    std::vector<Case> cases = {
        {"SS K-major/K-major (sanity)", 0, 0, kmaj(M), kmaj(N), 128, KK, 256, 128, KK, 256},
        {"MN L1, desc(lbo=kblk,sbo=mnblk)", 1, 1, mn_L1(M), mn_L1(N), 128, K8, 128, 128, K8, 128},
        {"MN L1, desc(lbo=mnblk,sbo=kblk)", 1, 1, mn_L1(M), mn_L1(N), K8, 128, 128, K8, 128, 128},
        {"MN L2, desc(lbo=kblk,sbo=mnblk)", 1, 1, mn_L2(M), mn_L2(N), MM, 128, MM, NN, 128, NN},
        {"MN L2, desc(lbo=mnblk,sbo=kblk)", 1, 1, mn_L2(M), mn_L2(N), 128, MM, MM, 128, NN, NN},
        {"A K-major, B MN L1 (lbo=kblk,sbo=mnblk)", 0, 1, kmaj(M), mn_L1(N), 128, KK, 256, 128, K8, 128},
        {"A K-major, B MN L1 (lbo=mnblk,sbo=kblk)", 0, 1, kmaj(M), mn_L1(N), 128, KK, 256, K8, 128, 128},
        {"A K-major, B MN L2 (lbo=mnblk,sbo=kblk)", 0, 1, kmaj(M), mn_L2(N), 128, KK, 256, 128, NN, NN},
        {"A K-major, B MN L2 (lbo=kblk,sbo=mnblk)", 0, 1, kmaj(M), mn_L2(N), 128, KK, 256, NN, 128, NN},
    };
    for (auto &cs : cases) {
        std::vector<float> ia(M * K, 0.f), ib(N * K, 0.f);
        for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) ia[cs.la(m, k)] = A[m * K + k];
        for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) ib[cs.lb(n, k)] = B[n * K + k];
        cudaMemcpy(dA, ia.data(), ia.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, ib.data(), ib.size() * 4, cudaMemcpyHostToDevice);
        Cfg c{umma_idesc_tf32(M, N, cs.a_mn, cs.b_mn), cs.a_lbo, cs.a_sbo, cs.a_ks, cs.b_lbo, cs.b_sbo, cs.b_ks, M, N, K};
        probe<<<1, 128, (M + N) * K * 4>>>(dA, M * K, dB, N * K, dD, c);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-45s CUDA error: %s\n", cs.name, cudaGetErrorString(e)); return 1; }
        std::vector<float> D(M * N);
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double num = 0, den = 0;
        for (int i = 0; i < M * N; ++i) { num += (D[i] - ref[i]) * (D[i] - ref[i]); den += ref[i] * ref[i]; }
        printf("%-45s rel err %.3e   D[0]=%9.4f ref %9.4f   D[77*64+13]=%9.4f ref %9.4f\n", cs.name, sqrt(num / den), D[0], ref[0], D[77 * 64 + 13], ref[77 * 64 + 13]);
    }
    return 0;
}
