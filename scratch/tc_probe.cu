// Standalone probe: tcgen05.mma kind::tf32, A from TMEM (TS), B from smem (K-major, no swizzle), 3xTF32 split.
// D[128x64] = A[128x64] * B^T  with B given as [n=64][k=64] (torch Linear weight layout).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
    return d;                // layout_type = 0 (no swizzle), base_offset = 0
}

__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accum)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accum) : "memory");
}

#define ST64(taddr, v) \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x64.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63,%64};" \
        :: "r"(taddr), \
        "r"(v[0]),"r"(v[1]),"r"(v[2]),"r"(v[3]),"r"(v[4]),"r"(v[5]),"r"(v[6]),"r"(v[7]),"r"(v[8]),"r"(v[9]),"r"(v[10]),"r"(v[11]),"r"(v[12]),"r"(v[13]),"r"(v[14]),"r"(v[15]), \
        "r"(v[16]),"r"(v[17]),"r"(v[18]),"r"(v[19]),"r"(v[20]),"r"(v[21]),"r"(v[22]),"r"(v[23]),"r"(v[24]),"r"(v[25]),"r"(v[26]),"r"(v[27]),"r"(v[28]),"r"(v[29]),"r"(v[30]),"r"(v[31]), \
        "r"(v[32]),"r"(v[33]),"r"(v[34]),"r"(v[35]),"r"(v[36]),"r"(v[37]),"r"(v[38]),"r"(v[39]),"r"(v[40]),"r"(v[41]),"r"(v[42]),"r"(v[43]),"r"(v[44]),"r"(v[45]),"r"(v[46]),"r"(v[47]), \
        "r"(v[48]),"r"(v[49]),"r"(v[50]),"r"(v[51]),"r"(v[52]),"r"(v[53]),"r"(v[54]),"r"(v[55]),"r"(v[56]),"r"(v[57]),"r"(v[58]),"r"(v[59]),"r"(v[60]),"r"(v[61]),"r"(v[62]),"r"(v[63]) : "memory")

#define LD64(taddr, v) \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];" \
        : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]), \
        "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]), \
        "=r"(v[32]),"=r"(v[33]),"=r"(v[34]),"=r"(v[35]),"=r"(v[36]),"=r"(v[37]),"=r"(v[38]),"=r"(v[39]),"=r"(v[40]),"=r"(v[41]),"=r"(v[42]),"=r"(v[43]),"=r"(v[44]),"=r"(v[45]),"=r"(v[46]),"=r"(v[47]), \
        "=r"(v[48]),"=r"(v[49]),"=r"(v[50]),"=r"(v[51]),"=r"(v[52]),"=r"(v[53]),"=r"(v[54]),"=r"(v[55]),"=r"(v[56]),"=r"(v[57]),"=r"(v[58]),"=r"(v[59]),"=r"(v[60]),"=r"(v[61]),"=r"(v[62]),"=r"(v[63]) \
        : "r"(taddr) : "memory")

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" :: "r"(bar), "r"(parity) : "memory");
}

// mode: 0 = 3xTF32, 1 = single tf32
__global__ void __launch_bounds__(128) probe(const float *A, const float *B, float *D, int mode, long long *cycles)
{
    __shared__ __align__(1024) float sBhi[64 * 64];
    __shared__ __align__(1024) float sBlo[64 * 64];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    // B -> canonical K-major no-swizzle: (n,k) at (n/8)*2048 + (k/4)*128 + (n%8)*16 + (k%4)*4 bytes
    for (int i = tid; i < 64 * 64; i += 128) {
        const int n = i / 64, k = i % 64;
        const float v = B[i];
        const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        const int off = (n / 8) * 512 + (k / 4) * 32 + (n % 8) * 4 + (k % 4);
        sBhi[off] = hi;
        sBlo[off] = v - hi;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base_s;
    const uint32_t lane_base = tb + ((uint32_t)(warp * 32) << 16);
    // A row -> TMEM: hi in cols [0,64), lo in [64,128)
    uint32_t hi[64], lo[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) {
        const float v = A[tid * 64 + k];
        const uint32_t h = __float_as_uint(v) & 0xFFFFE000u;
        hi[k] = h;
        lo[k] = __float_as_uint(v - __uint_as_float(h));
    }
    ST64(lane_base + 0, hi);
    ST64(lane_base + 64, lo);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    long long t0 = clock64();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = make_idesc_tf32(128, 64);
        const uint32_t d_tmem = tb + 128;
        for (int ks = 0; ks < 8; ++ks)
            mma_tf32_ts(d_tmem, tb + ks * 8, make_desc_kmajor(smem_u32(sBhi) + ks * 256, 128, 2048), idesc, ks > 0);
        if (mode == 0) {
            for (int ks = 0; ks < 8; ++ks)
                mma_tf32_ts(d_tmem, tb + 64 + ks * 8, make_desc_kmajor(smem_u32(sBhi) + ks * 256, 128, 2048), idesc, 1);
            for (int ks = 0; ks < 8; ++ks)
                mma_tf32_ts(d_tmem, tb + ks * 8, make_desc_kmajor(smem_u32(sBlo) + ks * 256, 128, 2048), idesc, 1);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t d[64];
    LD64(lane_base + 128, d);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 64; ++j) D[tid * 64 + j] = __uint_as_float(d[j]);
    if (tid == 0) *cycles = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(256));
}

int main()
{
    std::vector<float> A(128 * 64), B(64 * 64), D(128 * 64);
    srand(1);
    for (auto &v : A) v = fmaxf(0.f, (float)rand() / RAND_MAX * 2 - 0.7f);
    for (auto &v : B) v = ((float)rand() / RAND_MAX * 2 - 1) * 0.3f;
    float *dA, *dB, *dD; long long *dc;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dc, 8);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 2; ++mode) {
        cudaMemset(dD, 0, D.size() * 4);
        probe<<<1, 128>>>(dA, dB, dD, mode, dc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d CUDA error: %s\n", mode, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        long long cyc; cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
        double num = 0, den = 0, maxe = 0;
        for (int i = 0; i < 128; ++i)
            for (int j = 0; j < 64; ++j) {
                double s = 0;
                for (int k = 0; k < 64; ++k) s += (double)A[i * 64 + k] * (double)B[j * 64 + k];
                double e2 = D[i * 64 + j] - s;
                num += e2 * e2; den += s * s; maxe = fmax(maxe, fabs(e2));
            }
        printf("mode %d (%s): rel err %.3e  max abs err %.3e  cycles(issue->commit) %lld  D[0][0]=%f D[5][7]=%f\n", mode,
               mode == 0 ? "3xTF32" : "1xTF32", sqrt(num / den), maxe, cyc, D[0], D[5 * 64 + 7]);
    }
    return 0;
}
