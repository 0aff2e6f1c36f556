// tcgen05 / TMEM / mbarrier PTX wrappers for sm_100a (hand-written; encodings follow the PTX ISA as used by CUTLASS's
// cute/arch/mma_sm100_desc.hpp and copy_sm100.hpp, verified on a B200 by scratch/tc_probe.cu).
#pragma once
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, K-major, no swizzle: 8-row x 16-byte core matrices; LBO = byte distance between core
// matrices adjacent along K, SBO = between 8-row groups along M/N.
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100)
    return d;
}

// instruction descriptor for kind::tf32, fp32 accumulate. a_mn / b_mn: 1 = MN-major operand (0 = K-major).
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, int a_mn = 0, int b_mn = 0)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accum)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accum) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum) : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tGNAN_WAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra GNAN_DONE_%=;\n\tbra GNAN_WAIT_%=;\n\tGNAN_DONE_%=:\n\t}\n"
                 :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}

// fp32 -> (hi, lo) with hi = round-to-nearest tf32 and lo = v - hi (exact in fp32; the MMA truncates it to tf32)
__device__ __forceinline__ void split_tf32(float v, uint32_t &hi, uint32_t &lo)
{
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    hi = h;
    lo = __float_as_uint(v - __uint_as_float(h));
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" :: "r"(taddr), "r"(v[0]),"r"(v[1]),"r"(v[2]),"r"(v[3]),"r"(v[4]),"r"(v[5]),"r"(v[6]),"r"(v[7]),"r"(v[8]),"r"(v[9]),"r"(v[10]),"r"(v[11]),"r"(v[12]),"r"(v[13]),"r"(v[14]),"r"(v[15]) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]) : "r"(taddr) : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" :: "r"(taddr), "r"(v[0]),"r"(v[1]),"r"(v[2]),"r"(v[3]),"r"(v[4]),"r"(v[5]),"r"(v[6]),"r"(v[7]),"r"(v[8]),"r"(v[9]),"r"(v[10]),"r"(v[11]),"r"(v[12]),"r"(v[13]),"r"(v[14]),"r"(v[15]),"r"(v[16]),"r"(v[17]),"r"(v[18]),"r"(v[19]),"r"(v[20]),"r"(v[21]),"r"(v[22]),"r"(v[23]),"r"(v[24]),"r"(v[25]),"r"(v[26]),"r"(v[27]),"r"(v[28]),"r"(v[29]),"r"(v[30]),"r"(v[31]) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),"=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]) : "r"(taddr) : "memory");
}

__device__ __forceinline__ void tmem_st64(uint32_t taddr, const uint32_t (&v)[64])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x64.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63,%64};" :: "r"(taddr), "r"(v[0]),"r"(v[1]),"r"(v[2]),"r"(v[3]),"r"(v[4]),"r"(v[5]),"r"(v[6]),"r"(v[7]),"r"(v[8]),"r"(v[9]),"r"(v[10]),"r"(v[11]),"r"(v[12]),"r"(v[13]),"r"(v[14]),"r"(v[15]),"r"(v[16]),"r"(v[17]),"r"(v[18]),"r"(v[19]),"r"(v[20]),"r"(v[21]),"r"(v[22]),"r"(v[23]),"r"(v[24]),"r"(v[25]),"r"(v[26]),"r"(v[27]),"r"(v[28]),"r"(v[29]),"r"(v[30]),"r"(v[31]),"r"(v[32]),"r"(v[33]),"r"(v[34]),"r"(v[35]),"r"(v[36]),"r"(v[37]),"r"(v[38]),"r"(v[39]),"r"(v[40]),"r"(v[41]),"r"(v[42]),"r"(v[43]),"r"(v[44]),"r"(v[45]),"r"(v[46]),"r"(v[47]),"r"(v[48]),"r"(v[49]),"r"(v[50]),"r"(v[51]),"r"(v[52]),"r"(v[53]),"r"(v[54]),"r"(v[55]),"r"(v[56]),"r"(v[57]),"r"(v[58]),"r"(v[59]),"r"(v[60]),"r"(v[61]),"r"(v[62]),"r"(v[63]) : "memory");
}

__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[64])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];" : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),"=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]),"=r"(v[32]),"=r"(v[33]),"=r"(v[34]),"=r"(v[35]),"=r"(v[36]),"=r"(v[37]),"=r"(v[38]),"=r"(v[39]),"=r"(v[40]),"=r"(v[41]),"=r"(v[42]),"=r"(v[43]),"=r"(v[44]),"=r"(v[45]),"=r"(v[46]),"=r"(v[47]),"=r"(v[48]),"=r"(v[49]),"=r"(v[50]),"=r"(v[51]),"=r"(v[52]),"=r"(v[53]),"=r"(v[54]),"=r"(v[55]),"=r"(v[56]),"=r"(v[57]),"=r"(v[58]),"=r"(v[59]),"=r"(v[60]),"=r"(v[61]),"=r"(v[62]),"=r"(v[63]) : "r"(taddr) : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "r"(taddr), "r"(v[0]),"r"(v[1]),"r"(v[2]),"r"(v[3]),"r"(v[4]),"r"(v[5]),"r"(v[6]),"r"(v[7]) : "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]) : "r"(taddr) : "memory");
}


// a global load the compiler may neither sink to its use nor rematerialise (software prefetch into a register)
__device__ __forceinline__ float ldg_prefetch(const float *p)
{
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
