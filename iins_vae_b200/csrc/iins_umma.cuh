// Thin inline-PTX wrappers for the Blackwell (sm_100a) tensor-core path: tcgen05.mma with TMEM
// accumulators, mbarrier completion, TMEM loads, TMA bulk copies.  Descriptor formats follow the PTX ISA
// "tcgen05 matrix descriptors" (the same bit layout CUTLASS's cute/arch/mma_sm100_desc.hpp documents).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- shared-memory matrix descriptor (64 bit), no swizzle ("interleave") -------------------------------
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4   bits [46,48) version = 1 (Blackwell)     bits [61,64) layout = 0
// K-major operand:  core matrix = 8 rows x 16 bytes stored contiguously (row stride 16 B);
//                   SBO = byte distance between 8-row groups, LBO = byte distance between 16-byte K chunks.
// MN-major operand: core matrix = 8 K-rows x 16 bytes (8 MN elements of 2 B);
//                   SBO = distance between MN groups of 8 elements, LBO = distance between groups of 8 K-rows.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// ---- instruction descriptor (32 bit) for kind::f16 with BF16 inputs and FP32 accumulation ------------------
//   [4,6) D format 1=F32   [7,10) A format 1=BF16   [10,13) B format 1=BF16
//   [15] A major (0=K,1=MN)   [16] B major   [17,23) N>>3   [24,29) M>>4
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// one lane of a fully converged warp (warp-uniform code keeps the MMA descriptors in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// arrives on the mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void commit(uint32_t mbar_saddr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_saddr) : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t mbar_saddr, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_saddr), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_wait(uint32_t mbar_saddr, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}\n" ::"r"(mbar_saddr), "r"(parity)
        : "memory");
}

// same, for waits that are expected to be long (another warp role is working): each try suspends the thread for up to
// ~`ns` nanoseconds (it resumes as soon as the phase completes), so the spinning warps leave the issue slots to the others
__device__ __forceinline__ void mbar_wait_suspend(uint32_t mbar_saddr, uint32_t parity, uint32_t ns = 4000u) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}\n" ::"r"(mbar_saddr), "r"(parity), "r"(ns)
        : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t mbar_saddr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar_saddr) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t mbar_saddr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_saddr), "r"(bytes) : "memory");
}

// TMA bulk copy global -> shared (no tensor map: a contiguous byte range), completion on an mbarrier
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst_saddr, const void* src, uint32_t bytes, uint32_t mbar_saddr) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_saddr),
                 "l"(src), "r"(bytes), "r"(mbar_saddr)
                 : "memory");
}

// make generic-proxy smem writes (st.shared) visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM ---------------------------------------------------------------------------------------------
// executed by ONE full warp; writes the TMEM base address (lane<<16 | column) to *smem_slot
__device__ __forceinline__ void tmem_alloc(uint32_t smem_slot_saddr, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot_saddr), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// warp-collective: thread t of the warp receives 16 consecutive fp32 columns of TMEM lane (lane_base + t)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// non-blocking variant: issue several loads, then ONE tmem_ld_wait() before the registers are read
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- bf16 splitting -----------------------------------------------------------------------------------
// x = p0 + p1 + p2 (+ O(2^-24 x)), each piece a bf16 (round-to-nearest): three 8-bit mantissa slices.
__device__ __forceinline__ uint32_t bf16_rn_bits(float x) {          // round-to-nearest-even bf16, as the top 16 bits
    uint32_t u = __float_as_uint(x);
    uint32_t r = u + 0x7FFFu + ((u >> 16) & 1u);
    return r & 0xFFFF0000u;
}
__device__ __forceinline__ void split3(float x, uint32_t& p0, uint32_t& p1, uint32_t& p2) {
    p0 = bf16_rn_bits(x);
    float r1 = x - __uint_as_float(p0);
    p1 = bf16_rn_bits(r1);
    float r2 = r1 - __uint_as_float(p1);
    p2 = bf16_rn_bits(r2);
}
// Pair version on the hardware converter: cvt.rn.bf16x2.f32 packs two fp32 into {hi: b, lo: a} in one
// instruction; the residuals are formed exactly in fp32 (a - bf16(a) is representable).
__device__ __forceinline__ uint32_t cvt_bf16x2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));     // first source -> upper half
    return r;
}
__device__ __forceinline__ void split3_pair(float a, float b, uint32_t& w0, uint32_t& w1, uint32_t& w2) {
    w0 = cvt_bf16x2(a, b);
    float ra = a - __uint_as_float(w0 << 16), rb = b - __uint_as_float(w0 & 0xFFFF0000u);
    w1 = cvt_bf16x2(ra, rb);
    ra -= __uint_as_float(w1 << 16);
    rb -= __uint_as_float(w1 & 0xFFFF0000u);
    w2 = cvt_bf16x2(ra, rb);
}
// pack two bf16 (given as fp32 bit patterns with zero low halves) into one 32-bit word: lo = first element
__device__ __forceinline__ uint32_t pack2(uint32_t a_hi16, uint32_t b_hi16) { return (a_hi16 >> 16) | (b_hi16 & 0xFFFF0000u); }

}  // namespace umma
