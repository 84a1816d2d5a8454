// Fused residual trunk, forward (see iins_trunk.h).
//
// Layout of the resident operand in shared memory, per bf16 piece:   A[chunk of 8 channels (8)][sample (16)][row (10)][16 B]
// where the 10 rows of a sample are  [x(1) | x(0) .. x(7) | x(6)]:  the two reflected edge rows of ReflectionPad1d(1)
// (models.py:993,998) are MATERIALISED, so the three taps of the k3 convolution are three windows of the same array
// shifted by one row.  In the tcgen05 K-major no-swizzle operand format a core matrix is 8 rows x 16 bytes with the rows 16
// bytes apart: a sample's 8 output positions are exactly one core matrix, the stride between core matrices (SBO) is the 160
// bytes of a 10-row sample, the stride between 8-channel chunks (LBO) is 2560 bytes, and tap t is the SAME descriptor with
// its start address advanced by t * 16 bytes.  No im2col copy exists anywhere: the MMA reads the activation tile in place.
//
// Roles (320 threads, 2 CTAs / SM): warps 0-7 = epilogue (TMEM -> registers, bias, InstanceNorm / AdaIN over the 8 lanes of
// a sample, ReLU / residual, fp32 y / x-hat / rstd to HBM for the backward pass, and the NEXT convolution's operand written
// back into A as bf16 pieces), warp 8 = MMA issue (one elected lane, tcgen05.mma M=128, N = 64 x stacked pieces, K=16),
// warp 9 = weight stream: the packed weight slices (one per (tap, 16 channels)) of ALL convolutions of the chain flow
// through an 8-slot ring by tensor-map TMA (cp.async.bulk.tensor), running ahead across convolution boundaries.
// Synchronisation is mbarrier-only: w_full / w_empty per ring slot, acc_full (tcgen05.commit -> epilogue), a_ready
// (256 epilogue arrivals -> MMA warp).
#include "iins_tc.cuh"
#include "iins_trunk.h"
#include <cuda.h>
#include <stdio.h>

namespace {

constexpr int TR_SAMPLES = 16;                       // samples per tile (128 GEMM rows)
constexpr int TR_ROWS = TR_SAMPLES * 10;             // operand rows incl. the edge rows
constexpr uint32_t TR_LBO = TR_ROWS * 16 + 16;       // bytes between 8-channel chunks (+16: the epilogue's 8-byte stores of the
                                                     // 8 chunks of a row then fall into different banks)
constexpr uint32_t TR_SBO = 160;                     // bytes between 8-row groups (= samples)
constexpr uint32_t TR_APIECE = 8 * TR_LBO;           // one piece of the operand: 20608 B
constexpr int TR_NSLOT = 8;                          // weight ring slots
constexpr int TR_KS = 12;                            // k-steps of 16 per convolution: 3 taps x 4
constexpr int TR_STG_LD = 68;                        // floats per row of the accumulator staging tile (64 + 4: conflict-free)
constexpr uint32_t TR_STG_BYTES = 128 * TR_STG_LD * 4;

template <int PIECES>
struct TrunkSmem {
    static constexpr uint32_t SLICE = PIECES * 64 * 16 * 2;       // one (tap, 16-channel) weight slice: [2 chunks][PIECES*64 rows][16 B]
    static constexpr uint32_t A_BYTES = PIECES * TR_APIECE;
    static constexpr uint32_t RING = TR_NSLOT * SLICE;
    // the fp32 staging tile of the epilogue aliases the operand (dead once the MMAs of a convolution have completed) when
    // the operand is large enough (3 pieces); the 1-piece operand is smaller than the tile, which then gets its own space
    static constexpr bool STG_ALIAS = A_BYTES >= TR_STG_BYTES;
    static constexpr uint32_t TOTAL = A_BYTES + RING + (STG_ALIAS ? 0 : TR_STG_BYTES);
};

__device__ __forceinline__ void tma_tensor_2d(uint32_t dst_saddr, const CUtensorMap* map, int c0, int c1, uint32_t mbar_saddr) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst_saddr), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(mbar_saddr)
                 : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }     // the 8 epilogue warps

// ---- epilogue plumbing shared by the forward and the backward kernel -------------------------------------------------
// The accumulator row of TMEM lane r can only be read by thread (r & 31) of a warp with (warp & 3) == r >> 5, i.e. one
// thread sees ONE position of a sample -- but InstanceNorm and its backward reduce over the 8 positions.  Instead of
// 6 shuffles per value, the tile is transposed through shared memory once: every thread stages its row (2 x 16 columns),
// one named barrier, then thread t owns sample (t >> 4), channels 4 * (t & 15) .. +3 at ALL 8 positions (32 values in
// registers): the statistics are plain in-thread sums, the per-(sample, channel) operands (bias, AdaIN weight / bias, rstd)
// are loaded once instead of once per position, and every global access is a 16-byte piece of a 256-byte contiguous row.
template <int PIECES>
__device__ __forceinline__ void stage_accumulator(uint32_t tmem, int warp, int lane, float* stg) {
    const int q = warp & 3, hf = warp >> 2;
    const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
    float* dst = stg + (q * 32 + lane) * TR_STG_LD + hf * 32;
#pragma unroll
    for (int c0 = 0; c0 < 32; c0 += 16) {
        float v[16];
        iins_tmem_chunk16<64, PIECES>(tl, hf * 32 + c0, v);
#pragma unroll
        for (int k = 0; k < 4; ++k) *reinterpret_cast<float4*>(dst + c0 + 4 * k) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    }
}
__device__ __forceinline__ void load_columns(const float* stg, int s, int cg, float4* x) {
#pragma unroll
    for (int l = 0; l < 8; ++l) x[l] = *reinterpret_cast<const float4*>(stg + (s * 8 + l) * TR_STG_LD + 4 * cg);
}
// 4 channels x 8 positions -> bf16 pieces in the operand tile: rows 1..8 of the sample; the two edge rows are the reflected
// copies (forward: ReflectionPad1d(1), padded row 0 = x(1), padded row 9 = x(6)) or zeros (backward).  All ten rows are
// written every time: the staging tile of the epilogue may alias them.
template <int PIECES, bool REFLECT>
__device__ __forceinline__ void write_operand_columns(unsigned char* sA, int s, int cg, const float4* x) {
    unsigned char* base = sA + (cg >> 1) * TR_LBO + (s * 10) * 16 + (cg & 1) * 8;
#pragma unroll
    for (int l = 0; l < 8; ++l) iins_store4_split(x[l], base + (l + 1) * 16, TR_APIECE, PIECES);
    if (REFLECT) {
        iins_store4_split(x[1], base, TR_APIECE, PIECES);
        iins_store4_split(x[6], base + 9 * 16, TR_APIECE, PIECES);
    } else {
#pragma unroll
        for (int pc = 0; pc < PIECES; ++pc) {
            *reinterpret_cast<uint2*>(base + pc * TR_APIECE) = make_uint2(0u, 0u);
            *reinterpret_cast<uint2*>(base + pc * TR_APIECE + 9 * 16) = make_uint2(0u, 0u);
        }
    }
}
#define IINS_F4_OP(dst, a, op, b) do { (dst).x = (a).x op (b).x; (dst).y = (a).y op (b).y; (dst).z = (a).z op (b).z; (dst).w = (a).w op (b).w; } while (0)

// ======================================================================================================= forward
template <int PIECES, bool ADAIN, bool TMAP>
__global__ void __launch_bounds__(320, 2) iins_trunk_fwd_kernel(const IinsTrunkFwdParams p, const __grid_constant__ CUtensorMap wmap) {
    using SM = TrunkSmem<PIECES>;
    constexpr uint32_t SLICE = SM::SLICE;
    constexpr int TCOLS = IinsTmemCols<64, PIECES>::value;
    extern __shared__ __align__(1024) unsigned char dsm[];
    __shared__ __align__(8) unsigned long long w_full[TR_NSLOT];
    __shared__ __align__(8) unsigned long long w_empty[TR_NSLOT];
    __shared__ __align__(8) unsigned long long acc_full;
    __shared__ __align__(8) unsigned long long a_ready;
    __shared__ uint32_t tmem_slot;
    unsigned char* sA = dsm;
    unsigned char* sW = dsm + SM::A_BYTES;
    float* stg = reinterpret_cast<float*>(SM::STG_ALIAS ? dsm : dsm + SM::A_BYTES + SM::RING);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = (p.B + TR_SAMPLES - 1) / TR_SAMPLES;

    IINS_PERSISTENT_PDL_TRIGGER();
    if (tid == 0) {
        for (int i = 0; i < TR_NSLOT; ++i) {
            umma::mbar_init(umma::smem_u32(&w_full[i]), 1);
            umma::mbar_init(umma::smem_u32(&w_empty[i]), 1);
        }
        umma::mbar_init(umma::smem_u32(&acc_full), 1);
        umma::mbar_init(umma::smem_u32(&a_ready), 256);
        umma::fence_mbar_init();
        if (TMAP) prefetch_tensormap(&wmap);
    }
    if (warp == 8) umma::tmem_alloc(umma::smem_u32(&tmem_slot), TCOLS);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    iins_pdl_wait();                       // everything above overlapped the tail of the previous kernel in the stream

    if (warp == 9) {
        // ------------------------------------------------------------------ weight stream (runs ahead of the MMA warp)
        uint32_t i = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            for (int c = 0; c < p.nconv; ++c) {
                for (int ks = 0; ks < TR_KS; ++ks, ++i) {
                    const uint32_t slot = i % TR_NSLOT;
                    if (i >= TR_NSLOT) umma::mbar_wait(umma::smem_u32(&w_empty[slot]), ((i / TR_NSLOT) - 1) & 1);
                    if (umma::elect_one()) {
                        const uint32_t bar = umma::smem_u32(&w_full[slot]);
                        umma::mbar_arrive_expect_tx(bar, SLICE);
                        const int slice = c * TR_KS + ks;
                        if (TMAP) tma_tensor_2d(umma::smem_u32(sW + slot * SLICE), &wmap, 0, slice * (int)(SLICE / 512), bar);
                        else umma::tma_bulk_g2s(umma::smem_u32(sW + slot * SLICE), reinterpret_cast<const unsigned char*>(p.wpack) + (size_t)slice * SLICE, SLICE, bar);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 8) {
        // ------------------------------------------------------------------ MMA issue (warp-uniform code)
        uint32_t i = 0, j = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            for (int c = 0; c < p.nconv; ++c, ++j) {
                umma::mbar_wait(umma::smem_u32(&a_ready), j & 1);          // operand of convolution c is in A, TMEM is free
                umma::tc_fence_after();
                const bool leader = umma::elect_one();
                for (int ks = 0; ks < TR_KS; ++ks, ++i) {
                    const uint32_t slot = i % TR_NSLOT;
                    umma::mbar_wait(umma::smem_u32(&w_full[slot]), (i / TR_NSLOT) & 1);
                    umma::tc_fence_after();
                    const int tap = ks >> 2, kk = ks & 3;
                    const uint64_t ad = umma::make_desc(umma::smem_u32(sA) + tap * 16 + kk * 2 * TR_LBO, TR_LBO, TR_SBO);
                    const uint64_t bd = umma::make_desc(umma::smem_u32(sW + slot * SLICE), PIECES * 64 * 16, 128);
                    iins_issue_kstep<64, PIECES, 0, 0>(tmem, ad, bd, TR_APIECE >> 4, leader, ks > 0 ? 1u : 0u);
                    if (leader) umma::commit(umma::smem_u32(&w_empty[slot]));
                    __syncwarp();
                }
                if (leader) umma::commit(umma::smem_u32(&acc_full));
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps
        const int s = tid >> 4, cg = tid & 15;                              // after the transpose: sample, 4-channel group
        uint32_t j = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int gb = tile * TR_SAMPLES + s;
            const bool ok = gb < p.B;
            const long sb = ok ? gb : 0;
            const long o0 = sb * 512 + 4 * cg;                              // element (sample, position 0, channel 4 cg) of a (B, 8, 64) tensor
            float4 x[8];
            // ---- stage 0: the trunk input (fp32, channels-last) -> bf16 pieces in A
#pragma unroll
            for (int l = 0; l < 8; ++l) x[l] = ok ? __ldg(reinterpret_cast<const float4*>(p.x + o0 + l * 64)) : make_float4(0.f, 0.f, 0.f, 0.f);
            write_operand_columns<PIECES, true>(sA, s, cg, x);
            umma::fence_async_smem();
            umma::mbar_arrive(umma::smem_u32(&a_ready));
            for (int c = 0; c < p.nconv; ++c, ++j) {
                const IinsTrunkLayer& ly = p.layer[c];
                const bool second = (c & 1) != 0;                           // second convolution of a block: no ReLU, + skip
                umma::mbar_wait(umma::smem_u32(&acc_full), j & 1);
                umma::tc_fence_after();
                stage_accumulator<PIECES>(tmem, warp, lane, stg);
                umma::tc_fence_before();
                epi_bar();
                load_columns(stg, s, cg, x);
                epi_bar();                                                  // staging (aliasing A) may be overwritten from here on
                // (the bias is a parameter tensor inside the caller's flat buffer: 4-byte aligned only)
                const float4 b4 = make_float4(__ldg(ly.bias + 4 * cg), __ldg(ly.bias + 4 * cg + 1), __ldg(ly.bias + 4 * cg + 2), __ldg(ly.bias + 4 * cg + 3));
                // InstanceNorm over the 8 positions: biased variance, eps inside the square root (models.py:152, 1072)
                float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int l = 0; l < 8; ++l) { IINS_F4_OP(x[l], x[l], +, b4); IINS_F4_OP(sum, sum, +, x[l]); }
                const float4 mean = make_float4(sum.x * 0.125f, sum.y * 0.125f, sum.z * 0.125f, sum.w * 0.125f);
                float4 sq = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int l = 0; l < 8; ++l) {
                    IINS_F4_OP(x[l], x[l], -, mean);
                    sq.x = fmaf(x[l].x, x[l].x, sq.x); sq.y = fmaf(x[l].y, x[l].y, sq.y); sq.z = fmaf(x[l].z, x[l].z, sq.z); sq.w = fmaf(x[l].w, x[l].w, sq.w);
                }
                auto rstd_of = [](float ssq) {
                    const float vpe = fmaf(ssq, 0.125f, IINS_EPS);
                    const float r = rsqrtf(vpe);
                    return r * fmaf(-0.5f * vpe, r * r, 1.5f);              // one Newton step: full fp32 accuracy
                };
                const float4 rs = make_float4(rstd_of(sq.x), rstd_of(sq.y), rstd_of(sq.z), rstd_of(sq.w));
                if (ok) *reinterpret_cast<float4*>(ly.rstd + sb * 64 + 4 * cg) = rs;
#pragma unroll
                for (int l = 0; l < 8; ++l) {
                    IINS_F4_OP(x[l], x[l], *, rs);
                    if (ok) *reinterpret_cast<float4*>(ly.xhat + o0 + l * 64) = x[l];
                }
                if (ADAIN) {
                    const float* ab = p.adain + sb * p.adain_ld + 4 * cg;
                    const float4 w4 = __ldg(reinterpret_cast<const float4*>(ab + ly.adain_off_w));
                    const float4 a4 = __ldg(reinterpret_cast<const float4*>(ab + ly.adain_off_b));
#pragma unroll
                    for (int l = 0; l < 8; ++l) {
                        x[l].x = fmaf(x[l].x, w4.x, a4.x); x[l].y = fmaf(x[l].y, w4.y, a4.y); x[l].z = fmaf(x[l].z, w4.z, a4.z); x[l].w = fmaf(x[l].w, w4.w, a4.w);
                    }
                }
                if (!second) {
#pragma unroll
                    for (int l = 0; l < 8; ++l) { x[l].x = fmaxf(x[l].x, 0.f); x[l].y = fmaxf(x[l].y, 0.f); x[l].z = fmaxf(x[l].z, 0.f); x[l].w = fmaxf(x[l].w, 0.f); }
                } else if (ok) {
                    // plain (coherent) loads: for c >= 3 this thread wrote these very values earlier in the kernel
                    const float* skip = (c == 1 ? p.x : p.layer[c - 2].y) + o0;
#pragma unroll
                    for (int l = 0; l < 8; ++l) { const float4 h4 = *reinterpret_cast<const float4*>(skip + l * 64); IINS_F4_OP(x[l], x[l], +, h4); }
                }
                if (ok) {
#pragma unroll
                    for (int l = 0; l < 8; ++l) *reinterpret_cast<float4*>(ly.y + o0 + l * 64) = x[l];
                }
                if (c + 1 < p.nconv) {                                      // the next convolution reads it in place
                    write_operand_columns<PIECES, true>(sA, s, cg, x);
                    umma::fence_async_smem();
                    umma::mbar_arrive(umma::smem_u32(&a_ready));
                }
            }
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 8) umma::tmem_dealloc(tmem, TCOLS);
}

// ======================================================================================================= backward
// Data gradient of a k3 reflect-pad convolution as shifted windows.  With dz padded by ONE ZERO ROW on each side,
//   dx(pos) = sum_t W_t^T dz(pos + 1 - t)            -> tap t reads the window starting (2 - t) rows into the padded array,
// plus the two terms that the reflected edge rows of the forward pass add (padded row 0 was x(1), padded row 9 was x(6)):
//   dx(1) += W_0^T dz(0),        dx(6) += W_2^T dz(7).
// Those are two more windows of an array that holds ONLY dz(0) (in padded row 1) and dz(7) (in padded row 8): the same
// shared-memory tile after the epilogue warps have zeroed its six middle rows.  So a convolution is 12 k-steps on the tile,
// a barrier round trip in which rows 2..7 are cleared, and 8 more k-steps (tap-0 weights on the window at shift 0, tap-2
// weights on the window at shift 2) into the same TMEM accumulator; the weight ring streams 20 slices per convolution.
template <int PIECES, bool ADAIN, bool TMAP>
__global__ void __launch_bounds__(320, 2) iins_trunk_bwd_kernel(const IinsTrunkBwdParams p, const __grid_constant__ CUtensorMap wmap) {
    using SM = TrunkSmem<PIECES>;
    constexpr uint32_t SLICE = SM::SLICE;
    constexpr int TCOLS = IinsTmemCols<64, PIECES>::value;
    constexpr int NSL = 20;                                        // slices per convolution through the ring
    extern __shared__ __align__(1024) unsigned char dsm[];
    __shared__ __align__(8) unsigned long long w_full[TR_NSLOT];
    __shared__ __align__(8) unsigned long long w_empty[TR_NSLOT];
    __shared__ __align__(8) unsigned long long acc_full, main_done, a_ready, z_ready;
    __shared__ uint32_t tmem_slot;
    unsigned char* sA = dsm;
    unsigned char* sW = dsm + SM::A_BYTES;
    float* stg = reinterpret_cast<float*>(SM::STG_ALIAS ? dsm : dsm + SM::A_BYTES + SM::RING);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = (p.B + TR_SAMPLES - 1) / TR_SAMPLES;

    IINS_PERSISTENT_PDL_TRIGGER();
    if (tid == 0) {
        for (int i = 0; i < TR_NSLOT; ++i) {
            umma::mbar_init(umma::smem_u32(&w_full[i]), 1);
            umma::mbar_init(umma::smem_u32(&w_empty[i]), 1);
        }
        umma::mbar_init(umma::smem_u32(&acc_full), 1);
        umma::mbar_init(umma::smem_u32(&main_done), 1);
        umma::mbar_init(umma::smem_u32(&a_ready), 256);
        umma::mbar_init(umma::smem_u32(&z_ready), 256);
        umma::fence_mbar_init();
        if (TMAP) prefetch_tensormap(&wmap);
    }
    if (warp == 8) umma::tmem_alloc(umma::smem_u32(&tmem_slot), TCOLS);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    iins_pdl_wait();

    // slice sequence of one convolution: the 12 (tap, 16-channel) slices, then tap 0 and tap 2 again for the edge terms
    auto slice_of = [](int q) { return q < 12 ? q : (q < 16 ? q - 12 : q - 8); };

    if (warp == 9) {
        uint32_t i = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            for (int c = p.nconv - 1; c >= 0; --c) {
                for (int qq = 0; qq < NSL; ++qq, ++i) {
                    const uint32_t slot = i % TR_NSLOT;
                    if (i >= TR_NSLOT) umma::mbar_wait(umma::smem_u32(&w_empty[slot]), ((i / TR_NSLOT) - 1) & 1);
                    if (umma::elect_one()) {
                        const uint32_t bar = umma::smem_u32(&w_full[slot]);
                        umma::mbar_arrive_expect_tx(bar, SLICE);
                        const int slice = (p.nconv - 1 - c) * TR_KS + slice_of(qq);       // packs are in descending conv order
                        if (TMAP) tma_tensor_2d(umma::smem_u32(sW + slot * SLICE), &wmap, 0, slice * (int)(SLICE / 512), bar);
                        else umma::tma_bulk_g2s(umma::smem_u32(sW + slot * SLICE), reinterpret_cast<const unsigned char*>(p.wpack) + (size_t)slice * SLICE, SLICE, bar);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 8) {
        uint32_t i = 0, j = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            for (int c = p.nconv - 1; c >= 0; --c, ++j) {
                umma::mbar_wait(umma::smem_u32(&a_ready), j & 1);
                umma::tc_fence_after();
                const bool leader = umma::elect_one();
                for (int qq = 0; qq < NSL; ++qq, ++i) {
                    if (qq == 12) {
                        if (leader) umma::commit(umma::smem_u32(&main_done));     // the tile may be edited once these MMAs have read it
                        __syncwarp();
                        umma::mbar_wait(umma::smem_u32(&z_ready), j & 1);          // rows 2..7 cleared
                        umma::tc_fence_after();
                    }
                    const uint32_t slot = i % TR_NSLOT;
                    umma::mbar_wait(umma::smem_u32(&w_full[slot]), (i / TR_NSLOT) & 1);
                    umma::tc_fence_after();
                    const int ks = slice_of(qq);
                    const int tap = ks >> 2, kk = ks & 3;
                    const int shift = qq < 12 ? 2 - tap : tap;                     // main windows: 2 - t rows; edge windows: 0 (tap 0) / 2 (tap 2)
                    const uint64_t ad = umma::make_desc(umma::smem_u32(sA) + shift * 16 + kk * 2 * TR_LBO, TR_LBO, TR_SBO);
                    const uint64_t bd = umma::make_desc(umma::smem_u32(sW + slot * SLICE), PIECES * 64 * 16, 128);
                    iins_issue_kstep<64, PIECES, 0, 0>(tmem, ad, bd, TR_APIECE >> 4, leader, qq > 0 ? 1u : 0u);
                    if (leader) umma::commit(umma::smem_u32(&w_empty[slot]));
                    __syncwarp();
                }
                if (leader) umma::commit(umma::smem_u32(&acc_full));
                __syncwarp();
            }
        }
    } else {
        const int s = tid >> 4, cg = tid & 15;
        // InstanceNorm / AdaIN backward of layer `ly` applied to the gradient x (w.r.t. that layer's OUTPUT; 8 positions x 4
        // channels):  dz = rstd * w * (g - mean_l(g) - xhat * mean_l(g * xhat)),  g = x masked by the layer's ReLU;
        // AdaIN: d bias = sum_l g, d weight = sum_l g * xhat.  x is replaced by dz, which also goes to HBM (weight gradients).
        auto norm_backward = [&](const IinsTrunkBwdLayer& ly, bool relu, float4* x, long o0, long sb, bool ok, bool adain) {
            const float4 r4 = __ldg(reinterpret_cast<const float4*>(ly.rstd + sb * 64 + 4 * cg));
            float4 w4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (adain) {
                w4 = __ldg(reinterpret_cast<const float4*>(p.adain + sb * p.adain_ld + ly.adain_off_w + 4 * cg));
                b4 = __ldg(reinterpret_cast<const float4*>(p.adain + sb * p.adain_ld + ly.adain_off_b + 4 * cg));
            }
            float4 xh[8];
            float4 sr = make_float4(0.f, 0.f, 0.f, 0.f), srx = sr;
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                xh[l] = ok ? __ldg(reinterpret_cast<const float4*>(ly.xhat + o0 + l * 64)) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (relu) {
                    if (!(fmaf(xh[l].x, w4.x, b4.x) > 0.f)) x[l].x = 0.f;
                    if (!(fmaf(xh[l].y, w4.y, b4.y) > 0.f)) x[l].y = 0.f;
                    if (!(fmaf(xh[l].z, w4.z, b4.z) > 0.f)) x[l].z = 0.f;
                    if (!(fmaf(xh[l].w, w4.w, b4.w) > 0.f)) x[l].w = 0.f;
                }
                IINS_F4_OP(sr, sr, +, x[l]);
                srx.x = fmaf(x[l].x, xh[l].x, srx.x); srx.y = fmaf(x[l].y, xh[l].y, srx.y); srx.z = fmaf(x[l].z, xh[l].z, srx.z); srx.w = fmaf(x[l].w, xh[l].w, srx.w);
            }
            if (adain && ok) {
                *reinterpret_cast<float4*>(p.dadain + sb * p.adain_ld + ly.adain_off_b + 4 * cg) = sr;
                *reinterpret_cast<float4*>(p.dadain + sb * p.adain_ld + ly.adain_off_w + 4 * cg) = srx;
            }
            const float4 sc = make_float4(r4.x * w4.x, r4.y * w4.y, r4.z * w4.z, r4.w * w4.w);
            const float4 m = make_float4(sr.x * 0.125f, sr.y * 0.125f, sr.z * 0.125f, sr.w * 0.125f);
            const float4 mx = make_float4(srx.x * 0.125f, srx.y * 0.125f, srx.z * 0.125f, srx.w * 0.125f);
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                x[l].x = sc.x * (x[l].x - m.x - xh[l].x * mx.x); x[l].y = sc.y * (x[l].y - m.y - xh[l].y * mx.y);
                x[l].z = sc.z * (x[l].z - m.z - xh[l].z * mx.z); x[l].w = sc.w * (x[l].w - m.w - xh[l].w * mx.w);
                if (ok) *reinterpret_cast<float4*>(ly.dz + o0 + l * 64) = x[l];
            }
        };
        uint32_t j = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int gb = tile * TR_SAMPLES + s;
            const bool ok = gb < p.B;
            const long sb = ok ? gb : 0;
            const long o0 = sb * 512 + 4 * cg;
            float4 x[8];
            // ---- stage 0: norm backward of the LAST convolution's norm (second conv of the last block: no ReLU)
#pragma unroll
            for (int l = 0; l < 8; ++l) x[l] = ok ? __ldg(reinterpret_cast<const float4*>(p.dh + o0 + l * 64)) : make_float4(0.f, 0.f, 0.f, 0.f);
            norm_backward(p.layer[p.nconv - 1], false, x, o0, sb, ok, ADAIN);
            write_operand_columns<PIECES, false>(sA, s, cg, x);
            umma::fence_async_smem();
            umma::mbar_arrive(umma::smem_u32(&a_ready));
            for (int c = p.nconv - 1; c >= 0; --c, ++j) {
                // ---- the 12 main k-steps have read the tile: clear its six middle rows (positions 1..6) for the two edge terms
                umma::mbar_wait(umma::smem_u32(&main_done), j & 1);
                {
                    unsigned char* base = sA + (cg >> 1) * TR_LBO + (s * 10) * 16 + (cg & 1) * 8;
#pragma unroll
                    for (int l = 1; l <= 6; ++l)
#pragma unroll
                        for (int pc = 0; pc < PIECES; ++pc) *reinterpret_cast<uint2*>(base + pc * TR_APIECE + (l + 1) * 16) = make_uint2(0u, 0u);
                }
                umma::fence_async_smem();
                umma::mbar_arrive(umma::smem_u32(&z_ready));
                umma::mbar_wait(umma::smem_u32(&acc_full), j & 1);
                umma::tc_fence_after();
                stage_accumulator<PIECES>(tmem, warp, lane, stg);
                umma::tc_fence_before();
                epi_bar();
                load_columns(stg, s, cg, x);
                epi_bar();
                const bool first = (c & 1) == 0;                            // first convolution of its block: + the skip gradient
                if (first && ok) {   // plain loads: dh_scratch is written by this very thread two stages earlier
                    const float* skip = (c == p.nconv - 2 ? p.dh : p.dh_scratch) + o0;
#pragma unroll
                    for (int l = 0; l < 8; ++l) { const float4 h4 = *reinterpret_cast<const float4*>(skip + l * 64); IINS_F4_OP(x[l], x[l], +, h4); }
                }
                if (c == 0) {
                    if (p.dx != nullptr && ok) {
#pragma unroll
                        for (int l = 0; l < 8; ++l) *reinterpret_cast<float4*>(p.dx + o0 + l * 64) = x[l];
                    }
                    if (p.pre.xhat != nullptr) norm_backward(p.pre, p.pre_relu != 0, x, o0, sb, ok, false);
                    break;
                }
                if (first && ok) {                                          // gradient w.r.t. the previous block's output: the next skip term
#pragma unroll
                    for (int l = 0; l < 8; ++l) *reinterpret_cast<float4*>(p.dh_scratch + o0 + l * 64) = x[l];
                }
                norm_backward(p.layer[c - 1], ((c - 1) & 1) == 0, x, o0, sb, ok, ADAIN);
                write_operand_columns<PIECES, false>(sA, s, cg, x);
                umma::fence_async_smem();
                umma::mbar_arrive(umma::smem_u32(&a_ready));
            }
            ++j;                                                            // the `break` at c == 0 skipped the loop increment
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 8) umma::tmem_dealloc(tmem, TCOLS);
}

// cuTensorMapEncodeTiled is a DRIVER API entry point: it is resolved through the runtime (cudaGetDriverEntryPoint) so that the
// library carries no link-time dependency on libcuda.so.1 (it must load on a GPU-less build / CI host as well).
typedef CUresult (*IinsEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
IinsEncodeTiledFn encode_tiled_fn() {
    static IinsEncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<IinsEncodeTiledFn>(ptr);
        else (void)cudaGetLastError();
    }
    return fn;
}

bool make_weight_map(CUtensorMap* map, const void* base, int pieces, int nconv) {
    IinsEncodeTiledFn encode = encode_tiled_fn();
    if (encode == nullptr) return false;                      // -> the plain bulk-copy (cp.async.bulk) variant of the kernels
    // the packed weights as a 2-D array of 512-byte rows: one slice = SLICE / 512 consecutive rows (box = the whole slice)
    const unsigned slice = (unsigned)pieces * 64 * 16 * 2;
    const cuuint64_t dims[2] = {256, (cuuint64_t)(slice / 512) * TR_KS * nconv};
    const cuuint64_t strides[1] = {512};
    const cuuint32_t box[2] = {256, slice / 512};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int PIECES, bool ADAIN, bool TMAP>
void launch_variant(cudaStream_t st, const IinsTrunkFwdParams& p, const CUtensorMap& map, int grid) {
    constexpr int smem = (int)TrunkSmem<PIECES>::TOTAL;
    static bool attr = false;
    auto iins_trunk_fwd_kernel_ = iins_trunk_fwd_kernel<PIECES, ADAIN, TMAP>;
    if (!attr) { cudaFuncSetAttribute(iins_trunk_fwd_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
    IINS_LAUNCH(iins_trunk_fwd_kernel_, grid, 320, smem, st, p, map);
}

template <int PIECES, bool ADAIN, bool TMAP>
void launch_bwd_variant(cudaStream_t st, const IinsTrunkBwdParams& p, const CUtensorMap& map, int grid) {
    constexpr int smem = (int)TrunkSmem<PIECES>::TOTAL;
    static bool attr = false;
    auto iins_trunk_bwd_kernel_ = iins_trunk_bwd_kernel<PIECES, ADAIN, TMAP>;
    if (!attr) { cudaFuncSetAttribute(iins_trunk_bwd_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
    IINS_LAUNCH(iins_trunk_bwd_kernel_, grid, 320, smem, st, p, map);
}

}  // namespace

bool iins_trunk_backward_launch(cudaStream_t st, const IinsTrunkBwdParams& p) {
    if (p.B < 1 || p.nconv < 2 || p.nconv > IINS_TRUNK_MAX_CONVS || (p.nconv & 1) || (p.pieces != 1 && p.pieces != 3)) return false;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (!al16(p.dh) || !al16(p.dx) || !al16(p.dh_scratch) || !al16(p.adain) || !al16(p.dadain) || (reinterpret_cast<uintptr_t>(p.wpack) & 127) != 0) return false;
    if (p.dh == nullptr || (p.dx == nullptr && p.pre.xhat == nullptr) || (p.nconv > 2 && p.dh_scratch == nullptr)) return false;
    if (p.adain != nullptr && ((p.adain_ld & 3) || p.dadain == nullptr)) return false;
    for (int c = 0; c < p.nconv; ++c) {
        const IinsTrunkBwdLayer& l = p.layer[c];
        if (!l.xhat || !l.rstd || !l.dz || !al16(l.xhat) || !al16(l.rstd) || !al16(l.dz)) return false;
        if (p.adain != nullptr && ((l.adain_off_b & 3) || (l.adain_off_w & 3))) return false;
    }
    if (p.pre.xhat != nullptr && (!p.pre.rstd || !p.pre.dz || !al16(p.pre.xhat) || !al16(p.pre.rstd) || !al16(p.pre.dz) || p.adain != nullptr)) return false;
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    bool tmap = p.use_tmap != 0 && make_weight_map(&map, p.wpack, p.pieces, p.nconv);
    const int ntiles = (p.B + TR_SAMPLES - 1) / TR_SAMPLES;
    const int grid = ntiles < 2 * 148 ? ntiles : 2 * 148;
    IINS_SET_FLOPS(2.0 * (double)p.B * 8.0 * 64.0 * 192.0 * p.nconv); IINS_SET_SHAPE(p.B * 8, 64, 192 * p.nconv);
    // reads dh and every layer's x-hat, writes every layer's dz and the input gradient (+ the block-skip gradients once each way)
    IINS_SET_BYTES(2048.0 * p.B * (2.0 * p.nconv + 2.0 + p.nconv));
    const bool adain = p.adain != nullptr;
#define IINS_TRB(P_, A_, T_) if (p.pieces == P_ && adain == A_ && tmap == T_) { launch_bwd_variant<P_, A_, T_>(st, p, map, grid); return true; }
    IINS_TRB(3, false, true) IINS_TRB(3, true, true) IINS_TRB(1, false, true) IINS_TRB(1, true, true)
    IINS_TRB(3, false, false) IINS_TRB(3, true, false) IINS_TRB(1, false, false) IINS_TRB(1, true, false)
#undef IINS_TRB
    return false;
}

bool iins_trunk_forward_launch(cudaStream_t st, const IinsTrunkFwdParams& p) {
    if (p.B < 1 || p.nconv < 2 || p.nconv > IINS_TRUNK_MAX_CONVS || (p.nconv & 1) || (p.pieces != 1 && p.pieces != 3)) return false;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (!al16(p.x) || !al16(p.adain) || (reinterpret_cast<uintptr_t>(p.wpack) & 127) != 0) return false;
    if (p.adain != nullptr && (p.adain_ld & 3)) return false;
    for (int c = 0; c < p.nconv; ++c) {
        const IinsTrunkLayer& l = p.layer[c];
        if (!al16(l.y) || !al16(l.xhat) || !al16(l.rstd) || l.bias == nullptr) return false;
        if (p.adain != nullptr && ((l.adain_off_b & 3) || (l.adain_off_w & 3))) return false;
    }
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    bool tmap = p.use_tmap != 0 && make_weight_map(&map, p.wpack, p.pieces, p.nconv);
    const int ntiles = (p.B + TR_SAMPLES - 1) / TR_SAMPLES;
    const int grid = ntiles < 2 * 148 ? ntiles : 2 * 148;                  // persistent: 2 CTAs per SM
    IINS_SET_FLOPS(2.0 * (double)p.B * 8.0 * 64.0 * 192.0 * p.nconv); IINS_SET_SHAPE(p.B * 8, 64, 192 * p.nconv);
    // reads the input, writes y and x-hat of every layer, re-reads the block input for the skip
    IINS_SET_BYTES(2048.0 * p.B * (1.0 + 2.0 * p.nconv + 0.5 * p.nconv));
    const bool adain = p.adain != nullptr;
#define IINS_TRV(P_, A_, T_) if (p.pieces == P_ && adain == A_ && tmap == T_) { launch_variant<P_, A_, T_>(st, p, map, grid); return true; }
    IINS_TRV(3, false, true) IINS_TRV(3, true, true) IINS_TRV(1, false, true) IINS_TRV(1, true, true)
    IINS_TRV(3, false, false) IINS_TRV(3, true, false) IINS_TRV(1, false, false) IINS_TRV(1, true, false)
#undef IINS_TRV
    return false;
}
