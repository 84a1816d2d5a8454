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
constexpr int TR_ROWS = TR_SAMPLES * 10;             // operand rows incl. the reflected edge rows
constexpr uint32_t TR_LBO = TR_ROWS * 16;            // bytes between 8-channel chunks
constexpr uint32_t TR_SBO = 160;                     // bytes between 8-row groups (= samples)
constexpr uint32_t TR_APIECE = 8 * TR_LBO;           // one piece of the operand: 20480 B
constexpr int TR_NSLOT = 8;                          // weight ring slots
constexpr int TR_KS = 12;                            // k-steps of 16 per convolution: 3 taps x 4

__device__ __forceinline__ void tma_tensor_2d(uint32_t dst_saddr, const CUtensorMap* map, int c0, int c1, uint32_t mbar_saddr) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst_saddr), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(mbar_saddr)
                 : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

template <int PIECES, bool ADAIN, bool TMAP>
__global__ void __launch_bounds__(320, 2) iins_trunk_fwd_kernel(const IinsTrunkFwdParams p, const __grid_constant__ CUtensorMap wmap) {
    constexpr uint32_t SLICE = PIECES * 64 * 16 * 2;             // one (tap, 16-channel) weight slice: [2 chunks][PIECES*64 rows][16 B]
    constexpr uint32_t A_BYTES = PIECES * TR_APIECE;
    constexpr int TCOLS = IinsTmemCols<64, PIECES>::value;
    extern __shared__ __align__(1024) unsigned char dsm[];
    __shared__ __align__(8) unsigned long long w_full[TR_NSLOT];
    __shared__ __align__(8) unsigned long long w_empty[TR_NSLOT];
    __shared__ __align__(8) unsigned long long acc_full;
    __shared__ __align__(8) unsigned long long a_ready;
    __shared__ uint32_t tmem_slot;
    unsigned char* sA = dsm;
    unsigned char* sW = dsm + A_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = (p.B + TR_SAMPLES - 1) / TR_SAMPLES;

    iins_pdl_launch_dependents();
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
        // ------------------------------------------------------------------ epilogue warps: thread = one GEMM row, 32 channels
        const int q = warp & 3, hf = warp >> 2;
        const int row = q * 32 + lane;
        const int s = row >> 3, l = row & 7;
        const int cbeg = hf * 32;
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
        unsigned char* a_row = sA + (s * 10 + l + 1) * 16;                  // this row inside a chunk (piece 0)
        // the reflected copies of ReflectionPad1d(1): position 1 also fills row 0, position 6 also fills row 9
        const int extra = l == 1 ? -2 * 16 : (l == 6 ? 2 * 16 : 0);
        auto write_operand = [&](const float* v, int c0) {                  // 16 channels starting at cbeg + c0
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                unsigned char* dst = a_row + ((cbeg + c0) / 8 + h) * TR_LBO;
                iins_store8_split(v + 8 * h, dst, TR_APIECE, PIECES);
                if (extra != 0) iins_store8_split(v + 8 * h, dst + extra, TR_APIECE, PIECES);
            }
        };
        uint32_t j = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int gb = tile * TR_SAMPLES + s;
            const bool row_ok = gb < p.B;
            const long orow = ((long)(row_ok ? gb : 0) * 8 + l) * 64 + cbeg;
            // ---- stage 0: the trunk input (fp32, channels-last) -> bf16 pieces in A
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 16) {
                float v[16];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 t = row_ok ? __ldg(reinterpret_cast<const float4*>(p.x + orow + c0) + k) : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
                }
                write_operand(v, c0);
            }
            umma::fence_async_smem();
            umma::mbar_arrive(umma::smem_u32(&a_ready));
            for (int c = 0; c < p.nconv; ++c, ++j) {
                const IinsTrunkLayer& ly = p.layer[c];
                const bool second = (c & 1) != 0;                           // second convolution of a block: no ReLU, + skip
                const float* skip = !second ? nullptr : (c == 1 ? p.x : p.layer[c - 2].y);
                umma::mbar_wait(umma::smem_u32(&acc_full), j & 1);
                umma::tc_fence_after();
#pragma unroll 1
                for (int c0 = 0; c0 < 32; c0 += 16) {
                    const int gn = cbeg + c0;
                    float4 a4[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        // plain (coherent) load: for c >= 3 this thread wrote these very values earlier in the kernel
                        a4[k] = (second && row_ok) ? *(reinterpret_cast<const float4*>(skip + orow + c0) + k) : make_float4(0.f, 0.f, 0.f, 0.f);
                    float v[16];
                    iins_tmem_chunk16<64, PIECES>(tl, gn, v);
#pragma unroll
                    for (int k = 0; k < 16; ++k) v[k] += __ldg(ly.bias + gn + k);
                    // InstanceNorm statistics over the 8 positions of the sample = the 8 lanes of this lane group;
                    // biased variance, eps inside the square root (models.py:152, 1072)
                    float4* xh_dst = row_ok ? reinterpret_cast<float4*>(ly.xhat + orow + c0) : nullptr;
                    float4* rs_dst = (row_ok && l == 0) ? reinterpret_cast<float4*>(ly.rstd + (long)gb * 64 + gn) : nullptr;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float rs[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float mean = iins_lanes_sum<8>(v[4 * k + e]) * 0.125f;
                            const float d = v[4 * k + e] - mean;
                            const float vpe = fmaf(iins_lanes_sum<8>(d * d), 0.125f, IINS_EPS);
                            float r = rsqrtf(vpe);
                            r = r * fmaf(-0.5f * vpe, r * r, 1.5f);          // one Newton step: full fp32 accuracy
                            rs[e] = r;
                            v[4 * k + e] = d * r;
                        }
                        if (rs_dst != nullptr) rs_dst[k] = make_float4(rs[0], rs[1], rs[2], rs[3]);
                        if (xh_dst != nullptr) xh_dst[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                    }
                    if (ADAIN) {
                        const float* ab = p.adain + (long)(row_ok ? gb : 0) * p.adain_ld + gn;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float4 w4 = __ldg(reinterpret_cast<const float4*>(ab + ly.adain_off_w) + k);
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(ab + ly.adain_off_b) + k);
                            v[4 * k] = fmaf(v[4 * k], w4.x, b4.x); v[4 * k + 1] = fmaf(v[4 * k + 1], w4.y, b4.y);
                            v[4 * k + 2] = fmaf(v[4 * k + 2], w4.z, b4.z); v[4 * k + 3] = fmaf(v[4 * k + 3], w4.w, b4.w);
                        }
                    }
                    if (!second) {
#pragma unroll
                        for (int k = 0; k < 16; ++k) v[k] = fmaxf(v[k], 0.f);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) { v[4 * k] += a4[k].x; v[4 * k + 1] += a4[k].y; v[4 * k + 2] += a4[k].z; v[4 * k + 3] += a4[k].w; }
                    }
                    if (row_ok) {
                        float4* dst = reinterpret_cast<float4*>(ly.y + orow + c0);
#pragma unroll
                        for (int k = 0; k < 4; ++k) dst[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                    }
                    if (c + 1 < p.nconv) write_operand(v, c0);              // the next convolution reads it in place
                }
                if (c + 1 < p.nconv) {
                    umma::tc_fence_before();                               // TMEM reads done before the next MMAs overwrite it
                    umma::fence_async_smem();
                    umma::mbar_arrive(umma::smem_u32(&a_ready));
                }
            }
            umma::tc_fence_before();
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
    constexpr uint32_t SLICE = PIECES * 64 * 16 * 2;
    constexpr uint32_t A_BYTES = PIECES * TR_APIECE;
    constexpr int TCOLS = IinsTmemCols<64, PIECES>::value;
    constexpr int NSL = 20;                                        // slices per convolution through the ring
    extern __shared__ __align__(1024) unsigned char dsm[];
    __shared__ __align__(8) unsigned long long w_full[TR_NSLOT];
    __shared__ __align__(8) unsigned long long w_empty[TR_NSLOT];
    __shared__ __align__(8) unsigned long long acc_full, main_done, a_ready, z_ready;
    __shared__ uint32_t tmem_slot;
    unsigned char* sA = dsm;
    unsigned char* sW = dsm + A_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = (p.B + TR_SAMPLES - 1) / TR_SAMPLES;

    iins_pdl_launch_dependents();
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
    // the zero rows 0 and 9 of every sample are written once (nothing overwrites them afterwards)
    for (int e = tid; e < (int)(A_BYTES / 16); e += 320) {
        const int r = (e % TR_ROWS) % 10;
        if (r == 0 || r == 9) *reinterpret_cast<uint4*>(sA + (size_t)e * 16) = make_uint4(0u, 0u, 0u, 0u);
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
        const int q = warp & 3, hf = warp >> 2;
        const int row = q * 32 + lane;
        const int s = row >> 3, l = row & 7;
        const int cbeg = hf * 32;
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
        unsigned char* a_row = sA + (s * 10 + l + 1) * 16;
        auto write_operand = [&](const float* v, int c0) {
#pragma unroll
            for (int h = 0; h < 2; ++h) iins_store8_split(v + 8 * h, a_row + ((cbeg + c0) / 8 + h) * TR_LBO, TR_APIECE, PIECES);
        };
        // InstanceNorm / AdaIN backward of layer `ly` applied to the 16 gradient values v (w.r.t. that layer's OUTPUT):
        //   dz = rstd * w * (g - mean_l(g) - xhat * mean_l(g * xhat)),  g = v masked by the layer's ReLU;  AdaIN: d bias = sum_l g,
        //   d weight = sum_l g * xhat.  v is replaced by dz; dz goes to HBM (the weight-gradient kernels read it).
        auto norm_backward16 = [&](const IinsTrunkBwdLayer& ly, bool relu, float* v, long orow, int gn, int c0, long sb, bool row_ok, bool lead) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 xq = row_ok ? __ldg(reinterpret_cast<const float4*>(ly.xhat + orow + c0) + k) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 r4 = __ldg(reinterpret_cast<const float4*>(ly.rstd + sb * 64 + gn) + k);
                float4 w4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ADAIN) {
                    w4 = __ldg(reinterpret_cast<const float4*>(p.adain + sb * p.adain_ld + ly.adain_off_w + gn) + k);
                    b4 = __ldg(reinterpret_cast<const float4*>(p.adain + sb * p.adain_ld + ly.adain_off_b + gn) + k);
                }
                const float xv[4] = {xq.x, xq.y, xq.z, xq.w}, rv[4] = {r4.x, r4.y, r4.z, r4.w};
                const float wv[4] = {w4.x, w4.y, w4.z, w4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
                float sr[4], srx[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float u = fmaf(xv[e], wv[e], bv[e]);
                    const float raw = (relu && !(u > 0.f)) ? 0.f : v[4 * k + e];
                    sr[e] = iins_lanes_sum<8>(raw);
                    srx[e] = iins_lanes_sum<8>(raw * xv[e]);
                    v[4 * k + e] = rv[e] * wv[e] * (raw - sr[e] * 0.125f - xv[e] * srx[e] * 0.125f);
                }
                if (row_ok) reinterpret_cast<float4*>(ly.dz + orow + c0)[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                if (ADAIN && lead) {
                    *reinterpret_cast<float4*>(p.dadain + sb * p.adain_ld + ly.adain_off_b + gn + 4 * k) = make_float4(sr[0], sr[1], sr[2], sr[3]);
                    *reinterpret_cast<float4*>(p.dadain + sb * p.adain_ld + ly.adain_off_w + gn + 4 * k) = make_float4(srx[0], srx[1], srx[2], srx[3]);
                }
            }
        };
        uint32_t j = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int gb = tile * TR_SAMPLES + s;
            const bool row_ok = gb < p.B;
            const long sb = row_ok ? gb : 0;
            const long orow = (sb * 8 + l) * 64 + cbeg;
            const bool lead = row_ok && l == 0;
            // ---- stage 0: norm backward of the LAST convolution's norm (second conv of the last block: no ReLU)
#pragma unroll 1
            for (int c0 = 0; c0 < 32; c0 += 16) {
                float v[16];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 t = row_ok ? __ldg(reinterpret_cast<const float4*>(p.dh + orow + c0) + k) : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
                }
                norm_backward16(p.layer[p.nconv - 1], false, v, orow, cbeg + c0, c0, sb, row_ok, lead);
                write_operand(v, c0);
            }
            umma::fence_async_smem();
            umma::mbar_arrive(umma::smem_u32(&a_ready));
            for (int c = p.nconv - 1; c >= 0; --c, ++j) {
                // ---- the 12 main k-steps have read the tile: clear its six middle rows for the two edge terms
                umma::mbar_wait(umma::smem_u32(&main_done), j & 1);
                if (l >= 1 && l <= 6) {
#pragma unroll
                    for (int h = 0; h < 4; ++h)
#pragma unroll
                        for (int pc = 0; pc < PIECES; ++pc)
                            *reinterpret_cast<uint4*>(a_row + (cbeg / 8 + h) * TR_LBO + pc * TR_APIECE) = make_uint4(0u, 0u, 0u, 0u);
                }
                umma::fence_async_smem();
                umma::mbar_arrive(umma::smem_u32(&z_ready));
                umma::mbar_wait(umma::smem_u32(&acc_full), j & 1);
                umma::tc_fence_after();
                const bool first = (c & 1) == 0;                            // first convolution of its block: + the skip gradient
                const float* skip = !first ? nullptr : (c == p.nconv - 2 ? p.dh : p.dh_scratch);
#pragma unroll 1
                for (int c0 = 0; c0 < 32; c0 += 16) {
                    const int gn = cbeg + c0;
                    float4 a4[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)   // plain loads: dh_scratch is written by this very thread two stages earlier
                        a4[k] = (first && row_ok) ? *(reinterpret_cast<const float4*>(skip + orow + c0) + k) : make_float4(0.f, 0.f, 0.f, 0.f);
                    float v[16];
                    iins_tmem_chunk16<64, PIECES>(tl, gn, v);
#pragma unroll
                    for (int k = 0; k < 4; ++k) { v[4 * k] += a4[k].x; v[4 * k + 1] += a4[k].y; v[4 * k + 2] += a4[k].z; v[4 * k + 3] += a4[k].w; }
                    if (c == 0) {
                        if (p.dx != nullptr && row_ok) {
                            float4* dst = reinterpret_cast<float4*>(p.dx + orow + c0);
#pragma unroll
                            for (int k = 0; k < 4; ++k) dst[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                        }
                        if (p.pre.xhat != nullptr) norm_backward16(p.pre, p.pre_relu != 0, v, orow, gn, c0, sb, row_ok, false);
                        continue;
                    }
                    if (first && row_ok) {                                  // gradient w.r.t. the previous block's output: the next skip term
                        float4* dst = reinterpret_cast<float4*>(p.dh_scratch + orow + c0);
#pragma unroll
                        for (int k = 0; k < 4; ++k) dst[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                    }
                    norm_backward16(p.layer[c - 1], ((c - 1) & 1) == 0, v, orow, gn, c0, sb, row_ok, lead);
                    write_operand(v, c0);
                }
                if (c > 0) {
                    umma::tc_fence_before();
                    umma::fence_async_smem();
                    umma::mbar_arrive(umma::smem_u32(&a_ready));
                }
            }
            umma::tc_fence_before();
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 8) umma::tmem_dealloc(tmem, TCOLS);
}

bool make_weight_map(CUtensorMap* map, const void* base, int pieces, int nconv) {
    // the packed weights as a 2-D array of 512-byte rows: one slice = SLICE / 512 consecutive rows (box = the whole slice)
    const unsigned slice = (unsigned)pieces * 64 * 16 * 2;
    const cuuint64_t dims[2] = {256, (cuuint64_t)(slice / 512) * TR_KS * nconv};
    const cuuint64_t strides[1] = {512};
    const cuuint32_t box[2] = {256, slice / 512};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int PIECES, bool ADAIN, bool TMAP>
void launch_variant(cudaStream_t st, const IinsTrunkFwdParams& p, const CUtensorMap& map, int grid) {
    constexpr int smem = PIECES * (int)TR_APIECE + TR_NSLOT * PIECES * 64 * 16 * 2;
    static bool attr = false;
    auto iins_trunk_fwd_kernel_ = iins_trunk_fwd_kernel<PIECES, ADAIN, TMAP>;
    if (!attr) { cudaFuncSetAttribute(iins_trunk_fwd_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
    IINS_LAUNCH(iins_trunk_fwd_kernel_, grid, 320, smem, st, p, map);
}

template <int PIECES, bool ADAIN, bool TMAP>
void launch_bwd_variant(cudaStream_t st, const IinsTrunkBwdParams& p, const CUtensorMap& map, int grid) {
    constexpr int smem = PIECES * (int)TR_APIECE + TR_NSLOT * PIECES * 64 * 16 * 2;
    static bool attr = false;
    auto iins_trunk_bwd_kernel_ = iins_trunk_bwd_kernel<PIECES, ADAIN, TMAP>;
    if (!attr) { cudaFuncSetAttribute(iins_trunk_bwd_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
    IINS_LAUNCH(iins_trunk_bwd_kernel_, grid, 320, smem, st, p, map);
}

int trunk_use_tmap() {
    static int use_tmap = -1;
    if (use_tmap < 0) { const char* e = getenv("IINS_TRUNK_TMAP"); use_tmap = e ? atoi(e) : 1; }
    return use_tmap;
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
    bool tmap = trunk_use_tmap() != 0 && make_weight_map(&map, p.wpack, p.pieces, p.nconv);
    const int ntiles = (p.B + TR_SAMPLES - 1) / TR_SAMPLES;
    const int grid = ntiles < 2 * 148 ? ntiles : 2 * 148;
    IINS_SET_FLOPS(2.0 * (double)p.B * 8.0 * 64.0 * 192.0 * p.nconv); IINS_SET_SHAPE(p.B * 8, 64, 192 * p.nconv);
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
    bool tmap = trunk_use_tmap() != 0 && make_weight_map(&map, p.wpack, p.pieces, p.nconv);
    const int ntiles = (p.B + TR_SAMPLES - 1) / TR_SAMPLES;
    const int grid = ntiles < 2 * 148 ? ntiles : 2 * 148;                  // persistent: 2 CTAs per SM
    IINS_SET_FLOPS(2.0 * (double)p.B * 8.0 * 64.0 * 192.0 * p.nconv); IINS_SET_SHAPE(p.B * 8, 64, 192 * p.nconv);
    const bool adain = p.adain != nullptr;
#define IINS_TRV(P_, A_, T_) if (p.pieces == P_ && adain == A_ && tmap == T_) { launch_variant<P_, A_, T_>(st, p, map, grid); return true; }
    IINS_TRV(3, false, true) IINS_TRV(3, true, true) IINS_TRV(1, false, true) IINS_TRV(1, true, true)
    IINS_TRV(3, false, false) IINS_TRV(3, true, false) IINS_TRV(1, false, false) IINS_TRV(1, true, false)
#undef IINS_TRV
    return false;
}
