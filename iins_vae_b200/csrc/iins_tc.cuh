// Tensor-core (tcgen05 / TMEM) implicit-GEMM kernels for sm_100a.
//
//   iins_tc_nt_kernel : forward + data-gradient GEMMs.  A tile (128 rows x 32 k) is gathered from the
//       channels-last activations (im2col on the fly: reflection / zero padding, stride, nearest x2
//       upsampling, or the transposed tap map for dgrad), split into bf16 pieces and written to shared
//       memory in the UMMA K-major no-swizzle layout; the weight tile arrives pre-split through a TMA bulk
//       copy (cp.async.bulk + mbarrier); one thread issues tcgen05.mma (M=128, N=16..64, K=16), the fp32
//       accumulator lives in TMEM, the epilogue (bias, InstanceNorm / AdaIN / LayerNorm, activation,
//       residual) runs on the TMEM -> SMEM staged tile.
//   iins_tc_tn_kernel : weight gradient  dW^T[k][n] = sum_rows A[row][k] * dz[row][n]  with both operands
//       MN-major (the reduction runs over rows), accumulated in TMEM over the CTA's row range and flushed
//       with atomics.
//
// Precision: fp32 parity needs more than TF32/BF16 single-pass products, so each fp32 operand is split into
// three bf16 pieces (8+8+8 mantissa bits) and the six significant piece products are accumulated in fp32
// ("bf16x3": error ~2^-23, same class as an fp32 FMA chain).  pieces == 1 is the plain bf16 mode.
#pragma once
#include "iins_gemm.cuh"
#ifndef IINS_CPUSIM
#include "iins_umma.cuh"

// ----------------------------------------------------------------------------------------- 8-wide gathers
// generic (scalar) gathers for layouts / channel counts the 16-byte fast paths do not cover; kept out of line
// and rolled so the kernels stay small (instruction-cache footprint)
__device__ __noinline__ void iins_gather8_fwd_generic(const IinsGeom& g, const float* __restrict__ x, int K, int b, int l, int k0, float* v) {
    int t = k0 / g.Cin, c = k0 - t * g.Cin;
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        v[i] = (k0 + i < K) ? iins_a_fwd(g, x, b, l, t, c) : 0.f;
        if (++c == g.Cin) { c = 0; ++t; }
    }
}
__device__ __noinline__ void iins_dz8_generic(const IinsGeom& g, const IinsDz& d, int b, int l, int n0, float* v, bool accumulate) {
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        float u = (n0 + i < g.Cout) ? iins_dz_at(g, d, b, l, n0 + i) : 0.f;
        v[i] = accumulate ? v[i] + u : u;
    }
}
__device__ __noinline__ void iins_gather8_dgrad_generic(const IinsGeom& g, const IinsDz& d, int K, int b, int pos, int k0, float* v) {
    int t = k0 / g.Cout, c = k0 - t * g.Cout;
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        v[i] = (k0 + i < K) ? iins_a_dgrad(g, d, b, pos, t, c) : 0.f;
        if (++c == g.Cout) { c = 0; ++t; }
    }
}

// forward A operand: 8 consecutive k of row (b,l), k0 % 8 == 0
__device__ __forceinline__ void iins_gather8_fwd(const IinsGeom& g, const float* __restrict__ x, int K, int b, int l, int k0, float* v) {
    if (k0 >= K) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        return;
    }
    if ((g.Cin & 7) == 0 && g.in_layout == IINS_NLC) {
        int t = k0 / g.Cin, c0 = k0 - t * g.Cin;
        int pos = iins_src_pos(g, l, t);
        if (pos < 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
            return;
        }
        const float4* src = reinterpret_cast<const float4*>(x + ((long)b * g.Lin + pos) * g.Cin + c0);
        float4 a = __ldg(src), c = __ldg(src + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
        return;
    }
    iins_gather8_fwd_generic(g, x, K, b, l, k0, v);
}

// 8 consecutive output channels of dz at (b,l): n0 % 8 == 0
__device__ __forceinline__ void iins_dz8(const IinsGeom& g, const IinsDz& d, int b, int l, int n0, float* v, bool accumulate) {
    if ((g.Cout & 7) == 0 && g.out_layout == IINS_NLC) {
        long idx = ((long)b * g.Lout + l) * g.Cout + n0;
        const float4* src = reinterpret_cast<const float4*>(d.dy_bcast ? d.dy + (long)b * g.Cout + n0 : d.dy + idx);
        float4 a = __ldg(src), c = __ldg(src + 1);
        float u[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
        if (d.y != nullptr && d.act != IINS_ACT_NONE) {
            const float4* ys = reinterpret_cast<const float4*>(d.y + idx);
            float4 ya = __ldg(ys), yc = __ldg(ys + 1);
            float yy[8] = {ya.x, ya.y, ya.z, ya.w, yc.x, yc.y, yc.z, yc.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] *= iins_dact_from_y(yy[i], d.act, d.slope);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = accumulate ? v[i] + u[i] * d.dy_scale : u[i] * d.dy_scale;
        return;
    }
    iins_dz8_generic(g, d, b, l, n0, v, accumulate);
}

// dgrad A operand: 8 consecutive k = (t, co0..co0+7) of input row (b,pos)
__device__ __forceinline__ void iins_gather8_dgrad(const IinsGeom& g, const IinsDz& d, int K, int b, int pos, int k0, float* v) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    if (k0 >= K) return;
    if ((g.Cout & 7) == 0 && g.out_layout == IINS_NLC) {
        int t = k0 / g.Cout, c0 = k0 - t * g.Cout;
        int q[3], nq;
        if (g.mode == IINS_PAD_REFLECT) {
            q[0] = pos + g.pad; nq = 1;
            if (pos >= 1 && pos <= g.pad) q[nq++] = g.pad - pos;
            if (pos <= g.Lin - 2 && pos >= g.Lin - 1 - g.pad) q[nq++] = g.pad + 2 * (g.Lin - 1) - pos;
        } else if (g.mode == IINS_PAD_UP2) {
            q[0] = 2 * pos + g.pad; q[1] = 2 * pos + 1 + g.pad; nq = 2;
        } else {
            q[0] = pos + g.pad; nq = 1;
        }
#pragma unroll 1
        for (int j = 0; j < nq; ++j) {
            int r = q[j] - t;
            if (r < 0) continue;
            int l = r / g.stride;
            if (l * g.stride != r || l >= g.Lout) continue;
            iins_dz8(g, d, b, l, c0, v, true);
        }
        return;
    }
    iins_gather8_dgrad_generic(g, d, K, b, pos, k0, v);
}

// split 8 floats into bf16 pieces and store one 16-byte vector per piece
__device__ __forceinline__ void iins_store8_split(const float* v, unsigned char* base, uint32_t piece_stride, int pieces) {
    uint32_t w[3][4];
    if (pieces == 3) {
#pragma unroll
        for (int q = 0; q < 4; ++q) umma::split3_pair(v[2 * q], v[2 * q + 1], w[0][q], w[1][q], w[2][q]);
        *reinterpret_cast<uint4*>(base) = make_uint4(w[0][0], w[0][1], w[0][2], w[0][3]);
        *reinterpret_cast<uint4*>(base + piece_stride) = make_uint4(w[1][0], w[1][1], w[1][2], w[1][3]);
        *reinterpret_cast<uint4*>(base + 2 * piece_stride) = make_uint4(w[2][0], w[2][1], w[2][2], w[2][3]);
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) w[0][q] = umma::cvt_bf16x2(v[2 * q], v[2 * q + 1]);
        *reinterpret_cast<uint4*>(base) = make_uint4(w[0][0], w[0][1], w[0][2], w[0][3]);
    }
}

// ------------------------------------------------------------------------------------ weight packing
// Bw[n][k] (fwd: n = co, k = t*Cin+ci; dgrad: n = ci, k = t*Cout+co) -> bf16 pieces in the UMMA K-major
// tile layout  [n block][k block of 32][piece][chunk of 8 k][NT rows][8]  so that one (n block, k block)
// is ONE contiguous TMA bulk copy.  Zero fill outside (N, K).
struct IinsPackParams {
    IinsGeom g;
    int kind;            // 0 fwd, 1 dgrad
    const float* w;
    uint16_t* out;
    int N, K, NT, nkb, nblk;
};

__global__ void __launch_bounds__(256) iins_pack_kernel(const IinsPackParams p) {
    // one thread per 16-byte destination chunk (n block, k block, chunk, row)
    const long total = (long)p.nblk * p.nkb * 4 * p.NT;
    const int Cdim = p.kind == 0 ? p.g.Cin : p.g.Cout;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        int nn = (int)(e % p.NT);
        long r = e / p.NT;
        int chunk = (int)(r & 3);
        r >>= 2;
        int kb = (int)(r % p.nkb), nb = (int)(r / p.nkb);
        int n = nb * p.NT + nn;
        int k0 = kb * 32 + chunk * 8;
        float v[8];
        int t = k0 / Cdim, c = k0 - t * Cdim;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float x = 0.f;
            if (n < p.N && k0 + i < p.K) x = __ldg(p.w + (p.kind == 0 ? iins_w_index(p.g, n, c, t) : iins_w_index(p.g, c, n, t)));
            v[i] = x;
            if (++c == Cdim) { c = 0; ++t; }
        }
        const uint32_t piece_stride = 4u * p.NT * 16u;
        unsigned char* base = reinterpret_cast<unsigned char*>(p.out) +
                              ((long)(nb * p.nkb + kb) * 3) * piece_stride + ((long)chunk * p.NT + nn) * 16;
        iins_store8_split(v, base, piece_stride, 3);
    }
}

// --------------------------------------------------------------------------------- forward / dgrad GEMM
struct IinsTCParams {
    IinsNTParams nt;
    const uint16_t* wpack;
    int pieces;          // 3 (fp32-grade) or 1 (bf16)
    int nkb;             // K blocks of 32
};

template <int NT>
__global__ void __launch_bounds__(256) iins_tc_nt_kernel(const IinsTCParams tp) {
    constexpr int BM = 128;
    constexpr uint32_t A_PIECE = 4 * BM * 16;            // 8192 B : [chunk][row][16 B]
    constexpr uint32_t B_PIECE = 4 * NT * 16;
    constexpr uint32_t STAGE = 3 * A_PIECE + 3 * B_PIECE;
    constexpr int LD = NT + 1;
    extern __shared__ __align__(1024) unsigned char dsm[];
    __shared__ __align__(8) unsigned long long mbar_mma[2];
    __shared__ __align__(8) unsigned long long mbar_b[2];
    __shared__ uint32_t tmem_slot;
    const IinsNTParams& p = tp.nt;
    const IinsGeom& g = p.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile_m = blockIdx.x * BM, n0 = blockIdx.y * NT;
    float* Cs = reinterpret_cast<float*>(dsm);                         // aliases the stages after the MMAs
    float* st_mean = reinterpret_cast<float*>(dsm + 2 * STAGE);
    float* st_rstd = st_mean + 1024;

    if (tid == 0) {
        umma::mbar_init(umma::smem_u32(&mbar_mma[0]), 1);
        umma::mbar_init(umma::smem_u32(&mbar_mma[1]), 1);
        umma::mbar_init(umma::smem_u32(&mbar_b[0]), 1);
        umma::mbar_init(umma::smem_u32(&mbar_b[1]), 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(umma::smem_u32(&tmem_slot), 64);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;

    const int a_row = tid & 127, a_half = tid >> 7;
    const int grow = tile_m + a_row;
    const bool a_ok = grow < p.M;
    const int a_b = a_ok ? grow / p.Lrow : 0;
    const int a_l = a_ok ? grow - a_b * p.Lrow : 0;
    const int nkb = tp.nkb;
    const uint32_t idesc = umma::make_idesc_bf16(BM, NT, 0, 0);
    const uint32_t b_bytes = (tp.pieces == 3 ? 3u : 1u) * B_PIECE;

    for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb & 1;
        unsigned char* sA = dsm + s * STAGE;
        unsigned char* sB = sA + 3 * A_PIECE;
        if (kb >= 2) umma::mbar_wait(umma::smem_u32(&mbar_mma[s]), (uint32_t)(((kb >> 1) - 1) & 1));
        if (tid == 0) {
            const unsigned char* src = reinterpret_cast<const unsigned char*>(tp.wpack) +
                                       ((long)blockIdx.y * nkb + kb) * (3 * B_PIECE);
            umma::mbar_arrive_expect_tx(umma::smem_u32(&mbar_b[s]), b_bytes);
            umma::tma_bulk_g2s(umma::smem_u32(sB), src, b_bytes, umma::smem_u32(&mbar_b[s]));
        }
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
            const int j = a_half * 2 + jj;
            float v[8];
            if (a_ok) {
                if (p.a_kind == 0) iins_gather8_fwd(g, p.x, p.K, a_b, a_l, kb * 32 + j * 8, v);
                else iins_gather8_dgrad(g, p.dz, p.K, a_b, a_l, kb * 32 + j * 8, v);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = 0.f;
            }
            iins_store8_split(v, sA + (j * BM + a_row) * 16, A_PIECE, tp.pieces);
        }
        umma::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            umma::mbar_wait(umma::smem_u32(&mbar_b[s]), (uint32_t)((kb >> 1) & 1));
            umma::tc_fence_after();
            const uint32_t a0 = umma::smem_u32(sA), b0 = umma::smem_u32(sB);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                for (int pa = 0; pa < tp.pieces; ++pa)
                    for (int pb = 0; pb + pa < tp.pieces; ++pb) {
                        uint64_t ad = umma::make_desc(a0 + pa * A_PIECE + ks * 2 * BM * 16, BM * 16, 128);
                        uint64_t bd = umma::make_desc(b0 + pb * B_PIECE + ks * 2 * NT * 16, NT * 16, 128);
                        umma::mma_bf16_ss(tmem, ad, bd, idesc, (kb | ks | pa | pb) ? 1u : 0u);
                    }
            }
            umma::commit(umma::smem_u32(&mbar_mma[s]));
        }
    }
    if (nkb >= 2) umma::mbar_wait(umma::smem_u32(&mbar_mma[(nkb - 2) & 1]), (uint32_t)(((nkb - 2) >> 1) & 1));
    umma::mbar_wait(umma::smem_u32(&mbar_mma[(nkb - 1) & 1]), (uint32_t)(((nkb - 1) >> 1) & 1));
    umma::tc_fence_after();

    // ---- TMEM -> SMEM (+ bias).  Warp w owns TMEM lanes 32*(w&3) .. +31; with NT >= 32 the two warps sharing
    // a lane quarter split the columns.
    {
        constexpr int COLS_PER_WARP = NT >= 32 ? NT / 2 : NT;
        const int q = warp & 3, hf = warp >> 2;
        if (NT >= 32 || hf == 0) {
            const int row = q * 32 + lane;
            const int cbeg = NT >= 32 ? hf * COLS_PER_WARP : 0;
#pragma unroll
            for (int c0 = 0; c0 < COLS_PER_WARP; c0 += 16) {
                float v[16];
                umma::tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(cbeg + c0), v);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    int n = n0 + cbeg + c0 + i;
                    float bv = (p.ep.bias != nullptr && n < p.N) ? __ldg(p.ep.bias + n) : 0.f;
                    Cs[row * LD + cbeg + c0 + i] = v[i] + bv;
                }
            }
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    iins_epilogue_tile<NT, LD>(p, Cs, st_mean, st_rstd, tile_m, n0);
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 64);
}

// ------------------------------------------------------------------------------------------- weight grad
struct IinsTCTNParams {
    IinsTNParams tn;
    int pieces;
    int K;               // ks * Cin
};

// grid = (row parts, ceil(K/128), ceil(Cout/NT)).  D^T[k][n] accumulated in TMEM (128 lanes = 128 k entries).
template <int NT>
__global__ void __launch_bounds__(256) iins_tc_tn_kernel(const IinsTCTNParams tp) {
    constexpr int BR = 32;                               // rows per stage (2 MMA k-steps of 16)
    constexpr uint32_t A_PIECE = 16 * BR * 16;           // [k group of 8][row][16 B] = 8192 B
    constexpr uint32_t B_PIECE = (NT / 8) * BR * 16;
    constexpr uint32_t STAGE = 3 * A_PIECE + 3 * B_PIECE;
    extern __shared__ __align__(1024) unsigned char dsm[];
    __shared__ __align__(8) unsigned long long mbar_mma[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float s_bias[64];
    const IinsTNParams& p = tp.tn;
    const IinsGeom& g = p.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ktile0 = blockIdx.y * 128, n0 = blockIdx.z * NT;
    const int r_begin = blockIdx.x * p.rows_per_part;
    int r_end = r_begin + p.rows_per_part;
    if (r_end > p.M) r_end = p.M;

    if (tid == 0) {
        umma::mbar_init(umma::smem_u32(&mbar_mma[0]), 1);
        umma::mbar_init(umma::smem_u32(&mbar_mma[1]), 1);
        umma::fence_mbar_init();
    }
    if (tid < 64) s_bias[tid] = 0.f;
    if (warp == 0) umma::tmem_alloc(umma::smem_u32(&tmem_slot), 64);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = umma::make_idesc_bf16(128, NT, 1, 1);
    const bool do_bias = p.db != nullptr && blockIdx.y == 0;
    float bsum[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) bsum[i] = 0.f;

    const int nit = (r_end - r_begin + BR - 1) / BR;
    for (int it = 0; it < nit; ++it) {
        const int s = it & 1;
        unsigned char* sA = dsm + s * STAGE;
        unsigned char* sB = sA + 3 * A_PIECE;
        if (it >= 2) umma::mbar_wait(umma::smem_u32(&mbar_mma[s]), (uint32_t)(((it >> 1) - 1) & 1));
        const int row = r_begin + it * BR + lane;
        const bool ok = row < r_end;
        const int b = ok ? row / g.Lout : 0;
        const int l = ok ? row - b * g.Lout : 0;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
            const int kg = warp + 8 * jj;                // k group (8 consecutive k) of this tile
            float v[8];
            if (ok) iins_gather8_fwd(g, p.x, tp.K, b, l, ktile0 + kg * 8, v);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = 0.f;
            }
            iins_store8_split(v, sA + (kg * BR + lane) * 16, A_PIECE, tp.pieces);
        }
        if (warp < NT / 8) {
            float v[8];
            if (ok) iins_dz8(g, p.dz, b, l, n0 + warp * 8, v, false);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = 0.f;
            }
            if (do_bias) {
#pragma unroll
                for (int i = 0; i < 8; ++i) bsum[i] += v[i];
            }
            iins_store8_split(v, sB + (warp * BR + lane) * 16, B_PIECE, tp.pieces);
        }
        umma::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            umma::tc_fence_after();
            const uint32_t a0 = umma::smem_u32(sA), b0 = umma::smem_u32(sB);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                for (int pa = 0; pa < tp.pieces; ++pa)
                    for (int pb = 0; pb + pa < tp.pieces; ++pb) {
                        // MN-major: LBO = distance between 8-row groups (128 B), SBO = distance between MN groups
                        uint64_t ad = umma::make_desc(a0 + pa * A_PIECE + ks * 256, 128, BR * 16);
                        uint64_t bd = umma::make_desc(b0 + pb * B_PIECE + ks * 256, 128, BR * 16);
                        umma::mma_bf16_ss(tmem, ad, bd, idesc, (it | ks | pa | pb) ? 1u : 0u);
                    }
            }
            umma::commit(umma::smem_u32(&mbar_mma[s]));
        }
    }
    if (nit >= 2) umma::mbar_wait(umma::smem_u32(&mbar_mma[(nit - 2) & 1]), (uint32_t)(((nit - 2) >> 1) & 1));
    if (nit >= 1) umma::mbar_wait(umma::smem_u32(&mbar_mma[(nit - 1) & 1]), (uint32_t)(((nit - 1) >> 1) & 1));
    umma::tc_fence_after();

    if (nit >= 1) {
        // TMEM lane = k entry of this tile, column = n.  Warp w owns lanes 32*(w&3)..; column halves as above.
        constexpr int COLS_PER_WARP = NT >= 32 ? NT / 2 : NT;
        const int q = warp & 3, hf = warp >> 2;
        if (NT >= 32 || hf == 0) {
            const int k = ktile0 + q * 32 + lane;
            const int cbeg = NT >= 32 ? hf * COLS_PER_WARP : 0;
            const bool kok = k < tp.K;
            const int t = kok ? k / g.Cin : 0, ci = kok ? k - t * g.Cin : 0;
#pragma unroll
            for (int c0 = 0; c0 < COLS_PER_WARP; c0 += 16) {
                float v[16];
                umma::tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(cbeg + c0), v);
                if (kok) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        int n = n0 + cbeg + c0 + i;
                        if (n < g.Cout) atomicAdd(p.dw + iins_w_index(g, n, ci, t), v[i]);
                    }
                }
            }
        }
        if (do_bias && warp < NT / 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) atomicAdd(&s_bias[warp * 8 + i], bsum[i]);
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (do_bias && tid < NT && n0 + tid < g.Cout && nit >= 1) atomicAdd(p.db + n0 + tid, s_bias[tid]);
    if (warp == 0) umma::tmem_dealloc(tmem, 64);
}

#endif  // !IINS_CPUSIM
