// Tensor-core (tcgen05 / TMEM) implicit-GEMM kernels for sm_100a.
//
//   iins_tc_nt_kernel : forward + data-gradient GEMMs.  A tile (128 rows x 32 k) is gathered from the
//       channels-last activations (im2col on the fly: reflection / zero padding, stride, nearest x2
//       upsampling, or the transposed tap map for dgrad), split into bf16 pieces and written to shared
//       memory in the UMMA K-major no-swizzle layout; the weight tile arrives pre-split through a TMA bulk
//       copy (cp.async.bulk + mbarrier); one elected lane of the MMA warp issues tcgen05.mma (M=128, N=16..64 x
//       stacked pieces, K=16), the fp32 accumulator lives in TMEM, the epilogue (bias, InstanceNorm / AdaIN /
//       LayerNorm, activation, residual, or the fused norm BACKWARD of a data gradient) runs in registers straight
//       out of TMEM (iins_tc_epilogue_regs); a SMEM-staged generic epilogue remains as the fallback variant.
//   iins_tc_tn_kernel : weight gradient  dW^T[k][n] = sum_rows A[row][k] * dz[row][n]  with both operands
//       MN-major (the reduction runs over rows), accumulated in TMEM over the CTA's row range and flushed
//       with atomics.
//
// Precision: fp32 parity needs more than TF32/BF16 single-pass products, so each fp32 operand is split into
// three bf16 pieces (8+8+8 mantissa bits) and eight of the nine piece products are accumulated in fp32
// ("bf16x3": operand error ~2^-31; the fp32 accumulation inside the tensor core remains measurably noisier than an
// FMA chain).  pieces == 1 is the plain bf16 mode.
//
// Latency notes (measured with in-kernel clock traces during bring-up and with ncu): the MMA
// issue path must be warp-uniform (descriptors in uniform registers; a divergent `if (tid == 0)` costs ~170 cycles per
// MMA in R2UR traffic), the raw operand data of the next TWO K blocks is prefetched into registers, the row / k index
// splits are shifts (every L and channel count on the path is a power of 2), and every (tile width, operand kind,
// epilogue kind, rows per sample) combination is its own template instance (instruction-fetch stalls otherwise).
#pragma once
#include "iins_gemm.cuh"
#ifndef IINS_CPUSIM
#include "iins_umma.cuh"

__device__ __forceinline__ void iins_zero8(float* v) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
}
__device__ __forceinline__ void iins_ld8(const float* __restrict__ src, bool ok, float* v) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
    if (ok) {
        a = __ldg(reinterpret_cast<const float4*>(src));
        c = __ldg(reinterpret_cast<const float4*>(src) + 1);
    }
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
}

// forward A operand, fast path: NLC input, k = t * Cin + c with Cin = 1 << cs (>= 8), or a Linear layer (one tap, any
// Cin % 8 == 0: the host passes cs = 31 so that t = 0 and c0 = k0); 8 consecutive k = one tap, 8 channels
__device__ __forceinline__ void iins_gather8_fwd_fast(const IinsGeom& g, const float* __restrict__ x, int K, int cs, int b, int l,
                                                      int k0, float* v) {
    const int t = k0 >> cs, c0 = k0 & (int)((1u << cs) - 1u);
    const int pos = iins_src_pos(g, l, t);
    const bool ok = k0 < K && pos >= 0;
    iins_ld8(x + ((long)b * g.Lin + (ok ? pos : 0)) * g.Cin + c0, ok, v);
}

// 8 consecutive output channels of dz at (b,l), fast path: NLC, Cout % 8 == 0
__device__ __forceinline__ void iins_dz8_fast(const IinsGeom& g, const IinsDz& d, int b, int l, int n0, bool ok, float* v) {
    const long idx = ((long)b * g.Lout + (ok ? l : 0)) * g.Cout + n0;
    iins_ld8(d.dy_bcast ? d.dy + (long)b * g.Cout + n0 : d.dy + idx, ok, v);
    if (d.y != nullptr && d.act != IINS_ACT_NONE) {
        float yy[8];
        iins_ld8(d.y + idx, ok, yy);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= iins_dact_from_y(yy[i], d.act, d.slope);
    }
    if (d.dy_scale != 1.f) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= d.dy_scale;
    }
}

// dgrad A operand, fast path: sums the (<= 3) output rows whose tap t reads input position pos
__device__ __forceinline__ void iins_gather8_dgrad_fast(const IinsGeom& g, const IinsDz& d, int K, int cs, int b, int pos, int k0,
                                                        float* v) {
    const int t = k0 >> cs, c0 = k0 & (int)((1u << cs) - 1u);
    int q0 = pos + g.pad, q1 = -1, q2 = -1;
    if (g.mode == IINS_PAD_REFLECT) {
        if (pos >= 1 && pos <= g.pad) q1 = g.pad - pos;
        if (pos <= g.Lin - 2 && pos >= g.Lin - 1 - g.pad) q2 = g.pad + 2 * (g.Lin - 1) - pos;
        if (q1 < 0) { q1 = q2; q2 = -1; }              // a position has both reflected images only when Lin <= 2 * pad + 1
    } else if (g.mode == IINS_PAD_UP2) {
        q0 = 2 * pos + g.pad;
        q1 = q0 + 1;
    }
    const int sh = g.stride - 1;                       // stride is 1 or 2
    auto cand = [&](int q, float* u) {
        const int r = q - t;
        const int l = r >> sh;
        const bool ok = k0 < K && q >= 0 && r >= 0 && (r & sh) == 0 && l < g.Lout;
        iins_dz8_fast(g, d, b, l, c0, ok, u);
    };
    cand(q0, v);
    if (g.mode != IINS_PAD_ZERO) {                     // warp-uniform
        float u[8];
        cand(q1, u);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += u[i];
        if (q2 >= 0) {                                 // per-lane (no warp collective here: the caller may be divergent)
            cand(q2, u);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += u[i];
        }
    }
}

// Data gradient of a k4 / stride-2 / zero-pad-1 convolution split by the PARITY of the input position (AKIND 2).
// Input position pos = 2 j + par receives only the taps t with t = pos + 1 (mod 2): par 0 -> {1, 3}, par 1 -> {0, 2}; the
// plain gather would build K = 4 Cout entries per row of which every other 8-wide chunk is structurally zero.  Rows of one
// parity class form a GEMM with K' = 2 Cout:  k = u * Cout + co,  t = 2 u + (1 - par),  output row l = j - u + par.
__device__ __forceinline__ void iins_gather8_dgrad_parity(const IinsGeom& g, const IinsDz& d, int K2, int cs, int b, int j, int par,
                                                          int k0, float* v) {
    const int u = k0 >> cs, c0 = k0 & (int)((1u << cs) - 1u);
    const int l = j - u + par;
    const bool ok = k0 < K2 && l >= 0 && l < g.Lout;
    iins_dz8_fast(g, d, b, l, c0, ok, v);
}

// ---- quad gathers for the forward / data-gradient kernel -----------------------------------------------------
// A thread fetches 4 consecutive k (one 16-byte load) of one row; the 8 lanes that share a row cover one whole
// 32-wide K block of it (128 contiguous bytes when the block lies in one tap), so a warp-level load touches 4 rows x
// 128 B instead of 32 rows x 16 B: ~5x fewer L1 wavefronts than the lane-per-row mapping (measured: the L1 data
// pipe was the busiest unit of these kernels).
__device__ __forceinline__ float4 iins_ld4(const float* __restrict__ src, bool ok) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) a = __ldg(reinterpret_cast<const float4*>(src));
    return a;
}
__device__ __forceinline__ float4 iins_gather4_fwd(const IinsGeom& g, const float* __restrict__ x, int K, int cs, int b, int l, int k0) {
    const int t = k0 >> cs, c0 = k0 & (int)((1u << cs) - 1u);
    const int pos = iins_src_pos(g, l, t);
    const bool ok = k0 < K && pos >= 0;
    return iins_ld4(x + ((long)b * g.Lin + (ok ? pos : 0)) * g.Cin + c0, ok);
}
__device__ __forceinline__ float4 iins_dz4(const IinsGeom& g, const IinsDz& d, int b, int l, int n0, bool ok) {
    const long idx = ((long)b * g.Lout + (ok ? l : 0)) * g.Cout + n0;
    float4 v = iins_ld4(d.dy_bcast ? d.dy + (long)b * g.Cout + n0 : d.dy + idx, ok);
    if (d.y != nullptr && d.act != IINS_ACT_NONE) {
        const float4 y = iins_ld4(d.y + idx, ok);
        v.x *= iins_dact_from_y(y.x, d.act, d.slope); v.y *= iins_dact_from_y(y.y, d.act, d.slope);
        v.z *= iins_dact_from_y(y.z, d.act, d.slope); v.w *= iins_dact_from_y(y.w, d.act, d.slope);
    }
    if (d.dy_scale != 1.f) { v.x *= d.dy_scale; v.y *= d.dy_scale; v.z *= d.dy_scale; v.w *= d.dy_scale; }
    return v;
}
__device__ __forceinline__ float4 iins_gather4_dgrad(const IinsGeom& g, const IinsDz& d, int K, int cs, int b, int pos, int k0) {
    const int t = k0 >> cs, c0 = k0 & (int)((1u << cs) - 1u);
    int q0 = pos + g.pad, q1 = -1, q2 = -1;
    if (g.mode == IINS_PAD_REFLECT) {
        if (pos >= 1 && pos <= g.pad) q1 = g.pad - pos;
        if (pos <= g.Lin - 2 && pos >= g.Lin - 1 - g.pad) q2 = g.pad + 2 * (g.Lin - 1) - pos;
        if (q1 < 0) { q1 = q2; q2 = -1; }              // a position has both reflected images only when Lin <= 2 * pad + 1
    } else if (g.mode == IINS_PAD_UP2) {
        q0 = 2 * pos + g.pad;
        q1 = q0 + 1;
    }
    const int sh = g.stride - 1;                       // stride is 1 or 2
    auto cand = [&](int q) {
        const int r = q - t;
        const int l = r >> sh;
        const bool ok = k0 < K && q >= 0 && r >= 0 && (r & sh) == 0 && l < g.Lout;
        return iins_dz4(g, d, b, l, c0, ok);
    };
    float4 v = cand(q0);
    if (g.mode != IINS_PAD_ZERO) {                     // warp-uniform
        const float4 u = cand(q1);
        v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
        if (q2 >= 0) {                                 // per-lane (no warp collective here: the caller may be divergent)
            const float4 w = cand(q2);
            v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
    }
    return v;
}
// split 4 floats into bf16 pieces and store one 8-byte half of a 16-byte K chunk per piece
__device__ __forceinline__ void iins_store4_split(const float4 v, unsigned char* base, uint32_t piece_stride, int pieces) {
    if (pieces == 3) {
        uint32_t a0, a1, a2, b0, b1, b2;
        umma::split3_pair(v.x, v.y, a0, a1, a2);
        umma::split3_pair(v.z, v.w, b0, b1, b2);
        *reinterpret_cast<uint2*>(base) = make_uint2(a0, b0);
        *reinterpret_cast<uint2*>(base + piece_stride) = make_uint2(a1, b1);
        *reinterpret_cast<uint2*>(base + 2 * piece_stride) = make_uint2(a2, b2);
    } else {
        *reinterpret_cast<uint2*>(base) = make_uint2(umma::cvt_bf16x2(v.x, v.y), umma::cvt_bf16x2(v.z, v.w));
    }
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
    int pieces;          // 3 or 1: the pieces are STACKED along the tile's row (n) dimension
};

static __global__ void __launch_bounds__(256) iins_pack_kernel(const IinsPackParams p) {
    iins_pdl_enter();
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
        // tile = [chunk (4)][piece][NT rows][16 B]: row index of the stacked operand = piece * NT + n
        const uint32_t piece_stride = (uint32_t)p.NT * 16u;
        const long tile_bytes = 4L * p.pieces * p.NT * 16;
        unsigned char* base = reinterpret_cast<unsigned char*>(p.out) + (long)(nb * p.nkb + kb) * tile_bytes +
                              ((long)chunk * p.pieces * p.NT + nn) * 16;
        iins_store8_split(v, base, piece_stride, p.pieces);
    }
}

// All layers of one module pass packed by ONE launch: a table of jobs in the kernel parameters.
#define IINS_PACK_MAX_JOBS 48
struct IinsPackJob {
    const float* w;
    uint16_t* out;
    int Cin, Cout, ks, kind, N, K, NT, nkb, nblk;
    long chunk_begin;            // prefix sum of 16-byte destination chunks
};
struct IinsPackAllParams {
    int njobs, pieces;
    long total;
    IinsPackJob jobs[IINS_PACK_MAX_JOBS];
};

static __global__ void __launch_bounds__(256) iins_pack_all_kernel(const IinsPackAllParams pp) {
    iins_pdl_enter();
    int j = 0;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < pp.total; e += (long)gridDim.x * blockDim.x) {
        while (j + 1 < pp.njobs && e >= pp.jobs[j + 1].chunk_begin) ++j;
        const IinsPackJob& q = pp.jobs[j];
        const long le = e - q.chunk_begin;
        const int Cdim = q.kind == 0 ? q.Cin : q.Cout;
        int nn = (int)(le % q.NT);
        long r = le / q.NT;
        int chunk = (int)(r & 3);
        r >>= 2;
        int kb = (int)(r % q.nkb), nb = (int)(r / q.nkb);
        int n = nb * q.NT + nn;
        int k0 = kb * 32 + chunk * 8;
        float v[8];
        int t = k0 / Cdim, c = k0 - t * Cdim;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float x = 0.f;
            if (n < q.N && k0 + i < q.K) {
                // kinds 2 / 3: parity-split data gradient (even / odd input positions): k = u * Cout + co, tap = 2 u + 1 - parity
                const int tt = q.kind >= 2 ? 2 * t + (3 - q.kind) : t;
                long wi = q.kind == 0 ? ((long)n * q.Cin + c) * q.ks + tt : ((long)c * q.Cin + n) * q.ks + tt;
                x = __ldg(q.w + wi);
            }
            v[i] = x;
            if (++c == Cdim) { c = 0; ++t; }
        }
        const uint32_t piece_stride = (uint32_t)q.NT * 16u;
        const long tile_bytes = 4L * pp.pieces * q.NT * 16;
        unsigned char* base = reinterpret_cast<unsigned char*>(q.out) + (long)(nb * q.nkb + kb) * tile_bytes +
                              ((long)chunk * pp.pieces * q.NT + nn) * 16;
        iins_store8_split(v, base, piece_stride, pp.pieces);
    }
}

// Issue the MMAs of one k-step (K = 16) for the bf16x3 scheme with the B pieces stacked along N:
//   D[:, 0:3N) += A0 * [B0|B1|B2],  D[:, 0:3N) += A1 * [B0|B1|B2],  D[:, 0:2N) += A2 * [B0|B1]
// so eight of the nine piece products take 3 instructions; the epilogue adds the three N-wide column blocks.
// adesc / bdesc address piece 0 of this k-step; a_piece16 = byte distance between A pieces >> 4.
template <int NT, int PIECES, int AMAJ, int BMAJ>
__device__ __forceinline__ void iins_issue_kstep(uint32_t tmem, uint64_t adesc, uint64_t bdesc, uint32_t a_piece16, bool leader,
                                                 uint32_t first_acc) {
    if (PIECES == 3) {
        if (leader) {
            // A1 and A2 also take the wider stacked operand: the extra products (A1*B2, A2*B1) cost two more N blocks of
            // tensor-pipe time (not the bottleneck) and push the split error from 2^-23 to 2^-31 (only A2*B2 is dropped)
            umma::mma_bf16_ss(tmem, adesc, bdesc, umma::make_idesc_bf16(128, 3 * NT, AMAJ, BMAJ), first_acc);
            umma::mma_bf16_ss(tmem, adesc + a_piece16, bdesc, umma::make_idesc_bf16(128, 3 * NT, AMAJ, BMAJ), 1u);
            umma::mma_bf16_ss(tmem, adesc + 2 * a_piece16, bdesc, umma::make_idesc_bf16(128, 2 * NT, AMAJ, BMAJ), 1u);
        }
    } else {
        if (leader) umma::mma_bf16_ss(tmem, adesc, bdesc, umma::make_idesc_bf16(128, NT, AMAJ, BMAJ), first_acc);
    }
}

// TMEM -> registers for 16 accumulator columns starting at c (summing the stacked piece blocks)
template <int NT, int PIECES>
__device__ __forceinline__ void iins_tmem_acc16(uint32_t taddr_lane, int c, float* v) {
    umma::tmem_ld16(taddr_lane + (uint32_t)c, v);
    if (PIECES == 3) {
        float u[16], w[16];
        umma::tmem_ld16(taddr_lane + (uint32_t)(c + NT), u);
        umma::tmem_ld16(taddr_lane + (uint32_t)(c + 2 * NT), w);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += u[i] + w[i];
    }
}

template <int NT, int PIECES>
struct IinsTmemCols {            // power of two >= 32 covering PIECES * NT accumulator columns
    static constexpr int need = PIECES * NT;
    static constexpr int value = need <= 32 ? 32 : (need <= 64 ? 64 : (need <= 128 ? 128 : 256));
};


// ------------------------------------------------------------------------- register-resident tile epilogue
// The accumulator row of TMEM lane r belongs to thread (r & 31) of warp (r >> 5): a thread owns one GEMM row =
// one (sample, position) and CW consecutive channels.  The L rows of a sample are L consecutive lanes (L <= 32,
// a power of two), so the InstanceNorm / AdaIN statistics are xor-shuffle trees over lane bits < log2(L) and the
// custom LayerNorm's per-sample statistics add one exchange between the two warps that share a lane quarter.
// Nothing is staged in shared memory; x-hat / y / rstd leave the registers as 16-byte stores (each thread writes
// whole 64-byte runs of its own row).  Processed in chunks of 16 columns to keep the register footprint small
// (the LayerNorm passes re-read TMEM, which is cheap).
// Preconditions (checked on the host, IinsTCParams::ep_regs): NLC output, N % NT == 0, 16-byte aligned y / xhat /
// add / rstd / adain, Lrow <= 32 when a norm is fused, N == NT for LayerNorm.
template <int NT, int PIECES>
__device__ __forceinline__ void iins_tmem_chunk16(uint32_t taddr_lane, int c, float* v) {
    uint32_t r0[16], r1[16], r2[16];
    umma::tmem_ld16_nowait(taddr_lane + (uint32_t)c, r0);
    if (PIECES == 3) {
        umma::tmem_ld16_nowait(taddr_lane + (uint32_t)(c + NT), r1);
        umma::tmem_ld16_nowait(taddr_lane + (uint32_t)(c + 2 * NT), r2);
    }
    umma::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        v[i] = __uint_as_float(r0[i]);
        if (PIECES == 3) v[i] += __uint_as_float(r1[i]) + __uint_as_float(r2[i]);
    }
}

// sum over aligned groups of LL lanes (LL a compile-time power of two <= 32): a fully unrolled xor-shuffle tree
template <int LL>
__device__ __forceinline__ float iins_lanes_sum(float v) {
#pragma unroll
    for (int o = 1; o < LL; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// activation with the kind hoisted out of the element loop (the compiler does not unswitch it)
template <int N>
__device__ __forceinline__ void iins_act_vec(float* v, int act, float slope) {
    if (act == IINS_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = fmaxf(v[i], 0.f);
    } else if (act == IINS_ACT_LRELU) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * slope;
    }
}

// Epilogue kinds: every (tile width, operand kind, epilogue) combination is its own lean kernel -- one generic kernel
// with run-time switches was ~19k SASS instructions and stalled on instruction fetch more than on anything else.
enum { IINS_EPI_PLAIN = 0,     // bias, ReLU / LeakyReLU, residual / accumulate operand
       IINS_EPI_IN = 1,        // + InstanceNorm or AdaIN over the LL rows of a sample
       IINS_EPI_LN = 2,        // + the reference's custom LayerNorm over a whole sample (LL rows x NT channels)
       IINS_EPI_SMEM = 3,      // anything else: SMEM-staged generic tile epilogue (iins_epilogue_tile)
       IINS_EPI_NBWD = 4 };    // data gradient + fused InstanceNorm / AdaIN backward of the producing layer (LL rows / sample)

template <int NT, int PIECES, int EPI, int LL>
__device__ __forceinline__ void iins_tc_epilogue_regs(const IinsNTParams& p, uint32_t tmem, int tile_m, int n0, int warp, int lane,
                                                      float* xch, int row_scale = 1, int row_off = 0) {
    constexpr int CW = NT >= 32 ? NT / 2 : NT;          // columns per thread
    const IinsEpilogue& ep = p.ep;
    const int q = warp & 3, hf = warp >> 2;
    const bool active = NT >= 32 || hf == 0;            // warp-uniform
    const int row = q * 32 + lane;
    const int gr = tile_m + row;
    const bool row_ok = gr < p.M;
    const int cbeg = NT >= 32 ? hf * CW : 0;
    constexpr int L = LL;                               // rows per sample (== p.Lrow when a norm is fused)
    const int b = gr >> p.lshift, l = gr & (p.Lrow - 1);
    const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
    // (row_scale, row_off) = (2, parity) for the parity-split data gradient: GEMM row (b, j) is input position 2 j + parity
    const long orow = ((long)gr * row_scale + row_off) * p.N + n0 + cbeg;

    float ln_mean = 0.f, ln_rs = 0.f;
    if (EPI == IINS_EPI_LN) {
        // per-sample mean and UNBIASED std over (C*L), eps added to std (models.py:976-981); two passes
        const float nel = (float)(L * NT);
        float part = 0.f;
        if (active) {
#pragma unroll
            for (int c0 = 0; c0 < CW; c0 += 16) {
                float v[16];
                iins_tmem_chunk16<NT, PIECES>(tl, cbeg + c0, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) part += v[i] + (ep.bias != nullptr ? __ldg(ep.bias + n0 + cbeg + c0 + i) : 0.f);
            }
        }
        part = iins_lanes_sum<LL>(part);
        if (NT >= 32) {
            xch[hf * 128 + row] = part;
            iins_epi_sync<true>();
            part += xch[(hf ^ 1) * 128 + row];
        }
        ln_mean = part / nel;
        float sq = 0.f;
        if (active) {
#pragma unroll
            for (int c0 = 0; c0 < CW; c0 += 16) {
                float v[16];
                iins_tmem_chunk16<NT, PIECES>(tl, cbeg + c0, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) { const float d = v[i] + (ep.bias != nullptr ? __ldg(ep.bias + n0 + cbeg + c0 + i) : 0.f) - ln_mean; sq = fmaf(d, d, sq); }
            }
        }
        sq = iins_lanes_sum<LL>(sq);
        if (NT >= 32) {
            xch[256 + hf * 128 + row] = sq;
            iins_epi_sync<true>();
            sq += xch[256 + (hf ^ 1) * 128 + row];
        }
        ln_rs = 1.0f / (sqrtf(sq / (nel - 1.f)) + IINS_EPS);
        if (active && row_ok && l == 0 && hf == 0 && ep.rstd != nullptr) ep.rstd[b] = ln_rs;
    }
    if (!active) return;

#pragma unroll 1
    for (int c0 = 0; c0 < CW; c0 += 16) {
        const int gn = n0 + cbeg + c0;
        // operands that do not depend on the accumulator first: their latency hides behind the TMEM loads
        float4 a4[4];
        if (ep.add != nullptr && row_ok) {
#pragma unroll
            for (int j = 0; j < 4; ++j) a4[j] = __ldg(reinterpret_cast<const float4*>(ep.add + orow + c0) + j);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) a4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float4 x4[EPI == IINS_EPI_NBWD ? 4 : 1];            // fused norm backward: saved x-hat of the producing layer
        if (EPI == IINS_EPI_NBWD) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                x4[j] = row_ok ? __ldg(reinterpret_cast<const float4*>(ep.nb_xhat + orow + c0) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float v[16];
        iins_tmem_chunk16<NT, PIECES>(tl, cbeg + c0, v);
        if (ep.bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += __ldg(ep.bias + gn + i);
        }
        // x-hat and rstd are stored as soon as they exist (short live ranges: the kernel runs 2 CTAs / SM at <= 96 registers)
        float4* xh_dst = ((EPI == IINS_EPI_IN || EPI == IINS_EPI_LN) && ep.xhat != nullptr && row_ok) ? reinterpret_cast<float4*>(ep.xhat + orow + c0) : nullptr;
        if (EPI == IINS_EPI_IN) {
            // per (sample, channel) statistics over the L rows; biased variance (models.py:152, 1072)
            constexpr float invL = 1.0f / (float)L;
            float4* rs_dst = (ep.rstd != nullptr && row_ok && l == 0) ? reinterpret_cast<float4*>(ep.rstd + (long)b * p.N + gn) : nullptr;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float rs[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float mean = iins_lanes_sum<LL>(v[4 * j + i]) * invL;
                    const float d = v[4 * j + i] - mean;
                    const float vpe = fmaf(iins_lanes_sum<LL>(d * d), invL, IINS_EPS);
                    float r = rsqrtf(vpe);
                    r = r * fmaf(-0.5f * vpe, r * r, 1.5f);      // one Newton step: full fp32 accuracy
                    rs[i] = r;
                    v[4 * j + i] = d * r;
                }
                if (rs_dst != nullptr) rs_dst[j] = make_float4(rs[0], rs[1], rs[2], rs[3]);
                if (xh_dst != nullptr) xh_dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            if (ep.norm == IINS_NORM_ADAIN) {
                const float* ab = ep.adain + (long)(row_ok ? b : 0) * ep.adain_ld + gn;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 w4 = __ldg(reinterpret_cast<const float4*>(ab + ep.adain_off_w) + j);
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(ab + ep.adain_off_b) + j);
                    v[4 * j] = fmaf(v[4 * j], w4.x, b4.x); v[4 * j + 1] = fmaf(v[4 * j + 1], w4.y, b4.y);
                    v[4 * j + 2] = fmaf(v[4 * j + 2], w4.z, b4.z); v[4 * j + 3] = fmaf(v[4 * j + 3], w4.w, b4.w);
                }
            }
        } else if (EPI == IINS_EPI_LN) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int i = 0; i < 4; ++i) v[4 * j + i] = (v[4 * j + i] - ln_mean) * ln_rs;
                if (xh_dst != nullptr) xh_dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
#pragma unroll
                for (int i = 0; i < 4; ++i) v[4 * j + i] = fmaf(v[4 * j + i], __ldg(ep.gamma + gn + 4 * j + i), __ldg(ep.beta + gn + 4 * j + i));
            }
        }
        if (EPI == IINS_EPI_NBWD) {
            // v (+ residual operand) = gradient w.r.t. the previous layer's output; that layer's norm backward follows here
#pragma unroll
            for (int j = 0; j < 4; ++j) { v[4 * j] += a4[j].x; v[4 * j + 1] += a4[j].y; v[4 * j + 2] += a4[j].z; v[4 * j + 3] += a4[j].w; }
            if (ep.y != nullptr && row_ok) {
                float4* dst = reinterpret_cast<float4*>(ep.y + orow + c0);
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            constexpr float invL = 1.0f / (float)L;
            const bool relu = ep.nb_act == IINS_ACT_RELU;
            const long sb = (long)(row_ok ? b : 0);
            const bool lead = row_ok && l == 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 r4 = __ldg(reinterpret_cast<const float4*>(ep.nb_rstd + sb * p.N + gn) + j);
                float4 w4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ep.nb_adain != nullptr) {
                    w4 = __ldg(reinterpret_cast<const float4*>(ep.nb_adain + sb * ep.nb_ld + ep.nb_off_w + gn) + j);
                    b4 = __ldg(reinterpret_cast<const float4*>(ep.nb_adain + sb * ep.nb_ld + ep.nb_off_b + gn) + j);
                }
                const float4 xq = x4[EPI == IINS_EPI_NBWD ? j : 0];
                const float xv[4] = {xq.x, xq.y, xq.z, xq.w}, rv[4] = {r4.x, r4.y, r4.z, r4.w};
                const float wv[4] = {w4.x, w4.y, w4.z, w4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
                float sr[4], srx[4], o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float u = fmaf(xv[i], wv[i], bv[i]);
                    const float raw = (relu && !(u > 0.f)) ? 0.f : v[4 * j + i];
                    sr[i] = iins_lanes_sum<LL>(raw);                 // sum_l raw       (= d bias of AdaIN)
                    srx[i] = iins_lanes_sum<LL>(raw * xv[i]);        // sum_l raw * xhat (= d weight of AdaIN)
                    // gx = raw * scale with scale constant over l: mean_l(gx) = scale * sr / L, mean_l(gx * xhat) = scale * srx / L
                    o[i] = rv[i] * wv[i] * (raw - sr[i] * invL - xv[i] * srx[i] * invL);
                }
                if (row_ok) reinterpret_cast<float4*>(ep.nb_dz + orow + c0)[j] = make_float4(o[0], o[1], o[2], o[3]);
                if (lead && ep.nb_dadain != nullptr) {
                    *reinterpret_cast<float4*>(ep.nb_dadain + sb * ep.nb_ld + ep.nb_off_b + gn + 4 * j) = make_float4(sr[0], sr[1], sr[2], sr[3]);
                    *reinterpret_cast<float4*>(ep.nb_dadain + sb * ep.nb_ld + ep.nb_off_w + gn + 4 * j) = make_float4(srx[0], srx[1], srx[2], srx[3]);
                }
            }
            continue;
        }
        iins_act_vec<16>(v, ep.act, ep.slope);
        if (row_ok) {
            float4* dst = reinterpret_cast<float4*>(ep.y + orow + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                dst[j] = make_float4(v[4 * j] + a4[j].x, v[4 * j + 1] + a4[j].y, v[4 * j + 2] + a4[j].z, v[4 * j + 3] + a4[j].w);
        }
    }
}


// --------------------------------------------------------------------------------- forward / dgrad GEMM
struct IinsTCParams {
    IinsNTParams nt;
    const uint16_t* wpack;
    const uint16_t* wpack_odd;   // AKIND 2: packed weights of the odd-position class (blockIdx.z = 1)
    int pieces;          // 3 (fp32-grade) or 1 (bf16)
    int nkb;             // K blocks of 32
    int lin_dz;          // 1: the A operand of this AKIND 0 launch is dz of a Linear layer (nt.x = dy, row stride nt.g.Cin): apply nt.dz's mask / scale
};

// 288 threads: warps 0-7 are PRODUCERS (gather / split / store the A tile, later the epilogue), warp 8 is the
// MMA warp (TMA for the weight tile, tcgen05.mma issue, commit).  Hand-off through mbarriers only -- there is no
// CTA-wide barrier inside the K loop:
//   full[s]   (count 256)  producers -> MMA warp : stage s holds the A tile of this K block
//   bready[s] (tx bytes)   TMA       -> MMA warp : stage s holds the weight tile
//   done[s]   (tcgen05.commit) MMA   -> everyone : the MMAs reading stage s have completed (stage reusable)
template <int NT, int PIECES, int AKIND, int EPI, int LL>
static __global__ void __launch_bounds__(288, 2) iins_tc_nt_kernel(const IinsTCParams tp) {
    constexpr int BM = 128;
    // A stage: [chunk of 8 k][row][16 B] per piece.  The chunk stride (the descriptor's LBO) is padded by 64 bytes so that
    // the 8-byte stores of a warp (4 rows x 8 quads) spread over all banks: 2 wavefronts for 256 bytes, the minimum.
    constexpr uint32_t A_LBO = BM * 16 + 64;
    constexpr uint32_t A_PIECE = 4 * A_LBO;              // 8448 B
    constexpr uint32_t B_TILE = 4 * PIECES * NT * 16;    // [chunk][piece * NT + n][16 B]
    constexpr uint32_t STAGE = 3 * A_PIECE + 3 * 4 * NT * 16;
    constexpr int TCOLS = IinsTmemCols<NT, PIECES>::value;
    constexpr int LD = NT + 1;
    extern __shared__ __align__(1024) unsigned char dsm[];
    __shared__ __align__(8) unsigned long long mbar_done[2];
    __shared__ __align__(8) unsigned long long mbar_b[2];
    __shared__ __align__(8) unsigned long long mbar_full[2];
    __shared__ uint32_t tmem_slot;
    const IinsNTParams& p = tp.nt;
    const IinsGeom& g = p.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile_m = blockIdx.x * BM, n0 = blockIdx.y * NT;
    float* Cs = reinterpret_cast<float*>(dsm);                         // aliases the stages after the MMAs
    float* st_mean = reinterpret_cast<float*>(dsm + 2 * STAGE);
    float* st_rstd = st_mean + 1024;
    const int nkb = tp.nkb;

    // ---- producer set-up (index arithmetic only: no global access before iins_pdl_wait())
    const bool is_prod = warp < 8;
    const int cs = p.cshift;
    // FORWARD gathers (AKIND 0): producer warp w owns rows 16w .. 16w+15 of the tile; instruction jj covers rows
    // 16w + 4jj + (lane >> 3) and the 8 lanes of a row take the 8 quads of the K block (coalesced: 4 rows x 128 B per
    // warp-level load).  DATA-GRADIENT gathers (AKIND 1) keep one row per lane and two 8-wide chunks per thread: their
    // per-row candidate arithmetic (reflected / strided taps) would be done four times per thread otherwise, and that
    // kernel is bound by instruction issue, not by L1 wavefronts (measured both ways).
    const int a_quad = lane & 7;
    const int a_row0 = AKIND == 0 ? (warp & 7) * 16 + (lane >> 3) : (tid & 127);      // AKIND 1 / 2: one row per lane
    const int a_half = (tid >> 7) & 1;
    int a_b[4], a_l[4];
    bool a_ok[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        const int grow = tile_m + a_row0 + (AKIND == 0 ? 4 * jj : 0);
        a_ok[jj] = is_prod && grow < p.M;
        a_b[jj] = a_ok[jj] ? grow >> p.lshift : 0;
        a_l[jj] = a_ok[jj] ? grow & (p.Lrow - 1) : 0;
    }
    // raw operand data of TWO K blocks ahead lives in registers (the loads of block kb+2 are issued right after
    // block kb is handed to the MMA warp), so a K block costs its convert/store work, not a global-load latency
    float4 raw[2][4];
    // a Linear layer (one tap, one position per sample) reads a plain row-major matrix: row pointers are set up once and a K block
    // costs an add per load (the generic gather spends ~45 % of this kernel's instructions on tap / padding arithmetic there: ncu)
    const bool a_lin = AKIND == 0 && g.ks == 1 && g.Lin == 1 && g.Lout == 1 && g.stride == 1 && g.pad == 0 && g.in_layout == IINS_NLC;
    const float* a_ptr[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) a_ptr[jj] = p.x + (long)(a_ok[jj] ? tile_m + a_row0 + 4 * jj : 0) * g.Cin;
    auto load_raw = [&](int kb, float4* dst) {
        // the host routes layers without 16-byte gathers (< 8 channels, NCL operand) to the SIMT kernels
        if (AKIND == 0) {
            const int k0 = kb * 32 + a_quad * 4;
            if (a_lin) {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) dst[jj] = iins_ld4(a_ptr[jj] + k0, a_ok[jj] && k0 < p.K);
                if (tp.lin_dz) {
                    // data gradient of a Linear layer through the forward mapping (4 rows x 128 B per warp-level load instead of
                    // 32 rows x 16 B):  dz = dy * act'(y) * scale
                    if (p.dz.y != nullptr && p.dz.act != IINS_ACT_NONE) {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const float4 y = iins_ld4(p.dz.y + (a_ptr[jj] - p.x) + k0, a_ok[jj] && k0 < p.K);
                            dst[jj].x *= iins_dact_from_y(y.x, p.dz.act, p.dz.slope); dst[jj].y *= iins_dact_from_y(y.y, p.dz.act, p.dz.slope);
                            dst[jj].z *= iins_dact_from_y(y.z, p.dz.act, p.dz.slope); dst[jj].w *= iins_dact_from_y(y.w, p.dz.act, p.dz.slope);
                        }
                    }
                    if (p.dz.dy_scale != 1.f) {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) { dst[jj].x *= p.dz.dy_scale; dst[jj].y *= p.dz.dy_scale; dst[jj].z *= p.dz.dy_scale; dst[jj].w *= p.dz.dy_scale; }
                    }
                }
            } else {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                    dst[jj] = a_ok[jj] ? iins_gather4_fwd(g, p.x, p.K, cs, a_b[jj], a_l[jj], k0) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else if (AKIND == 2) {
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int k0 = kb * 32 + (a_half * 2 + jj) * 8;
                float v[8];
                if (!a_ok[0]) iins_zero8(v);
                else iins_gather8_dgrad_parity(g, p.dz, p.K, cs, a_b[0], a_l[0], (int)blockIdx.z, k0, v);
                dst[2 * jj] = make_float4(v[0], v[1], v[2], v[3]);
                dst[2 * jj + 1] = make_float4(v[4], v[5], v[6], v[7]);
            }
        } else {
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int k0 = kb * 32 + (a_half * 2 + jj) * 8;
                float v[8];
                if (!a_ok[0]) iins_zero8(v);
                else iins_gather8_dgrad_fast(g, p.dz, p.K, cs, a_b[0], a_l[0], k0, v);
                dst[2 * jj] = make_float4(v[0], v[1], v[2], v[3]);
                dst[2 * jj + 1] = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
    };
    iins_pdl_launch_dependents();
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            umma::mbar_init(umma::smem_u32(&mbar_done[i]), 1);
            umma::mbar_init(umma::smem_u32(&mbar_b[i]), 1);
            umma::mbar_init(umma::smem_u32(&mbar_full[i]), 256);
        }
        umma::fence_mbar_init();
    }
    if (warp == 8) umma::tmem_alloc(umma::smem_u32(&tmem_slot), TCOLS);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    // everything above overlapped the tail of the previous kernel in the stream; its results are needed from here on
    iins_pdl_wait();
    if (is_prod) {
        load_raw(0, raw[0]);
        if (nkb > 1) load_raw(1, raw[1]);
    }

    if (warp == 8) {
        // ------------------------------------------------------------------ MMA warp (warp-uniform code)
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb & 1;
            unsigned char* sA = dsm + s * STAGE;
            unsigned char* sB = sA + 3 * A_PIECE;
            if (kb >= 2) umma::mbar_wait(umma::smem_u32(&mbar_done[s]), (uint32_t)(((kb >> 1) - 1) & 1));
            const bool leader = umma::elect_one();
            if (leader) {
                const unsigned char* src = reinterpret_cast<const unsigned char*>(AKIND == 2 && blockIdx.z == 1 ? tp.wpack_odd : tp.wpack) +
                                           ((long)blockIdx.y * nkb + kb) * B_TILE;
                umma::mbar_arrive_expect_tx(umma::smem_u32(&mbar_b[s]), B_TILE);
                umma::tma_bulk_g2s(umma::smem_u32(sB), src, B_TILE, umma::smem_u32(&mbar_b[s]));
            }
            umma::mbar_wait(umma::smem_u32(&mbar_full[s]), (uint32_t)((kb >> 1) & 1));
            umma::mbar_wait(umma::smem_u32(&mbar_b[s]), (uint32_t)((kb >> 1) & 1));
            umma::tc_fence_after();
            // K-major, no swizzle: LBO = distance between 8-wide k chunks, SBO = 128 B between 8-row groups
            const uint64_t ad = umma::make_desc(umma::smem_u32(sA), A_LBO, 128);
            const uint64_t bd = umma::make_desc(umma::smem_u32(sB), PIECES * NT * 16, 128);
            iins_issue_kstep<NT, PIECES, 0, 0>(tmem, ad, bd, A_PIECE >> 4, leader, kb > 0 ? 1u : 0u);
            iins_issue_kstep<NT, PIECES, 0, 0>(tmem, ad + ((2 * A_LBO) >> 4), bd + ((2 * PIECES * NT * 16) >> 4), A_PIECE >> 4,
                                               leader, 1u);
            if (leader) umma::commit(umma::smem_u32(&mbar_done[s]));
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ producers
        auto produce = [&](int kb, float4* src) {
            const int s = kb & 1;
            unsigned char* sA = dsm + s * STAGE;
            if (kb >= 2) umma::mbar_wait(umma::smem_u32(&mbar_done[s]), (uint32_t)(((kb >> 1) - 1) & 1));
            if (AKIND == 0) {
                unsigned char* dst = sA + (a_quad >> 1) * A_LBO + a_row0 * 16 + (a_quad & 1) * 8;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) iins_store4_split(src[jj], dst + jj * 64, A_PIECE, PIECES);
            } else {
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const float v[8] = {src[2 * jj].x, src[2 * jj].y, src[2 * jj].z, src[2 * jj].w,
                                        src[2 * jj + 1].x, src[2 * jj + 1].y, src[2 * jj + 1].z, src[2 * jj + 1].w};
                    iins_store8_split(v, sA + (a_half * 2 + jj) * A_LBO + a_row0 * 16, A_PIECE, PIECES);
                }
            }
            umma::fence_async_smem();
            umma::mbar_arrive(umma::smem_u32(&mbar_full[s]));
            if (kb + 2 < nkb) load_raw(kb + 2, src);      // in flight while the next block is converted
        };
        for (int kb = 0; kb < nkb; kb += 2) {
            produce(kb, raw[0]);
            if (kb + 1 < nkb) produce(kb + 1, raw[1]);
        }
    }
    // every thread observes the completion of the last MMAs (also orders the smem reuse by the epilogue)
    if (nkb >= 2) umma::mbar_wait(umma::smem_u32(&mbar_done[(nkb - 2) & 1]), (uint32_t)(((nkb - 2) >> 1) & 1));
    umma::mbar_wait(umma::smem_u32(&mbar_done[(nkb - 1) & 1]), (uint32_t)(((nkb - 1) >> 1) & 1));
    umma::tc_fence_after();

    if (EPI != IINS_EPI_SMEM) {
        if (warp < 8) iins_tc_epilogue_regs<NT, PIECES, EPI, LL>(p, tmem, tile_m, n0, warp, lane, st_mean, AKIND == 2 ? 2 : 1,
                                                                 AKIND == 2 ? (int)blockIdx.z : 0);
    } else if (warp < 8) {
        // ---- generic path: TMEM -> SMEM (+ bias).  Warp w owns TMEM lanes 32*(w&3) .. +31; with NT >= 32 the two
        // warps sharing a lane quarter split the columns.
        constexpr int COLS_PER_WARP = NT >= 32 ? NT / 2 : NT;
        const int q = warp & 3, hf = warp >> 2;
        if (NT >= 32 || hf == 0) {
            const int row = q * 32 + lane;
            const int cbeg = NT >= 32 ? hf * COLS_PER_WARP : 0;
#pragma unroll
            for (int c0 = 0; c0 < COLS_PER_WARP; c0 += 16) {
                float v[16];
                iins_tmem_acc16<NT, PIECES>(tmem + ((uint32_t)(q * 32) << 16), cbeg + c0, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    int n = n0 + cbeg + c0 + i;
                    float bv = (p.ep.bias != nullptr && n < p.N) ? __ldg(p.ep.bias + n) : 0.f;
                    Cs[row * LD + cbeg + c0 + i] = v[i] + bv;
                }
            }
        }
        umma::tc_fence_before();
        iins_epi_sync<true>();
        iins_epilogue_tile<NT, LD, true>(p, Cs, st_mean, st_rstd, tile_m, n0);
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 8) umma::tmem_dealloc(tmem, TCOLS);
}

// ------------------------------------------------------------------------------------------- weight grad
struct IinsTCTNParams {
    IinsTNParams tn;
    int pieces;
    int K;               // ks * Cin
    int lshift;          // log2(Lout) (rows per sample)
    int cshift_in;       // log2(Cin) or -1
    int cshift_out;      // log2(Cout) or -1
    // Batched launch: `nbatch` > 0 problems of IDENTICAL geometry (the 2 * n_residual trunk convolutions, whose dz tensors
    // the fused backward kernel leaves behind all at once) in ONE grid: blockIdx.z selects the problem (Cout <= NT then),
    // the operand / output pointers come from these arrays instead of tn.x / tn.dz.dy / tn.dw / tn.db.
    int nbatch;
    const float* bx[8]; const float* bdy[8]; float* bdw[8]; float* bdb[8];
};

// grid = (row parts, ceil(K/128), ceil(Cout/NT)).  D^T[k][n] accumulated in TMEM (128 lanes = 128 k entries).
template <int NT, int PIECES>
static __global__ void __launch_bounds__(288, 2) iins_tc_tn_kernel(const IinsTCTNParams tp) {
    iins_pdl_launch_dependents();
    constexpr int BR = 32;                               // rows per stage (2 MMA k-steps of 16)
    constexpr uint32_t A_PIECE = 16 * BR * 16;           // [k group of 8][row][16 B] = 8192 B
    constexpr uint32_t B_PIECE = (NT / 8) * BR * 16;     // [n group of 8][row][16 B]; pieces stacked = more n groups
    constexpr uint32_t STAGE = 3 * A_PIECE + 3 * B_PIECE;
    constexpr int TCOLS = IinsTmemCols<NT, PIECES>::value;
    extern __shared__ __align__(1024) unsigned char dsm[];
    __shared__ __align__(8) unsigned long long mbar_done[2];
    __shared__ __align__(8) unsigned long long mbar_full[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float s_bias[64];
    const IinsTNParams& p = tp.tn;
    const IinsGeom& g = p.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool batched = tp.nbatch > 0;
    const int ktile0 = blockIdx.y * 128, n0 = batched ? 0 : blockIdx.z * NT;
    const float* const px = batched ? tp.bx[blockIdx.z] : p.x;
    IinsDz pdz = p.dz;
    if (batched) pdz.dy = tp.bdy[blockIdx.z];
    float* const pdw = batched ? tp.bdw[blockIdx.z] : p.dw;
    float* const pdb = batched ? tp.bdb[blockIdx.z] : p.db;
    const int r_begin = blockIdx.x * p.rows_per_part;
    int r_end = r_begin + p.rows_per_part;
    if (r_end > p.M) r_end = p.M;
    const int nit = (r_end - r_begin + BR - 1) / BR;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            umma::mbar_init(umma::smem_u32(&mbar_done[i]), 1);
            umma::mbar_init(umma::smem_u32(&mbar_full[i]), 256);
        }
        umma::fence_mbar_init();
    }
    if (tid < 64) s_bias[tid] = 0.f;
    if (warp == 8) umma::tmem_alloc(umma::smem_u32(&tmem_slot), TCOLS);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    iins_pdl_wait();                                   // prologue overlapped the previous kernel's tail
    const bool do_bias = pdb != nullptr && blockIdx.y == 0;
    const bool has_z = warp < NT / 8;
    float bsum[8];
    iins_zero8(bsum);

    if (warp == 8) {
        for (int it = 0; it < nit; ++it) {
            const int s = it & 1;
            unsigned char* sA = dsm + s * STAGE;
            unsigned char* sB = sA + 3 * A_PIECE;
            umma::mbar_wait(umma::smem_u32(&mbar_full[s]), (uint32_t)((it >> 1) & 1));
            umma::tc_fence_after();
            // MN-major: LBO = distance between 8-row groups (128 B), SBO = distance between MN groups of 8
            const uint64_t ad = umma::make_desc(umma::smem_u32(sA), 128, BR * 16);
            const uint64_t bd = umma::make_desc(umma::smem_u32(sB), 128, BR * 16);
            const bool leader = umma::elect_one();
            iins_issue_kstep<NT, PIECES, 1, 1>(tmem, ad, bd, A_PIECE >> 4, leader, it > 0 ? 1u : 0u);
            iins_issue_kstep<NT, PIECES, 1, 1>(tmem, ad + (256 >> 4), bd + (256 >> 4), A_PIECE >> 4, leader, 1u);
            if (leader) umma::commit(umma::smem_u32(&mbar_done[s]));
            __syncwarp();
        }
    } else {
        // two row blocks of raw operand data in flight in registers (see the forward kernel)
        float rawa[2][2][8], rawz[2][8];
        auto load_raw = [&](int it, float (*ra)[8], float* rz) {
            const int row = r_begin + it * BR + lane;
            const bool ok = row < r_end;
            const int b = ok ? row >> tp.lshift : 0;
            const int l = ok ? row & (g.Lout - 1) : 0;
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int k0 = ktile0 + (warp + 8 * jj) * 8;
                if (!ok || k0 >= tp.K) iins_zero8(ra[jj]);
                else iins_gather8_fwd_fast(g, px, tp.K, tp.cshift_in, b, l, k0, ra[jj]);
            }
            if (has_z) {
                if (!ok) iins_zero8(rz);
                else iins_dz8_fast(g, pdz, b, l, n0 + warp * 8, true, rz);
            }
        };
        auto produce = [&](int it, float (*ra)[8], float* rz) {
            const int s = it & 1;
            unsigned char* sA = dsm + s * STAGE;
            unsigned char* sB = sA + 3 * A_PIECE;
            if (it >= 2) umma::mbar_wait(umma::smem_u32(&mbar_done[s]), (uint32_t)(((it >> 1) - 1) & 1));
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                // k chunks beyond K (a half-empty last k tile) are all zero: written once per stage, then skipped (warp-uniform)
                if (ktile0 + (warp + 8 * jj) * 8 < tp.K || it < 2)
                    iins_store8_split(ra[jj], sA + ((warp + 8 * jj) * BR + lane) * 16, A_PIECE, PIECES);
            }
            if (has_z) {
                if (do_bias) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) bsum[i] += rz[i];
                }
                iins_store8_split(rz, sB + (warp * BR + lane) * 16, B_PIECE, PIECES);
            }
            umma::fence_async_smem();
            umma::mbar_arrive(umma::smem_u32(&mbar_full[s]));
            if (it + 2 < nit) load_raw(it + 2, ra, rz);
        };
        if (nit > 0) load_raw(0, rawa[0], rawz[0]);
        if (nit > 1) load_raw(1, rawa[1], rawz[1]);
        for (int it = 0; it < nit; it += 2) {
            produce(it, rawa[0], rawz[0]);
            if (it + 1 < nit) produce(it + 1, rawa[1], rawz[1]);
        }
    }
    if (nit >= 2) umma::mbar_wait(umma::smem_u32(&mbar_done[(nit - 2) & 1]), (uint32_t)(((nit - 2) >> 1) & 1));
    if (nit >= 1) umma::mbar_wait(umma::smem_u32(&mbar_done[(nit - 1) & 1]), (uint32_t)(((nit - 1) >> 1) & 1));
    umma::tc_fence_after();

    if (nit >= 1 && warp < 8) {
        // TMEM lane = k entry of this tile, column = n.  Warp w owns lanes 32*(w&3)..; column halves as above.
        constexpr int COLS_PER_WARP = NT >= 32 ? NT / 2 : NT;
        const int q = warp & 3, hf = warp >> 2;
        if (NT >= 32 || hf == 0) {
            const int k = ktile0 + q * 32 + lane;
            const int cbeg = NT >= 32 ? hf * COLS_PER_WARP : 0;
            const bool kok = k < tp.K;
            const int t = kok ? k / g.Cin : 0, ci = kok ? k - t * g.Cin : 0;
            // dW[n][ci][t]: this lane's (ci, t) is fixed, n walks the accumulator columns with a constant stride
            const long nstride = (long)g.Cin * g.ks;
            float* dst0 = pdw + (long)(n0 + cbeg) * nstride + (long)ci * g.ks + t;
#pragma unroll
            for (int c0 = 0; c0 < COLS_PER_WARP; c0 += 16) {
                float v[16];
                iins_tmem_acc16<NT, PIECES>(tmem + ((uint32_t)(q * 32) << 16), cbeg + c0, v);
                if (kok) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if (n0 + cbeg + c0 + i < g.Cout) atomicAdd(dst0 + (c0 + i) * nstride, v[i]);
                    }
                }
            }
        }
        if (do_bias && has_z) {
            // bias gradient: each producer warp owns 8 distinct output channels -> warp-shuffle sum over its 32 rows, plain store
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float tot = iins_warp_sum(bsum[i]);
                if (lane == 0) s_bias[warp * 8 + i] = tot;
            }
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (do_bias && tid < NT && n0 + tid < g.Cout && nit >= 1) atomicAdd(pdb + n0 + tid, s_bias[tid]);
    if (warp == 8) umma::tmem_dealloc(tmem, TCOLS);
}

#endif  // !IINS_CPUSIM
