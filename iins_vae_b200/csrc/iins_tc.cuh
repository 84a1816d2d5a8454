// Tensor-core (tcgen05 / TMEM) implicit-GEMM kernels for sm_100a.
//
//   iins_tc_nt_kernel : forward + data-gradient GEMMs.  A tile (128 rows x 32 k) is gathered from the
//       channels-last activations (im2col on the fly: reflection / zero padding, stride, nearest x2
//       upsampling, or the transposed tap map for dgrad), split into bf16 pieces and written to shared
//       memory in the UMMA K-major no-swizzle layout; the weight tile arrives pre-split through a TMA bulk
//       copy (cp.async.bulk + mbarrier); one elected lane of warp 0 issues tcgen05.mma (M=128, N=16..64,
//       K=16), the fp32 accumulator lives in TMEM, the epilogue (bias, InstanceNorm / AdaIN / LayerNorm,
//       activation, residual) runs on the TMEM -> SMEM staged tile.
//   iins_tc_tn_kernel : weight gradient  dW^T[k][n] = sum_rows A[row][k] * dz[row][n]  with both operands
//       MN-major (the reduction runs over rows), accumulated in TMEM over the CTA's row range and flushed
//       with atomics.
//
// Precision: fp32 parity needs more than TF32/BF16 single-pass products, so each fp32 operand is split into
// three bf16 pieces (8+8+8 mantissa bits) and the six significant piece products are accumulated in fp32
// ("bf16x3": error ~2^-23, same class as an fp32 FMA chain).  pieces == 1 is the plain bf16 mode.
//
// Latency notes (measured with the IINS_TL clock trace, tools/timeline.py): the MMA issue path must be
// warp-uniform (descriptors in uniform registers; a divergent `if (tid == 0)` costs ~170 cycles per MMA in
// R2UR traffic), the raw operand data of K block kb+1 is prefetched into registers before the barrier of
// block kb, and the row / k index splits are shifts (every L and channel count on the path is a power of 2).
#pragma once
#include "iins_gemm.cuh"
#ifndef IINS_CPUSIM
#include "iins_umma.cuh"

// generic (scalar) gathers for layouts / channel counts the 16-byte fast paths do not cover; kept out of line
__device__ __noinline__ void iins_gather8_fwd_generic(const IinsGeom& g, const float* __restrict__ x, int K, int b, int l, int k0, float* v) {
    int t = k0 / g.Cin, c = k0 - t * g.Cin;
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        v[i] = (k0 + i < K) ? iins_a_fwd(g, x, b, l, t, c) : 0.f;
        if (++c == g.Cin) { c = 0; ++t; }
    }
}
__device__ __noinline__ void iins_dz8_generic(const IinsGeom& g, const IinsDz& d, int b, int l, int n0, float* v) {
#pragma unroll 1
    for (int i = 0; i < 8; ++i) v[i] = (n0 + i < g.Cout) ? iins_dz_at(g, d, b, l, n0 + i) : 0.f;
}
__device__ __noinline__ void iins_gather8_dgrad_generic(const IinsGeom& g, const IinsDz& d, int K, int b, int pos, int k0, float* v) {
    int t = k0 / g.Cout, c = k0 - t * g.Cout;
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        v[i] = (k0 + i < K) ? iins_a_dgrad(g, d, b, pos, t, c) : 0.f;
        if (++c == g.Cout) { c = 0; ++t; }
    }
}

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

// forward A operand, fast path: Cin = 1 << cs (>= 8), NLC input; 8 consecutive k = one tap, 8 channels
__device__ __forceinline__ void iins_gather8_fwd_fast(const IinsGeom& g, const float* __restrict__ x, int K, int cs, int b, int l,
                                                      int k0, float* v) {
    const int t = k0 >> cs, c0 = k0 & (g.Cin - 1);
    const int pos = iins_src_pos(g, l, t);
    const bool ok = k0 < K && pos >= 0;
    iins_ld8(x + (((long)b * g.Lin + (ok ? pos : 0)) << cs) + c0, ok, v);
}

// 8 consecutive output channels of dz at (b,l), fast path: Cout = 1 << cs (>= 8), NLC
__device__ __forceinline__ void iins_dz8_fast(const IinsGeom& g, const IinsDz& d, int cs, int b, int l, int n0, bool ok, float* v) {
    const long idx = (((long)b * g.Lout + (ok ? l : 0)) << cs) + n0;
    iins_ld8(d.dy_bcast ? d.dy + ((long)b << cs) + n0 : d.dy + idx, ok, v);
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
    const int t = k0 >> cs, c0 = k0 & (g.Cout - 1);
    int q0 = pos + g.pad, q1 = -1, q2 = -1;
    if (g.mode == IINS_PAD_REFLECT) {
        if (pos >= 1 && pos <= g.pad) q1 = g.pad - pos;
        if (pos <= g.Lin - 2 && pos >= g.Lin - 1 - g.pad) q2 = g.pad + 2 * (g.Lin - 1) - pos;
    } else if (g.mode == IINS_PAD_UP2) {
        q0 = 2 * pos + g.pad;
        q1 = q0 + 1;
    }
    const int sh = g.stride - 1;                       // stride is 1 or 2
    iins_zero8(v);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int q = j == 0 ? q0 : (j == 1 ? q1 : q2);
        const int r = q - t;
        const int l = r >> sh;
        const bool ok = k0 < K && q >= 0 && r >= 0 && (r & sh) == 0 && l < g.Lout;
        if (j > 0 && q < 0) continue;                  // warp-divergent only at the padded borders
        float u[8];
        iins_dz8_fast(g, d, cs, b, l, c0, ok, u);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += u[i];
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

__global__ void __launch_bounds__(256) iins_pack_all_kernel(const IinsPackAllParams pp) {
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
                long wi = q.kind == 0 ? ((long)n * q.Cin + c) * q.ks + t : ((long)c * q.Cin + n) * q.ks + t;
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

#define IINS_TL(tag) do { if (tl_ != nullptr && tl_n_ < 500) { tl_[2 * tl_n_] = (tag); tl_[2 * tl_n_ + 1] = clock64(); ++tl_n_; tl_[1022] = tl_n_; } } while (0)

// --------------------------------------------------------------------------------- forward / dgrad GEMM
struct IinsTCParams {
    IinsNTParams nt;
    const uint16_t* wpack;
    int pieces;          // 3 (fp32-grade) or 1 (bf16)
    int nkb;             // K blocks of 32
    long long* timeline; // debug: (tag, clock64) pairs of CTA (0,0) thread 0, or nullptr
};

// 288 threads: warps 0-7 are PRODUCERS (gather / split / store the A tile, later the epilogue), warp 8 is the
// MMA warp (TMA for the weight tile, tcgen05.mma issue, commit).  Hand-off through mbarriers only -- there is no
// CTA-wide barrier inside the K loop:
//   full[s]   (count 256)  producers -> MMA warp : stage s holds the A tile of this K block
//   bready[s] (tx bytes)   TMA       -> MMA warp : stage s holds the weight tile
//   done[s]   (tcgen05.commit) MMA   -> everyone : the MMAs reading stage s have completed (stage reusable)
template <int NT, int PIECES>
__global__ void __launch_bounds__(288) iins_tc_nt_kernel(const IinsTCParams tp) {
    constexpr int BM = 128;
    constexpr uint32_t A_PIECE = 4 * BM * 16;            // 8192 B : [chunk][row][16 B]
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
    long long* tl_ = (tp.timeline != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) ? tp.timeline : nullptr;
    int tl_n_ = 0;
    IINS_TL(0);

    if (warp == 8) {
        // ------------------------------------------------------------------ MMA warp (warp-uniform code)
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb & 1;
            unsigned char* sA = dsm + s * STAGE;
            unsigned char* sB = sA + 3 * A_PIECE;
            if (kb >= 2) umma::mbar_wait(umma::smem_u32(&mbar_done[s]), (uint32_t)(((kb >> 1) - 1) & 1));
            const bool leader = umma::elect_one();
            if (leader) {
                const unsigned char* src = reinterpret_cast<const unsigned char*>(tp.wpack) + ((long)blockIdx.y * nkb + kb) * B_TILE;
                umma::mbar_arrive_expect_tx(umma::smem_u32(&mbar_b[s]), B_TILE);
                umma::tma_bulk_g2s(umma::smem_u32(sB), src, B_TILE, umma::smem_u32(&mbar_b[s]));
            }
            umma::mbar_wait(umma::smem_u32(&mbar_full[s]), (uint32_t)((kb >> 1) & 1));
            umma::mbar_wait(umma::smem_u32(&mbar_b[s]), (uint32_t)((kb >> 1) & 1));
            umma::tc_fence_after();
            // K-major, no swizzle: LBO = distance between 8-wide k chunks, SBO = 128 B between 8-row groups
            const uint64_t ad = umma::make_desc(umma::smem_u32(sA), BM * 16, 128);
            const uint64_t bd = umma::make_desc(umma::smem_u32(sB), PIECES * NT * 16, 128);
            iins_issue_kstep<NT, PIECES, 0, 0>(tmem, ad, bd, A_PIECE >> 4, leader, kb > 0 ? 1u : 0u);
            iins_issue_kstep<NT, PIECES, 0, 0>(tmem, ad + ((2 * BM * 16) >> 4), bd + ((2 * PIECES * NT * 16) >> 4), A_PIECE >> 4,
                                               leader, 1u);
            if (leader) umma::commit(umma::smem_u32(&mbar_done[s]));
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ producers
        const int a_row = tid & 127, a_half = tid >> 7;
        const int grow = tile_m + a_row;
        const bool a_ok = grow < p.M;
        const int a_b = a_ok ? grow >> p.lshift : 0;
        const int a_l = a_ok ? grow & (p.Lrow - 1) : 0;
        const int cs = p.cshift;
        // 16-byte gathers need >= 8 channels (power of two) in channels-last order
        const bool fast = cs >= 3 && (p.a_kind == 0 ? g.in_layout == IINS_NLC : g.out_layout == IINS_NLC);
        float raw[2][8];
        auto load_raw = [&](int kb) {
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int k0 = kb * 32 + (a_half * 2 + jj) * 8;
                if (!a_ok) iins_zero8(raw[jj]);
                else if (fast) {
                    if (p.a_kind == 0) iins_gather8_fwd_fast(g, p.x, p.K, cs, a_b, a_l, k0, raw[jj]);
                    else iins_gather8_dgrad_fast(g, p.dz, p.K, cs, a_b, a_l, k0, raw[jj]);
                } else {
                    float tmp[8];                          // address-taken copy: keeps raw[][] in registers
                    if (p.a_kind == 0) iins_gather8_fwd_generic(g, p.x, p.K, a_b, a_l, k0, tmp);
                    else iins_gather8_dgrad_generic(g, p.dz, p.K, a_b, a_l, k0, tmp);
#pragma unroll
                    for (int i = 0; i < 8; ++i) raw[jj][i] = tmp[i];
                }
            }
        };
        load_raw(0);
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb & 1;
            unsigned char* sA = dsm + s * STAGE;
            if (kb >= 2) umma::mbar_wait(umma::smem_u32(&mbar_done[s]), (uint32_t)(((kb >> 1) - 1) & 1));
            IINS_TL(1);
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
                iins_store8_split(raw[jj], sA + ((a_half * 2 + jj) * BM + a_row) * 16, A_PIECE, PIECES);
            umma::fence_async_smem();
            umma::mbar_arrive(umma::smem_u32(&mbar_full[s]));
            IINS_TL(2);
            if (kb + 1 < nkb) load_raw(kb + 1);           // in flight while the MMA warp works on this block
            IINS_TL(3);
        }
    }
    // every thread observes the completion of the last MMAs (also orders the smem reuse by the epilogue)
    if (nkb >= 2) umma::mbar_wait(umma::smem_u32(&mbar_done[(nkb - 2) & 1]), (uint32_t)(((nkb - 2) >> 1) & 1));
    umma::mbar_wait(umma::smem_u32(&mbar_done[(nkb - 1) & 1]), (uint32_t)(((nkb - 1) >> 1) & 1));
    umma::tc_fence_after();
    IINS_TL(13);

    if (warp < 8) {
        // ---- TMEM -> SMEM (+ bias).  Warp w owns TMEM lanes 32*(w&3) .. +31; with NT >= 32 the two warps
        // sharing a lane quarter split the columns.
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
        IINS_TL(14);
        iins_epilogue_tile<NT, LD, true>(p, Cs, st_mean, st_rstd, tile_m, n0);
        IINS_TL(15);
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
};

// grid = (row parts, ceil(K/128), ceil(Cout/NT)).  D^T[k][n] accumulated in TMEM (128 lanes = 128 k entries).
template <int NT, int PIECES>
__global__ void __launch_bounds__(288) iins_tc_tn_kernel(const IinsTCTNParams tp) {
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
    const int ktile0 = blockIdx.y * 128, n0 = blockIdx.z * NT;
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
    const bool do_bias = p.db != nullptr && blockIdx.y == 0;
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
        const bool fast_a = tp.cshift_in >= 3 && g.in_layout == IINS_NLC;
        const bool fast_z = tp.cshift_out >= 3 && g.out_layout == IINS_NLC;
        float rawa[2][8], rawz[8];
        auto load_raw = [&](int it) {
            const int row = r_begin + it * BR + lane;
            const bool ok = row < r_end;
            const int b = ok ? row >> tp.lshift : 0;
            const int l = ok ? row & (g.Lout - 1) : 0;
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int k0 = ktile0 + (warp + 8 * jj) * 8;
                if (!ok) iins_zero8(rawa[jj]);
                else if (fast_a) iins_gather8_fwd_fast(g, p.x, tp.K, tp.cshift_in, b, l, k0, rawa[jj]);
                else {
                    float tmp[8];
                    iins_gather8_fwd_generic(g, p.x, tp.K, b, l, k0, tmp);
#pragma unroll
                    for (int i = 0; i < 8; ++i) rawa[jj][i] = tmp[i];
                }
            }
            if (has_z) {
                if (!ok) iins_zero8(rawz);
                else if (fast_z) iins_dz8_fast(g, p.dz, tp.cshift_out, b, l, n0 + warp * 8, true, rawz);
                else {
                    float tmp[8];
                    iins_dz8_generic(g, p.dz, b, l, n0 + warp * 8, tmp);
#pragma unroll
                    for (int i = 0; i < 8; ++i) rawz[i] = tmp[i];
                }
            }
        };
        if (nit > 0) load_raw(0);
        for (int it = 0; it < nit; ++it) {
            const int s = it & 1;
            unsigned char* sA = dsm + s * STAGE;
            unsigned char* sB = sA + 3 * A_PIECE;
            if (it >= 2) umma::mbar_wait(umma::smem_u32(&mbar_done[s]), (uint32_t)(((it >> 1) - 1) & 1));
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
                iins_store8_split(rawa[jj], sA + ((warp + 8 * jj) * BR + lane) * 16, A_PIECE, PIECES);
            if (has_z) {
                if (do_bias) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) bsum[i] += rawz[i];
                }
                iins_store8_split(rawz, sB + (warp * BR + lane) * 16, B_PIECE, PIECES);
            }
            umma::fence_async_smem();
            umma::mbar_arrive(umma::smem_u32(&mbar_full[s]));
            if (it + 1 < nit) load_raw(it + 1);
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
#pragma unroll
            for (int c0 = 0; c0 < COLS_PER_WARP; c0 += 16) {
                float v[16];
                iins_tmem_acc16<NT, PIECES>(tmem + ((uint32_t)(q * 32) << 16), cbeg + c0, v);
                if (kok) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        int n = n0 + cbeg + c0 + i;
                        if (n < g.Cout) atomicAdd(p.dw + iins_w_index(g, n, ci, t), v[i]);
                    }
                }
            }
        }
        if (do_bias && has_z) {
#pragma unroll
            for (int i = 0; i < 8; ++i) atomicAdd(&s_bias[warp * 8 + i], bsum[i]);
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (do_bias && tid < NT && n0 + tid < g.Cout && nit >= 1) atomicAdd(p.db + n0 + tid, s_bias[tid]);
    if (warp == 8) umma::tmem_dealloc(tmem, TCOLS);
}

#endif  // !IINS_CPUSIM
