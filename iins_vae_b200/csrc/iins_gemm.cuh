// Implicit-GEMM kernels (FP32 SIMT core) shared by every Conv1d / Linear layer of the path.
//
//   nt_kernel : C[row][n] = sum_k A[row][k] * Bw[n][k]        forward (A = im2col gather of the layer
//               input) and data-gradient (A = gather of dz through the transposed tap map)
//   tn_kernel : dW[n][k] += sum_row dz[row][n] * A[row][k]    weight gradient (+ bias gradient)
//
// Tiles are 128 rows = a whole number of samples (every L on the path is a power of two <= 128), so
// InstanceNorm / AdaIN / the reference's custom LayerNorm, the activation and the residual add run in
// the epilogue on an SMEM-staged tile and the activation tensor is written exactly once.
#pragma once
#include "iins_common.cuh"

struct IinsEpilogue {
    const float* bias;         // [N] or nullptr
    int norm;                  // IINS_NORM_*
    int act;
    float slope;
    const float* add;          // same indexing as y; added after norm+act (residual / grad accumulate)
    float* y;                  // output
    float* xhat;               // normalised pre-affine values (saved for backward) or nullptr
    float* rstd;               // IN/AdaIN: [B*N]; LN: [B] (= 1/(std+eps))
    const float* gamma;        // LN per-channel scale / shift
    const float* beta;
    const float* adain;        // AdaIN params (B, adain_ld): bias at +off_b, weight at +off_w
    int adain_ld, adain_off_b, adain_off_w;
    // Fused InstanceNorm / AdaIN BACKWARD of the layer whose output gradient this data-gradient GEMM produces (the
    // tile holds whole samples, so the per-(sample, channel) sums over L are available in the epilogue):
    //   dz = rstd * (gx - mean_l(gx) - xhat * mean_l(gx * xhat)),  gx = act'(.) * dy * scale
    // nb_dz != nullptr switches it on; y (the plain gradient) is then optional (nullptr = not needed afterwards).
    float* nb_dz;
    const float* nb_xhat;      // saved normalised values of that layer (same indexing as y)
    const float* nb_rstd;      // [B*N]
    int nb_act;                // IINS_ACT_NONE / IINS_ACT_RELU of that layer
    const float* nb_adain;     // AdaIN params of that layer or nullptr (plain InstanceNorm)
    float* nb_dadain;          // AdaIN parameter gradients (B, ld): bias grad at +off_b, weight grad at +off_w
    int nb_ld, nb_off_b, nb_off_w;
};

struct IinsNTParams {
    IinsGeom g;
    int a_kind;                // 0 = forward gather from x, 1 = dgrad gather from dz
    const float* x;            // forward input
    IinsDz dz;                 // dgrad source
    const float* w;
    IinsEpilogue ep;
    int M, N, K;               // GEMM sizes (rows, cols, reduction)
    int Lrow;                  // rows per sample of the OUTPUT of this GEMM (Lout fwd, Lin dgrad); a power of two
    int lshift;                // log2(Lrow)
    int cshift;                // log2(channel extent of the k index) or -1 if it is not a power of two
    int out_layout;            // layout of the GEMM output
    long long* dbg;            // debug clock trace (CTA 0 / thread 0) or nullptr
};

// Norm / activation / residual / store on one SMEM-staged tile Cs[128][LD] (bias already added).
// st_mean / st_rstd: scratch for the statistics (>= 1024 floats each).  Called by every thread of a
// 256-thread CTA; contains __syncthreads().
// 256-thread barrier used inside the epilogue: the whole CTA for the SIMT kernels; named barrier 1 for the
// warp-specialised tensor-core kernels (their 9th warp, the MMA issuer, does not take part in the epilogue)
template <bool NAMED>
__device__ __forceinline__ void iins_epi_sync() {
#ifndef IINS_CPUSIM
    if (NAMED) { asm volatile("bar.sync 1, 256;" ::: "memory"); return; }
#endif
    __syncthreads();
}

template <int BN, int LD, bool NAMED = false>
__device__ __forceinline__ void iins_epilogue_tile(const IinsNTParams& p, const float* Cs, float* st_mean, float* st_rstd,
                                                   int tile_m, int n0) {
    constexpr int BM = 128;
    const int tid = threadIdx.x;
    const IinsGeom& g = p.g;
    const IinsEpilogue& ep = p.ep;
    const int L = p.Lrow;
    const int S = BM / L;                  // samples per tile (L <= 128 always when norm != NONE)
    const int ncols = (p.N - n0) < BN ? (p.N - n0) : BN;
    const int b0 = tile_m / L;             // first sample of the tile

    long long* edbg = (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) ? p.dbg : nullptr;
    if (edbg) edbg[900] = clock64();
    if (ep.norm == IINS_NORM_IN || ep.norm == IINS_NORM_ADAIN) {
        // per (sample, channel) statistics over the L rows; biased variance (models.py:152, 1072)
        const int pairs = S * ncols;
        int G = 1;
        while (G < 32 && pairs * G * 2 <= 256) G <<= 1;
        const int per_iter = 256 / G;
        const float invL = 1.0f / (float)L;
        const bool pow2 = (ncols & (ncols - 1)) == 0;
        const int nsh = 31 - __clz(ncols);
        for (int p0 = 0; p0 < pairs; p0 += per_iter) {
            int pr = p0 + tid / G, sub = tid % G;
            bool ok = pr < pairs;
            int s = !ok ? 0 : (pow2 ? pr >> nsh : pr / ncols), c = ok ? pr - s * ncols : 0;
            float sum = 0.f;
            if (ok) for (int l = sub; l < L; l += G) sum += Cs[(s * L + l) * LD + c];
            sum = iins_group_sum(sum, G);
            float mean = sum * invL;
            float sq = 0.f;
            if (ok) for (int l = sub; l < L; l += G) { float dv = Cs[(s * L + l) * LD + c] - mean; sq += dv * dv; }
            sq = iins_group_sum(sq, G);
            if (ok && sub == 0) {
                // rsqrt + one Newton step: full fp32 accuracy without the IEEE sqrt / divide sequences
                const float vpe = fmaf(sq, invL, IINS_EPS);
                float rs = rsqrtf(vpe);
                rs = rs * fmaf(-0.5f * vpe, rs * rs, 1.5f);
                st_mean[s * BN + c] = mean;
                st_rstd[s * BN + c] = rs;
                int b = b0 + s;
                if (b < g.B && ep.rstd != nullptr) ep.rstd[(long)b * p.N + n0 + c] = rs;
            }
        }
        iins_epi_sync<NAMED>();
    } else if (ep.norm == IINS_NORM_LN) {
        // per-sample mean and UNBIASED std over (C*L), eps added to std (models.py:976-981)
        const int warp = tid >> 5, lane = tid & 31;
        const int nel = L * ncols;
        for (int s0 = 0; s0 < S; s0 += 8) {
            int s = s0 + warp;
            bool ok = s < S;
            float sum = 0.f;
            if (ok) for (int e = lane; e < nel; e += 32) sum += Cs[(s * L + e / ncols) * LD + (e % ncols)];
            sum = iins_warp_sum(sum);
            float mean = sum / (float)nel;
            float sq = 0.f;
            if (ok) for (int e = lane; e < nel; e += 32) { float dv = Cs[(s * L + e / ncols) * LD + (e % ncols)] - mean; sq += dv * dv; }
            sq = iins_warp_sum(sq);
            if (ok && lane == 0) {
                float rs = 1.0f / (sqrtf(sq / (float)(nel - 1)) + IINS_EPS);
                st_mean[s] = mean;
                st_rstd[s] = rs;
                int b = b0 + s;
                if (b < g.B && ep.rstd != nullptr) ep.rstd[b] = rs;
            }
        }
        iins_epi_sync<NAMED>();
    }

    if (edbg) edbg[901] = clock64();
    // ---- apply + store (coalesced over the contiguous NLC tile)
    // Each thread owns 4 consecutive columns and walks the rows with a fixed stride; L is a power of two on
    // this path (p.lshift), so the row -> (sample, position) split is a shift, and the NLC stores are 16 bytes.
    constexpr int CG = BN / 4;                 // column groups
    constexpr int RSTEP = 256 / CG;            // rows covered per pass
    const int cg = tid % CG, n = cg * 4, gn = n0 + n;
    const int lsh = p.lshift;
    const bool vec = p.out_layout == IINS_NLC && (p.N & 3) == 0 && gn + 3 < p.N;
    float gam[4], bet[4];
    if (ep.norm == IINS_NORM_LN) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            gam[j] = gn + j < p.N ? __ldg(ep.gamma + gn + j) : 0.f;
            bet[j] = gn + j < p.N ? __ldg(ep.beta + gn + j) : 0.f;
        }
    }
    if (gn < p.N) {
#pragma unroll 4
        for (int r = tid / CG; r < BM; r += RSTEP) {
            const int gr = tile_m + r;
            if (gr >= p.M) break;
            const int s = r >> lsh, b = gr >> lsh, l = gr & (L - 1);
            float v[4], xh[4];
            // residual / accumulate operand first (read-only path: the load may be hoisted over the stores of the
            // previous rows; every element is read before the same thread overwrites it when add == y)
            float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (vec && ep.add != nullptr) a4 = __ldg(reinterpret_cast<const float4*>(ep.add + (long)gr * p.N + gn));
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = Cs[r * LD + n + j];
            if (ep.norm == IINS_NORM_IN || ep.norm == IINS_NORM_ADAIN) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[j] = (v[j] - st_mean[s * BN + n + j]) * st_rstd[s * BN + n + j];
                    xh[j] = v[j];
                }
                if (ep.norm == IINS_NORM_ADAIN) {
                    const float* ab = ep.adain + (long)b * ep.adain_ld;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (gn + j < p.N) v[j] = fmaf(v[j], __ldg(ab + ep.adain_off_w + gn + j), __ldg(ab + ep.adain_off_b + gn + j));
                }
            } else if (ep.norm == IINS_NORM_LN) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[j] = (v[j] - st_mean[s]) * st_rstd[s];
                    xh[j] = v[j];
                    v[j] = fmaf(v[j], gam[j], bet[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = iins_act(v[j], ep.act, ep.slope);
            if (vec) {
                const long oi = (long)gr * p.N + gn;
                if (ep.norm != IINS_NORM_NONE && ep.xhat != nullptr)
                    *reinterpret_cast<float4*>(ep.xhat + oi) = make_float4(xh[0], xh[1], xh[2], xh[3]);
                v[0] += a4.x; v[1] += a4.y; v[2] += a4.z; v[3] += a4.w;
                *reinterpret_cast<float4*>(ep.y + oi) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (gn + j >= p.N) continue;
                    const long oi = p.out_layout == IINS_NCL ? (((long)b * p.N + gn + j) << lsh) + l : (long)gr * p.N + gn + j;
                    if (ep.norm != IINS_NORM_NONE && ep.xhat != nullptr) ep.xhat[oi] = xh[j];
                    float o = v[j];
                    if (ep.add != nullptr) o += ep.add[oi];
                    ep.y[oi] = o;
                }
            }
        }
    }
    if (edbg) edbg[902] = clock64();
}

template <int BN>
static __global__ void __launch_bounds__(256) iins_nt_kernel(const IinsNTParams p) {
    iins_pdl_enter();
    constexpr int BM = 128, BK = 16, TN = 4;
    constexpr int CT = BN / TN;            // threads along n
    constexpr int RT = 256 / CT;           // threads along m
    constexpr int TM = BM / RT;            // rows per thread
    constexpr int B_PER_THREAD = (BN * BK + 255) / 256;
    __shared__ __align__(16) float smem[2048 + BK * BN + BM * BN];
    float* As = smem;                       // [BK][BM]
    float* Bs = smem + 2048;                // [BK][BN]
    float* Cs = smem + 2048 + BK * BN;      // [BM][BN]
    float* st_mean = smem;                  // epilogue stats alias As (<= 1024 entries each)
    float* st_rstd = smem + 1024;

    const int tid = threadIdx.x;
    const int tile_m = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const IinsGeom& g = p.g;
    const int Cdim = p.a_kind == 0 ? g.Cin : g.Cout;      // channel extent of the k index

    // ---- A loader: row = tid & 127, 8 consecutive k starting at (tid >> 7) * 8
    const int a_row = tid & 127;
    const int a_k0 = (tid >> 7) * 8;
    const int grow = tile_m + a_row;
    const bool a_ok = grow < p.M;
    const int a_b = a_ok ? grow / p.Lrow : 0;
    const int a_l = a_ok ? grow - a_b * p.Lrow : 0;

    float a_reg[8];
    float b_reg[B_PER_THREAD];

    auto load_regs = [&](int kb) {
        int k = kb * BK + a_k0;
        int t = k / Cdim;
        int c = k - t * Cdim;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v = 0.f;
            if (a_ok && k + i < p.K) {
                v = p.a_kind == 0 ? iins_a_fwd(g, p.x, a_b, a_l, t, c) : iins_a_dgrad(g, p.dz, a_b, a_l, t, c);
            }
            a_reg[i] = v;
            if (++c == Cdim) { c = 0; ++t; }
        }
#pragma unroll
        for (int j = 0; j < B_PER_THREAD; ++j) {
            int e = tid + j * 256;
            float v = 0.f;
            if (e < BN * BK) {
                int kk = e / BN, n = e - kk * BN;
                int kg = kb * BK + kk, ng = n0 + n;
                if (kg < p.K && ng < p.N) {
                    int tt = kg / Cdim, cc = kg - tt * Cdim;
                    long wi = p.a_kind == 0 ? iins_w_index(g, ng, cc, tt) : iins_w_index(g, cc, ng, tt);
                    v = __ldg(p.w + wi);
                }
            }
            b_reg[j] = v;
        }
    };
    auto store_smem = [&]() {
#pragma unroll
        for (int i = 0; i < 8; ++i) As[(a_k0 + i) * BM + a_row] = a_reg[i];
#pragma unroll
        for (int j = 0; j < B_PER_THREAD; ++j) {
            int e = tid + j * 256;
            if (e < BN * BK) Bs[e] = b_reg[j];          // e == kk*BN + n
        }
    };

    const int tx = tid % CT, ty = tid / CT;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int nkb = (p.K + BK - 1) / BK;
    load_regs(0);
    store_smem();
    __syncthreads();
    for (int kb = 0; kb < nkb; ++kb) {
        if (kb + 1 < nkb) load_regs(kb + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[kk * BM + ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[kk * BN + tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
        if (kb + 1 < nkb) {
            store_smem();
            __syncthreads();
        }
    }

    // ---- epilogue: stage (acc + bias) in smem
    const IinsEpilogue& ep = p.ep;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        int n = n0 + tx * TN + j;
        float bv = (ep.bias != nullptr && n < p.N) ? __ldg(ep.bias + n) : 0.f;
#pragma unroll
        for (int i = 0; i < TM; ++i) Cs[(ty * TM + i) * BN + tx * TN + j] = acc[i][j] + bv;
    }
    __syncthreads();

    iins_epilogue_tile<BN, BN>(p, Cs, st_mean, st_rstd, tile_m, n0);
}

// ------------------------------------------------------------------------------------ wgrad
struct IinsTNParams {
    IinsGeom g;
    const float* x;            // forward input of the layer (A operand via the forward gather)
    IinsDz dz;
    float* dw;                 // [Cout][Cin][ks], accumulated with atomicAdd
    float* db;                 // [Cout] or nullptr
    int M;                     // rows = B * Lout
    int rows_per_part;         // multiple of 32
};

// grid = (row parts, ceil(K/64), ceil(N/64)); each CTA owns a 64(n) x 64(k) block of dW, reduces its
// row range in registers (4x4 per thread) and flushes once with atomicAdd.
static __global__ void __launch_bounds__(256) iins_tn_kernel(const IinsTNParams p) {
    iins_pdl_enter();
    constexpr int BR = 32, BNK = 64;
    __shared__ __align__(16) float Ds[BR * BNK];
    __shared__ __align__(16) float Xs[BR * BNK];
    const IinsGeom& g = p.g;
    const int tid = threadIdx.x;
    const int N = g.Cout, K = g.ks * g.Cin;
    const int k0 = blockIdx.y * BNK, n0 = blockIdx.z * BNK;
    const int r_begin = blockIdx.x * p.rows_per_part;
    int r_end = r_begin + p.rows_per_part;
    if (r_end > p.M) r_end = p.M;
    const int tx = tid & 15, ty = tid >> 4;       // tx -> 4 k, ty -> 4 n
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float bsum = 0.f;

    for (int r0 = r_begin; r0 < r_end; r0 += BR) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int e = tid + j * 256;
            int r = e >> 6, c = e & 63;
            int row = r0 + r;
            float dv = 0.f, xv = 0.f;
            if (row < r_end) {
                int b = row / g.Lout, l = row - b * g.Lout;
                if (n0 + c < N) dv = iins_dz_at(g, p.dz, b, l, n0 + c);
                int k = k0 + c;
                if (k < K) { int t = k / g.Cin, ci = k - t * g.Cin; xv = iins_a_fwd(g, p.x, b, l, t, ci); }
            }
            Ds[e] = dv;
            Xs[e] = xv;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < BR; ++r) {
            float d[4], x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) d[i] = Ds[r * BNK + ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = Xs[r * BNK + tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(d[i], x[j], acc[i][j]);
        }
        if (p.db != nullptr && blockIdx.y == 0 && tid < BNK) {
            for (int r = 0; r < BR; ++r) bsum += Ds[r * BNK + tid];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int n = n0 + ty * 4 + i;
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = k0 + tx * 4 + j;
            if (k >= K) continue;
            int t = k / g.Cin, ci = k - t * g.Cin;
            atomicAdd(p.dw + iins_w_index(g, n, ci, t), acc[i][j]);
        }
    }
    if (p.db != nullptr && blockIdx.y == 0 && tid < BNK && n0 + tid < N) atomicAdd(p.db + n0 + tid, bsum);
}

// ===================================================================== small-channel layers (SIMT, direct)
// Layers with N <= 16 output columns and K <= 64 (the stems 1->4 / 1->16 k7, 4->8 k4, the last upsampling
// conv 8->4 k5, the output conv 4->1 k7, their data gradients, the 2-channel range-code convs, the tiny
// classifier): a 128x16 tensor-core tile would be almost all padding and these layers are pure HBM streams at
// L = 64..128, so they run as a direct convolution, one thread per output row, weights in shared memory,
// followed by the SAME fused tile epilogue (norm / activation / residual).
// acc[0..7] += a * w[0..7]  (w 32-byte aligned in shared memory: two 16-byte loads)
IINS_D void iins_row_fma8(float* acc, float a, const float* w) {
    const float4 w0 = *reinterpret_cast<const float4*>(w), w1 = *reinterpret_cast<const float4*>(w + 4);
    acc[0] = fmaf(a, w0.x, acc[0]); acc[1] = fmaf(a, w0.y, acc[1]); acc[2] = fmaf(a, w0.z, acc[2]); acc[3] = fmaf(a, w0.w, acc[3]);
    acc[4] = fmaf(a, w1.x, acc[4]); acc[5] = fmaf(a, w1.y, acc[5]); acc[6] = fmaf(a, w1.z, acc[6]); acc[7] = fmaf(a, w1.w, acc[7]);
}

struct IinsRowParams {
    IinsNTParams nt;           // geometry, operands, epilogue, M/N/K, Lrow, lshift
};

static __global__ void __launch_bounds__(256) iins_row_nt_kernel(const IinsRowParams rp) {
    iins_pdl_enter();
    constexpr int BM = 128, NT = 16, LD = NT + 1, KMAX = 64;
    __shared__ __align__(16) float Ws[KMAX * NT];            // [k][n]
    __shared__ float Cs[BM * LD];
    __shared__ float st_mean[1024];
    __shared__ float st_rstd[1024];
    const IinsNTParams& p = rp.nt;
    const IinsGeom& g = p.g;
    const int tid = threadIdx.x;
    const int tile_m = blockIdx.x * BM;
    const int Cdim = p.a_kind == 0 ? g.Cin : g.Cout;
    // weights -> smem, Ws[k][n] with k = t * Cdim + c
    for (int e = tid; e < p.K * NT; e += 256) {     // rows k >= K are never read
        int k = e / NT, n = e - k * NT;
        float v = 0.f;
        if (n < p.N) {
            int t = k / Cdim, c = k - t * Cdim;
            v = __ldg(p.w + (p.a_kind == 0 ? iins_w_index(g, n, c, t) : iins_w_index(g, c, n, t)));
        }
        Ws[e] = v;
    }
    __syncthreads();
    const int row = tid & 127, half = tid >> 7;            // two threads per row: columns [8*half, 8*half+8)
    const int grow = tile_m + row;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    // With N <= 8 the second thread of a row would idle: instead the two threads split the TAPS and the
    // partial sums are combined through shared memory (Cs) below.
    const bool split_taps = p.N <= 8;
    const int ncol0 = split_taps ? 0 : half * 8;
    if (grow < p.M && (split_taps || half * 8 < p.N)) {
        const int b = grow >> p.lshift, l = grow & (p.Lrow - 1);
        const int tstep = split_taps ? 2 : 1;
        for (int t = split_taps ? half : 0; t < g.ks; t += tstep) {
            const float* wr0 = Ws + (t * Cdim) * NT + ncol0;
            if (p.a_kind == 0) {
                const int pos = iins_src_pos(g, l, t);
                if (pos < 0) continue;
                const float* xr = p.x + iins_in_index(g, b, pos, 0);
                if (g.in_layout == IINS_NLC && (Cdim & 3) == 0) {
                    for (int c = 0; c < Cdim; c += 4) {
                        const float4 a4 = __ldg(reinterpret_cast<const float4*>(xr + c));
                        iins_row_fma8(acc, a4.x, wr0 + c * NT);
                        iins_row_fma8(acc, a4.y, wr0 + (c + 1) * NT);
                        iins_row_fma8(acc, a4.z, wr0 + (c + 2) * NT);
                        iins_row_fma8(acc, a4.w, wr0 + (c + 3) * NT);
                    }
                } else {
                    const long cstride = g.in_layout == IINS_NCL ? g.Lin : 1;
                    for (int c = 0; c < Cdim; ++c) iins_row_fma8(acc, __ldg(xr + c * cstride), wr0 + c * NT);
                }
            } else {
                // data gradient: output rows whose tap t reads this input position (<= 3 candidates)
                int q[3] = {l + g.pad, -1, -1};
                if (g.mode == IINS_PAD_REFLECT) {
                    if (l >= 1 && l <= g.pad) q[1] = g.pad - l;
                    if (l <= g.Lin - 2 && l >= g.Lin - 1 - g.pad) q[2] = g.pad + 2 * (g.Lin - 1) - l;
                } else if (g.mode == IINS_PAD_UP2) {
                    q[0] = 2 * l + g.pad;
                    q[1] = q[0] + 1;
                }
                for (int jq = 0; jq < 3; ++jq) {
                    if (q[jq] < 0) continue;
                    const int r = q[jq] - t;
                    if (r < 0) continue;
                    const int lo = r / g.stride;
                    if (lo * g.stride != r || lo >= g.Lout) continue;
                    if (g.out_layout == IINS_NLC && (Cdim & 3) == 0 && !p.dz.dy_bcast) {
                        const long base = ((long)b * g.Lout + lo) * g.Cout;
                        for (int c = 0; c < Cdim; c += 4) {
                            float4 a4 = __ldg(reinterpret_cast<const float4*>(p.dz.dy + base + c));
                            if (p.dz.y != nullptr && p.dz.act != IINS_ACT_NONE) {
                                const float4 y4 = __ldg(reinterpret_cast<const float4*>(p.dz.y + base + c));
                                a4.x *= iins_dact_from_y(y4.x, p.dz.act, p.dz.slope); a4.y *= iins_dact_from_y(y4.y, p.dz.act, p.dz.slope);
                                a4.z *= iins_dact_from_y(y4.z, p.dz.act, p.dz.slope); a4.w *= iins_dact_from_y(y4.w, p.dz.act, p.dz.slope);
                            }
                            const float sc = p.dz.dy_scale;
                            iins_row_fma8(acc, a4.x * sc, wr0 + c * NT);
                            iins_row_fma8(acc, a4.y * sc, wr0 + (c + 1) * NT);
                            iins_row_fma8(acc, a4.z * sc, wr0 + (c + 2) * NT);
                            iins_row_fma8(acc, a4.w * sc, wr0 + (c + 3) * NT);
                        }
                    } else {
                        for (int c = 0; c < Cdim; ++c) iins_row_fma8(acc, iins_dz_at(g, p.dz, b, lo, c), wr0 + c * NT);
                    }
                }
            }
        }
    }
    if (split_taps) {
        if (half == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) Cs[row * LD + 8 + j] = acc[j];      // columns 8..15 are unused when N <= 8
        }
        __syncthreads();
        if (half == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += Cs[row * LD + 8 + j];
        }
        __syncthreads();
    }
    if (!split_taps || half == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int n = ncol0 + j;
            float bv = (p.ep.bias != nullptr && n < p.N) ? __ldg(p.ep.bias + n) : 0.f;
            Cs[row * LD + n] = acc[j] + bv;
        }
    }
    __syncthreads();
    iins_epilogue_tile<NT, LD>(p, Cs, st_mean, st_rstd, tile_m, 0);
}

// ---- one thread per output row ("row2") --------------------------------------------------------------------
// The small-channel layers are HBM streams (L = 32..128 positions, <= 16 channels) with next to no arithmetic.  One thread
// owns one GEMM row (or R rows) and all NACC (4 / 8 / 16) output channels in registers: a tap/channel step is NACC/4
// 16-byte weight loads from shared memory + NACC FMAs per row, the InstanceNorm / AdaIN / LayerNorm statistics are
// warp-shuffle sums (+ one shared-memory exchange between the warps of a sample), and the row leaves the registers as
// 16-byte stores -- no staging of the tile in shared memory.  128 threads = 128 R rows = whole samples (L <= 128).
// What limits them (measured, profiles/r02f_row_pair.txt) is the shared-memory pipe of those weight reads: see iins_row2_fma.
// Preconditions (host): N <= NACC, K <= IINS_ROW2_KMAX(NACC); with a fused norm: L in {32, 64, 128}.
#define IINS_ROW2_WMAX 1024            // floats of weights in shared memory: K * NACC <= 1024

// acc[r][0..NACC) += a[r] * w[0..NACC) for the R rows of a thread.  ONE 16-byte shared-memory read per four weights serves all R
// rows: a warp-uniform LDS.128 still occupies the shared-memory pipe for four passes, and with one row per thread that pipe
// (ncu: l1tex throughput 60-76 %, short-scoreboard stalls) -- not instruction issue, not HBM -- bounds these kernels.  The FMAs
// are packed (FFMA2: two IEEE fp32 FMAs per issue slot; bit-identical to scalar fmaf).
template <int NACC, int R>
IINS_D void iins_row2_fma(float (*acc)[NACC], const float* a, const float* w) {
#pragma unroll
    for (int j = 0; j < NACC; j += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(w + j);
#pragma unroll
        for (int r = 0; r < R; ++r) {
#if defined(__CUDA_ARCH__) && !defined(IINS_CPUSIM)
            unsigned long long a2;
            asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(a[r]));
            asm("{\n\t.reg .b64 rw, rc;\n\tmov.b64 rw, {%2, %3};\n\tmov.b64 rc, {%0, %1};\n\tfma.rn.f32x2 rc, %4, rw, rc;\n\t"
                "mov.b64 {%0, %1}, rc;\n\t}" : "+f"(acc[r][j]), "+f"(acc[r][j + 1]) : "f"(w4.x), "f"(w4.y), "l"(a2));
            asm("{\n\t.reg .b64 rw, rc;\n\tmov.b64 rw, {%2, %3};\n\tmov.b64 rc, {%0, %1};\n\tfma.rn.f32x2 rc, %4, rw, rc;\n\t"
                "mov.b64 {%0, %1}, rc;\n\t}" : "+f"(acc[r][j + 2]), "+f"(acc[r][j + 3]) : "f"(w4.z), "f"(w4.w), "l"(a2));
#else
            acc[r][j] = fmaf(a[r], w4.x, acc[r][j]); acc[r][j + 1] = fmaf(a[r], w4.y, acc[r][j + 1]);
            acc[r][j + 2] = fmaf(a[r], w4.z, acc[r][j + 2]); acc[r][j + 3] = fmaf(a[r], w4.w, acc[r][j + 3]);
#endif
        }
    }
}

// sum of v over the rows of this thread's sample that live in nw whole warps (nw = 1, 2, 4); xch: [4] floats per call site
IINS_D float iins_row2_sample_sum(float v, int nw, int warp, int lane, float* xch) {
    v = iins_warp_sum(v);
    if (nw > 1) {                                   // CTA-uniform
        if (lane == 0) xch[warp] = v;
        __syncthreads();
        const int w0 = warp & ~(nw - 1);
        float t = 0.f;
        for (int i = 0; i < nw; ++i) t += xch[w0 + i];
        v = t;
        __syncthreads();                            // xch is reused by the next reduction
    }
    return v;
}

// EPI: 0 plain, 1 InstanceNorm / AdaIN, 2 LayerNorm, 3 data gradient + fused IN backward.
// R rows per thread (1 or 2).  The R rows of a thread are H rows apart: H = 128 for the plain epilogue (a CTA owns 128 R
// consecutive rows), H = L / R with a fused norm, so that the L rows of a sample are the R rows of H consecutive lanes and the
// statistics stay warp-shuffle sums (R = 2 needs L in {64, 128}; M is a whole number of samples there).
template <int NACC, int AKIND, int EPI, int R>
static __global__ void __launch_bounds__(128) iins_row2_nt_kernel(const IinsRowParams rp) {
    iins_pdl_enter();
    constexpr int BM = 128;
    __shared__ __align__(16) float Ws[IINS_ROW2_WMAX];       // [k][NACC]
    __shared__ float s_bias[NACC < 16 ? 16 : NACC];
    __shared__ float xch[4];
    const IinsNTParams& p = rp.nt;
    const IinsGeom& g = p.g;
    const IinsEpilogue& ep = p.ep;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_m = blockIdx.x * BM * R;
    const int n0 = blockIdx.y * NACC;                    // column block (only the wide plain layers use more than one)
    const int Cdim = AKIND == 0 ? g.Cin : g.Cout;
    for (int e = tid; e < p.K * NACC; e += BM) {
        const int k = e / NACC, n = n0 + e - k * NACC;
        float v = 0.f;
        if (n < p.N) {
            const int t = k / Cdim, c = k - t * Cdim;
            v = __ldg(p.w + (AKIND == 0 ? iins_w_index(g, n, c, t) : iins_w_index(g, c, n, t)));
        }
        Ws[e] = v;
    }
    if (tid < NACC) s_bias[tid] = (ep.bias != nullptr && n0 + tid < p.N) ? __ldg(ep.bias + n0 + tid) : 0.f;
    __syncthreads();

    const int L = p.Lrow;
    const int hs = (R == 1 || EPI == 0) ? 7 : p.lshift - (R == 2 ? 1 : 0);      // log2 H
    int grow[R], b[R], l[R];
    bool ok[R], any_ok = false;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        grow[r] = tile_m + (((tid >> hs) * R + r) << hs) + (tid & ((1 << hs) - 1));
        ok[r] = grow[r] < p.M;
        any_ok |= ok[r];
        b[r] = ok[r] ? grow[r] >> p.lshift : 0;
        l[r] = ok[r] ? grow[r] & (L - 1) : 0;
    }
    float acc[R][NACC];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < NACC; ++j) acc[r][j] = s_bias[j];
    if (any_ok) {
        if (AKIND == 0) {
            // (index arithmetic hoisted out of the tap loop)
            const bool ncl = g.in_layout == IINS_NCL;
            const int pstride = ncl ? 1 : g.Cin;
            const int hi2 = 2 * (g.Lin - 1);
            const float* xb[R];
            int u0[R];
#pragma unroll
            for (int r = 0; r < R; ++r) { xb[r] = p.x + (long)b[r] * g.Lin * g.Cin; u0[r] = l[r] * g.stride - g.pad; }
            for (int t = 0; t < g.ks; ++t) {
                const float* xr[R];
                bool any = false;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    int pos = u0[r] + t;
                    if (g.mode == IINS_PAD_REFLECT) { pos = pos < 0 ? -pos : pos; const int rf = hi2 - pos; pos = rf < pos ? rf : pos; }   // pad < Lin
                    else if (g.mode == IINS_PAD_UP2) pos = (pos < 0 || pos >= 2 * g.Lin) ? -1 : (pos >> 1);
                    else if (pos >= g.Lin) pos = -1;
                    xr[r] = (pos >= 0 && ok[r]) ? xb[r] + pos * pstride : nullptr;
                    any |= xr[r] != nullptr;
                }
                if (!any) continue;
                const float* wr = Ws + t * Cdim * NACC;
                float a[R];
                if (Cdim == 1) {
#pragma unroll
                    for (int r = 0; r < R; ++r) a[r] = xr[r] != nullptr ? __ldg(xr[r]) : 0.f;
                    iins_row2_fma<NACC, R>(acc, a, wr);
                } else if (g.in_layout == IINS_NLC && (Cdim & 3) == 0) {
                    for (int c = 0; c < Cdim; c += 4) {
                        float4 a4[R];
#pragma unroll
                        for (int r = 0; r < R; ++r)
                            a4[r] = xr[r] != nullptr ? __ldg(reinterpret_cast<const float4*>(xr[r] + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int r = 0; r < R; ++r) a[r] = a4[r].x;
                        iins_row2_fma<NACC, R>(acc, a, wr + c * NACC);
#pragma unroll
                        for (int r = 0; r < R; ++r) a[r] = a4[r].y;
                        iins_row2_fma<NACC, R>(acc, a, wr + (c + 1) * NACC);
#pragma unroll
                        for (int r = 0; r < R; ++r) a[r] = a4[r].z;
                        iins_row2_fma<NACC, R>(acc, a, wr + (c + 2) * NACC);
#pragma unroll
                        for (int r = 0; r < R; ++r) a[r] = a4[r].w;
                        iins_row2_fma<NACC, R>(acc, a, wr + (c + 3) * NACC);
                    }
                } else {
                    const long cstride = g.in_layout == IINS_NCL ? g.Lin : 1;
                    for (int c = 0; c < Cdim; ++c) {
#pragma unroll
                        for (int r = 0; r < R; ++r) a[r] = xr[r] != nullptr ? __ldg(xr[r] + c * cstride) : 0.f;
                        iins_row2_fma<NACC, R>(acc, a, wr + c * NACC);
                    }
                }
            }
        } else {
            // data gradient: output rows whose tap t reads this input position (<= 3 candidates)
            int q[R][3];
            long zb0[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                q[r][0] = l[r] + g.pad; q[r][1] = -1; q[r][2] = -1;
                if (g.mode == IINS_PAD_REFLECT) {
                    if (l[r] >= 1 && l[r] <= g.pad) q[r][1] = g.pad - l[r];
                    if (l[r] <= g.Lin - 2 && l[r] >= g.Lin - 1 - g.pad) q[r][2] = g.pad + 2 * (g.Lin - 1) - l[r];
                } else if (g.mode == IINS_PAD_UP2) {
                    q[r][0] = 2 * l[r] + g.pad;
                    q[r][1] = q[r][0] + 1;
                }
                if (!ok[r]) { q[r][0] = -1; q[r][1] = -1; q[r][2] = -1; }
                zb0[r] = (long)b[r] * g.Lout * g.Cout;
            }
            const bool vec = g.out_layout == IINS_NLC && (Cdim & 3) == 0 && !p.dz.dy_bcast;
            const bool masked = p.dz.y != nullptr && p.dz.act != IINS_ACT_NONE;
            const float sc = p.dz.dy_scale;
            for (int t = 0; t < g.ks; ++t) {
                const float* wr = Ws + t * Cdim * NACC;
#pragma unroll
                for (int jq = 0; jq < 3; ++jq) {
                    int lo[R];
                    bool any = false;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        lo[r] = -1;
                        const int rr = q[r][jq] - t;
                        if (q[r][jq] >= 0 && rr >= 0) {
                            const int o = g.stride == 1 ? rr : (g.stride == 2 ? rr >> 1 : rr / g.stride);     // (uniform branches: no integer division)
                            if (o * g.stride == rr && o < g.Lout) lo[r] = o;
                        }
                        any |= lo[r] >= 0;
                    }
                    if (!any) continue;
                    float a[R];
                    if (vec) {
                        for (int c = 0; c < Cdim; c += 4) {
                            float4 a4[R];
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                a4[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (lo[r] >= 0) {
                                    const long base = zb0[r] + (long)lo[r] * g.Cout + c;
                                    a4[r] = __ldg(reinterpret_cast<const float4*>(p.dz.dy + base));
                                    if (masked) {
                                        const float4 y4 = __ldg(reinterpret_cast<const float4*>(p.dz.y + base));
                                        a4[r].x *= iins_dact_from_y(y4.x, p.dz.act, p.dz.slope); a4[r].y *= iins_dact_from_y(y4.y, p.dz.act, p.dz.slope);
                                        a4[r].z *= iins_dact_from_y(y4.z, p.dz.act, p.dz.slope); a4[r].w *= iins_dact_from_y(y4.w, p.dz.act, p.dz.slope);
                                    }
                                }
                            }
#pragma unroll
                            for (int r = 0; r < R; ++r) a[r] = a4[r].x * sc;
                            iins_row2_fma<NACC, R>(acc, a, wr + c * NACC);
#pragma unroll
                            for (int r = 0; r < R; ++r) a[r] = a4[r].y * sc;
                            iins_row2_fma<NACC, R>(acc, a, wr + (c + 1) * NACC);
#pragma unroll
                            for (int r = 0; r < R; ++r) a[r] = a4[r].z * sc;
                            iins_row2_fma<NACC, R>(acc, a, wr + (c + 2) * NACC);
#pragma unroll
                            for (int r = 0; r < R; ++r) a[r] = a4[r].w * sc;
                            iins_row2_fma<NACC, R>(acc, a, wr + (c + 3) * NACC);
                        }
                    } else {
                        for (int c = 0; c < Cdim; ++c) {
#pragma unroll
                            for (int r = 0; r < R; ++r) a[r] = lo[r] >= 0 ? iins_dz_at(g, p.dz, b[r], lo[r], c) : 0.f;
                            iins_row2_fma<NACC, R>(acc, a, wr + c * NACC);
                        }
                    }
                }
            }
        }
    }

    // ---- epilogue in registers
    const int nw = EPI == 0 ? 1 : (L / R) >> 5;          // warps that hold one sample (1, 2 or 4) when a norm is fused
    float xh[R][NACC];                                   // normalised pre-affine values (dead when EPI == 0)
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < NACC; ++j) xh[r][j] = 0.f;
    if constexpr (EPI == 1) {
        const float invL = 1.0f / (float)L;
        // all NACC columns share ONE shared-memory exchange per statistic (mean, then centred sum of squares)
        __shared__ __align__(16) float xchv[4][NACC];
        float tot[NACC];
        auto sample_sums = [&](float* v) {               // v[j] -> sum of v[j] over the rows of this thread's sample held by other lanes
            const float mine = iins_warp_sums<NACC>(v, lane);        // total of column iins_warp_sums_slot(lane) over this warp
            if ((lane & (32 / NACC - 1)) == 0) xchv[warp][iins_warp_sums_slot<NACC>(lane)] = mine;
            if (nw > 1) __syncthreads(); else __syncwarp();          // nw is CTA-uniform
            const int w0 = warp & ~(nw - 1);
#pragma unroll
            for (int j = 0; j < NACC; j += 4) {
                float4 t = *reinterpret_cast<const float4*>(&xchv[w0][j]);
                if (nw > 1) {
                    const float4 u = *reinterpret_cast<const float4*>(&xchv[w0 + 1][j]);
                    t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
                }
                if (nw == 4) {
                    const float4 u = *reinterpret_cast<const float4*>(&xchv[w0 + 2][j]);
                    const float4 c = *reinterpret_cast<const float4*>(&xchv[w0 + 3][j]);
                    t.x += u.x + c.x; t.y += u.y + c.y; t.z += u.z + c.z; t.w += u.w + c.w;
                }
                v[j] = t.x; v[j + 1] = t.y; v[j + 2] = t.z; v[j + 3] = t.w;
            }
            if (nw > 1) __syncthreads(); else __syncwarp();          // xchv is reused by the next statistic
        };
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
            tot[j] = acc[0][j];
#pragma unroll
            for (int r = 1; r < R; ++r) tot[j] += acc[r][j];
        }
        sample_sums(tot);
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
            const float mean = tot[j] * invL;
            tot[j] = 0.f;
#pragma unroll
            for (int r = 0; r < R; ++r) { xh[r][j] = acc[r][j] - mean; tot[j] = r == 0 ? xh[r][j] * xh[r][j] : fmaf(xh[r][j], xh[r][j], tot[j]); }
        }
        sample_sums(tot);
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
            const float vpe = fmaf(tot[j], invL, IINS_EPS);
            float rs = rsqrtf(vpe);
            rs = rs * fmaf(-0.5f * vpe, rs * rs, 1.5f);
            if (ok[0] && l[0] == 0 && j < p.N && ep.rstd != nullptr) ep.rstd[(long)b[0] * p.N + j] = rs;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                xh[r][j] *= rs;
                acc[r][j] = xh[r][j];
                if (ep.norm == IINS_NORM_ADAIN && j < p.N) {
                    const float* ab = ep.adain + (long)b[r] * ep.adain_ld;
                    acc[r][j] = fmaf(xh[r][j], __ldg(ab + ep.adain_off_w + j), __ldg(ab + ep.adain_off_b + j));
                }
            }
        }
    } else if constexpr (EPI == 2) {
        // per-sample mean and UNBIASED std over (C*L), eps added to std (models.py:976-981)
        const float nel = (float)(L * p.N);
        float part = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < NACC; ++j) part += j < p.N ? acc[r][j] : 0.f;
        const float mean = iins_row2_sample_sum(part, nw, warp, lane, xch) / nel;
        float sq = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < NACC; ++j) { const float d = acc[r][j] - mean; sq += j < p.N ? d * d : 0.f; }
        sq = iins_row2_sample_sum(sq, nw, warp, lane, xch);
        const float rs = 1.0f / (sqrtf(sq / (nel - 1.f)) + IINS_EPS);
        if (ok[0] && l[0] == 0 && ep.rstd != nullptr) ep.rstd[b[0]] = rs;
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < NACC; ++j) {
                xh[r][j] = (acc[r][j] - mean) * rs;
                acc[r][j] = j < p.N ? fmaf(xh[r][j], __ldg(ep.gamma + j), __ldg(ep.beta + j)) : 0.f;
            }
    }
    if constexpr (EPI == 3) {
        // acc = gradient w.r.t. the previous layer's output (this kernel is that layer's consumer's data gradient); the
        // InstanceNorm backward of the previous layer follows in registers (same fusion as IINS_EPI_NBWD of the tensor-core
        // kernel):  dz = rstd * (raw - mean_l(raw) - xhat * mean_l(raw * xhat)),  raw = relu'(xhat) * dy.  Needs N == NACC.
        __shared__ float xchn[4][2 * NACC];
        const float invL = 1.0f / (float)L;
        float s1[NACC], s2[NACC];
        const bool relu = ep.nb_act == IINS_ACT_RELU;
#pragma unroll
        for (int r = 0; r < R; ++r) {                                // xh holds xhat of the previous layer here
            const long oi = (long)grow[r] * NACC;
#pragma unroll
            for (int j = 0; j < NACC; j += 4) {
                const float4 x4 = ok[r] ? __ldg(reinterpret_cast<const float4*>(ep.nb_xhat + oi + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
                xh[r][j] = x4.x; xh[r][j + 1] = x4.y; xh[r][j + 2] = x4.z; xh[r][j + 3] = x4.w;
            }
            if (ep.y != nullptr && ok[r]) {
#pragma unroll
                for (int j = 0; j < NACC; j += 4)
                    *reinterpret_cast<float4*>(ep.y + oi + j) = make_float4(acc[r][j], acc[r][j + 1], acc[r][j + 2], acc[r][j + 3]);
            }
        }
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (relu && !(xh[r][j] > 0.f)) acc[r][j] = 0.f;      // raw
                s1[j] = r == 0 ? acc[r][j] : s1[j] + acc[r][j];
                s2[j] = r == 0 ? acc[r][j] * xh[r][j] : fmaf(acc[r][j], xh[r][j], s2[j]);
            }
        }
        {
            const float m1 = iins_warp_sums<NACC>(s1, lane), m2 = iins_warp_sums<NACC>(s2, lane);
            if ((lane & (32 / NACC - 1)) == 0) {
                const int slot = iins_warp_sums_slot<NACC>(lane);
                xchn[warp][slot] = m1; xchn[warp][NACC + slot] = m2;
            }
            if (nw > 1) __syncthreads(); else __syncwarp();          // nw is CTA-uniform
            const int w0 = warp & ~(nw - 1);
#pragma unroll
            for (int j = 0; j < NACC; ++j) {
                float a = 0.f, c2 = 0.f;
                if (nw == 1) { a = xchn[warp][j]; c2 = xchn[warp][NACC + j]; }
                else for (int i = 0; i < nw; ++i) { a += xchn[w0 + i][j]; c2 += xchn[w0 + i][NACC + j]; }
                s1[j] = a; s2[j] = c2;
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (!ok[r]) continue;
            const long oi = (long)grow[r] * NACC;
#pragma unroll
            for (int j = 0; j < NACC; j += 4) {
                const float4 r4 = __ldg(reinterpret_cast<const float4*>(ep.nb_rstd + (long)b[r] * NACC + j));
                const float rv[4] = {r4.x, r4.y, r4.z, r4.w};
                float o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) o[i] = rv[i] * (acc[r][j + i] - s1[j + i] * invL - xh[r][j + i] * s2[j + i] * invL);
                *reinterpret_cast<float4*>(ep.nb_dz + oi + j) = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
        return;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (ep.act == IINS_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < NACC; ++j) acc[r][j] = fmaxf(acc[r][j], 0.f);
        } else if (ep.act != IINS_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < NACC; ++j) acc[r][j] = iins_act(acc[r][j], ep.act, ep.slope);
        }
        if (!ok[r]) continue;
        if (p.out_layout == IINS_NLC && (p.N % NACC) == 0) {
            const long oi = (long)grow[r] * p.N + n0;
            if ((EPI == 1 || EPI == 2) && ep.xhat != nullptr) {
#pragma unroll
                for (int j = 0; j < NACC; j += 4)
                    *reinterpret_cast<float4*>(ep.xhat + oi + j) = make_float4(xh[r][j], xh[r][j + 1], xh[r][j + 2], xh[r][j + 3]);
            }
            if (ep.add != nullptr) {
#pragma unroll
                for (int j = 0; j < NACC; j += 4) {
                    const float4 a4 = __ldg(reinterpret_cast<const float4*>(ep.add + oi + j));
                    acc[r][j] += a4.x; acc[r][j + 1] += a4.y; acc[r][j + 2] += a4.z; acc[r][j + 3] += a4.w;
                }
            }
#pragma unroll
            for (int j = 0; j < NACC; j += 4)
                *reinterpret_cast<float4*>(ep.y + oi + j) = make_float4(acc[r][j], acc[r][j + 1], acc[r][j + 2], acc[r][j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < NACC; ++j) {
                if (n0 + j >= p.N) continue;
                const long oi = p.out_layout == IINS_NCL ? (((long)b[r] * p.N + n0 + j) << p.lshift) + l[r] : (long)grow[r] * p.N + n0 + j;
                if ((EPI == 1 || EPI == 2) && ep.xhat != nullptr) ep.xhat[oi] = xh[r][j];
                float o = acc[r][j];
                if (ep.add != nullptr) o += ep.add[oi];
                ep.y[oi] = o;
            }
        }
    }
}

// Weight gradient for the same small layers: dW[n][k] (N <= 16, K <= 64) = sum_rows dz[row][n] * A[row][k].
// Persistent CTAs: each stages 64 rows of dz and of the im2col'd input in shared memory, thread (n, k-quad)
// accumulates in registers over the CTA's row range and flushes once with atomics.
struct IinsRowTNParams {
    IinsTNParams tn;
    int K;
    int lshift;
};

static __global__ void __launch_bounds__(256) iins_row_tn_kernel(const IinsRowTNParams rp) {
    iins_pdl_enter();
    constexpr int BR = 64, NT = 16, KMAX = 64;
    __shared__ float Zs[BR * (NT + 1)];        // [r][n]
    __shared__ float As[BR * (KMAX + 1)];      // [r][k]
    const IinsTNParams& p = rp.tn;
    const IinsGeom& g = p.g;
    const int tid = threadIdx.x;
    const int N = g.Cout, K = rp.K;
    const int n = tid & 15, kq = tid >> 4;     // thread owns dW[n][kq*4 .. kq*4+3]
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float bsum = 0.f;
    const int r_begin = blockIdx.x * p.rows_per_part;
    int r_end = r_begin + p.rows_per_part;
    if (r_end > p.M) r_end = p.M;
    for (int r0 = r_begin; r0 < r_end; r0 += BR) {
        // stage: 4 threads per row; each fills a quarter of the taps / columns of that row
        {
            const int r = tid >> 2, part = tid & 3;
            const int row = r0 + r;
            const bool ok = row < r_end;
            const int b = ok ? row >> rp.lshift : 0, l = ok ? row & (g.Lout - 1) : 0;
            for (int nn = part; nn < NT; nn += 4) Zs[r * (NT + 1) + nn] = (ok && nn < N) ? iins_dz_at(g, p.dz, b, l, nn) : 0.f;
            for (int t = part; t < g.ks; t += 4) {
                const int pos = ok ? iins_src_pos(g, l, t) : -1;
                const float* xr = p.x + (pos >= 0 ? iins_in_index(g, b, pos, 0) : 0);
                const long cstride = g.in_layout == IINS_NCL ? g.Lin : 1;
                if (g.in_layout == IINS_NLC && (g.Cin & 3) == 0) {
                    for (int c = 0; c < g.Cin; c += 4) {
                        float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (pos >= 0) a4 = __ldg(reinterpret_cast<const float4*>(xr + c));
                        float* dst = As + r * (KMAX + 1) + t * g.Cin + c;
                        dst[0] = a4.x; dst[1] = a4.y; dst[2] = a4.z; dst[3] = a4.w;
                    }
                } else {
                    for (int c = 0; c < g.Cin; ++c) As[r * (KMAX + 1) + t * g.Cin + c] = pos >= 0 ? __ldg(xr + c * cstride) : 0.f;
                }
            }
        }
        __syncthreads();
        if (n < N && kq * 4 < K) {
#pragma unroll 4
            for (int r = 0; r < BR; ++r) {
                const float z = Zs[r * (NT + 1) + n];
                const float* ar = As + r * (KMAX + 1) + kq * 4;
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[j] = fmaf(z, ar[j], acc[j]);
            }
        }
        if (p.db != nullptr && tid < N) {
            for (int r = 0; r < BR; ++r) bsum += Zs[r * (NT + 1) + tid];
        }
        __syncthreads();
    }
    if (n < N) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = kq * 4 + j;
            if (k < K) { int t = k / g.Cin, c = k - t * g.Cin; atomicAdd(p.dw + iins_w_index(g, n, c, t), acc[j]); }
        }
    }
    if (p.db != nullptr && tid < N) atomicAdd(p.db + tid, bsum);
}

// ---- weight gradient of the small-channel conv layers, one thread per row ("row2") ---------------------------
// dW[n][k] = sum_rows dz[row][n] * A[row][k] with M = B*L huge (up to 524288) and N*K tiny (28..512): a tall-skinny
// reduction.  Lane = row: a thread loads dz[row][0..NACC) and the im2col slice A[row][k-slice] (TPS taps x CIN
// channels, compile-time so everything stays in registers), does the NACC x KS outer-product FMAs into its own
// accumulators and walks on to its next row; only at the very end the 128 threads of the CTA are summed through
// shared memory and the CTA issues one atomic per output.  grid = (row parts, k slices).
struct IinsRow2TNParams {
    IinsTNParams tn;
    int lshift;                // log2(Lout)
};

template <int NACC, int CIN, int TPS>
static __global__ void __launch_bounds__(128) iins_row2_tn_kernel(const IinsRow2TNParams rp) {
    iins_pdl_enter();
    constexpr int KS = CIN * TPS;              // k entries per thread
    constexpr int NA = NACC * KS;              // accumulators per thread
    __shared__ float red[128][33];
    const IinsTNParams& p = rp.tn;
    const IinsGeom& g = p.g;
    const int tid = threadIdx.x;
    const int t0 = blockIdx.y * TPS;           // first tap of this CTA's k slice
    const bool do_bias = p.db != nullptr && blockIdx.y == 0;
    float acc[NA], bacc[NACC];
#pragma unroll
    for (int i = 0; i < NA; ++i) acc[i] = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) bacc[i] = 0.f;
    const int r_begin = blockIdx.x * p.rows_per_part;
    int r_end = r_begin + p.rows_per_part;
    if (r_end > p.M) r_end = p.M;
    const bool zvec = g.out_layout == IINS_NLC && g.Cout == NACC && (NACC & 3) == 0 && !p.dz.dy_bcast;
    const bool masked = p.dz.y != nullptr && p.dz.act != IINS_ACT_NONE;
    for (int row = r_begin + tid; row < r_end; row += 128) {
        const int b = row >> rp.lshift, l = row & (g.Lout - 1);
        float z[NACC], a[KS];
        if (zvec) {
            const long zi = (long)row * NACC;
#pragma unroll
            for (int j = 0; j < NACC; j += 4) {
                float4 d4 = __ldg(reinterpret_cast<const float4*>(p.dz.dy + zi + j));
                if (masked) {
                    const float4 y4 = __ldg(reinterpret_cast<const float4*>(p.dz.y + zi + j));
                    d4.x *= iins_dact_from_y(y4.x, p.dz.act, p.dz.slope); d4.y *= iins_dact_from_y(y4.y, p.dz.act, p.dz.slope);
                    d4.z *= iins_dact_from_y(y4.z, p.dz.act, p.dz.slope); d4.w *= iins_dact_from_y(y4.w, p.dz.act, p.dz.slope);
                }
                z[j] = d4.x * p.dz.dy_scale; z[j + 1] = d4.y * p.dz.dy_scale; z[j + 2] = d4.z * p.dz.dy_scale; z[j + 3] = d4.w * p.dz.dy_scale;
            }
        } else {
#pragma unroll
            for (int j = 0; j < NACC; ++j) z[j] = j < g.Cout ? iins_dz_at(g, p.dz, b, l, j) : 0.f;
        }
#pragma unroll
        for (int tt = 0; tt < TPS; ++tt) {
            const int t = t0 + tt;
            const int pos = t < g.ks ? iins_src_pos(g, l, t) : -1;
            if (CIN >= 4 && g.in_layout == IINS_NLC) {
                const float* xr = p.x + ((long)b * g.Lin + (pos >= 0 ? pos : 0)) * CIN;
#pragma unroll
                for (int c = 0; c < CIN; c += 4) {
                    float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (pos >= 0) a4 = __ldg(reinterpret_cast<const float4*>(xr + c));
                    a[tt * CIN + c] = a4.x; a[tt * CIN + c + 1] = a4.y; a[tt * CIN + c + 2] = a4.z; a[tt * CIN + c + 3] = a4.w;
                }
            } else {
#pragma unroll
                for (int c = 0; c < CIN; ++c) a[tt * CIN + c] = pos >= 0 ? __ldg(p.x + iins_in_index(g, b, pos, c)) : 0.f;
            }
        }
#pragma unroll
        for (int n = 0; n < NACC; ++n) {
#pragma unroll
            for (int k = 0; k < KS; ++k) acc[n * KS + k] = fmaf(z[n], a[k], acc[n * KS + k]);
            bacc[n] += z[n];
        }
    }
    // ---- CTA reduction, 32 accumulators at a time: red[thread][value] -> 4 partial sums per value -> one atomic
    const int col = tid & 31, part = tid >> 5;
#pragma unroll
    for (int c0 = 0; c0 < NA + NACC; c0 += 32) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int e = c0 + i;                           // compile-time after unrolling
            red[tid][i] = e < NA ? acc[e < NA ? e : 0] : (e < NA + NACC ? bacc[e < NA + NACC && e >= NA ? e - NA : 0] : 0.f);
        }
        __syncthreads();
        float sum = 0.f;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) sum += red[part * 32 + r][col];
        __syncthreads();
        red[part][col] = sum;
        __syncthreads();
        if (part == 0) {
            sum = red[0][col] + red[1][col] + red[2][col] + red[3][col];
            const int e = c0 + col;
            if (e < NA) {
                const int n = e / KS, k = e - n * KS;
                const int tt = k / CIN, c = k - tt * CIN, t = t0 + tt;
                if (n < g.Cout && t < g.ks) atomicAdd(p.dw + iins_w_index(g, n, c, t), sum);
            } else if (e < NA + NACC && do_bias) {
                if (e - NA < g.Cout) atomicAdd(p.db + (e - NA), sum);
            }
        }
    }
}

// Weight gradient of "thin" layers: one of the two GEMM dims is <= 4 (the 1x1 conv on the 2-channel range code:
// K = 2; the Restorer's last Linear: N = 1).  dW is then T <= 4 rank-1 accumulations of a W-wide row vector:
// thread = one wide column, 256 / W row groups per CTA, coalesced reads of the wide operand, T accumulators per
// thread, one atomic flush.  These layers would waste a 128-lane tensor-core tile and their operands are not
// 8-channel gatherable (NCL range code, single output channel).
struct IinsThinTNParams {
    IinsTNParams tn;
    int K;                     // ks * Cin
    int lshift;                // log2(Lout)
    int thin_is_k;             // 1: K <= 4 (wide = output channels), 0: Cout <= 4 (wide = k)
};

static __global__ void __launch_bounds__(256) iins_thin_tn_kernel(const IinsThinTNParams tp) {
    iins_pdl_enter();
    const IinsTNParams& p = tp.tn;
    const IinsGeom& g = p.g;
    const int N = g.Cout, K = tp.K;
    const int W = tp.thin_is_k ? N : K, T = tp.thin_is_k ? K : N;
    const int col = threadIdx.x % W, grp = threadIdx.x / W, ngrp = 256 / W;       // host guarantees 256 % W == 0
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, bthin[4] = {0.f, 0.f, 0.f, 0.f};
    float bwide = 0.f;
    const int r_begin = blockIdx.x * p.rows_per_part;
    int r_end = r_begin + p.rows_per_part;
    if (r_end > p.M) r_end = p.M;
    const int wt = tp.thin_is_k ? 0 : col / g.Cin, wc = tp.thin_is_k ? 0 : col - wt * g.Cin;   // wide k -> (tap, channel)
#pragma unroll 4
    for (int row = r_begin + grp; row < r_end; row += ngrp) {
        const int b = row >> tp.lshift, l = row & (g.Lout - 1);
        float thin[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            float v = 0.f;
            if (t < T) {
                if (tp.thin_is_k) { const int tt = t / g.Cin, ci = t - tt * g.Cin; v = iins_a_fwd(g, p.x, b, l, tt, ci); }
                else v = iins_dz_at(g, p.dz, b, l, t);
            }
            thin[t] = v;
        }
        const float wv = tp.thin_is_k ? iins_dz_at(g, p.dz, b, l, col) : iins_a_fwd(g, p.x, b, l, wt, wc);
#pragma unroll
        for (int t = 0; t < 4; ++t) { acc[t] = fmaf(thin[t], wv, acc[t]); bthin[t] += thin[t]; }
        bwide += wv;
    }
    // combine the row groups of the CTA in shared memory, then one atomic per output element and CTA
    __shared__ float red[9][256];
#pragma unroll
    for (int t = 0; t < 4; ++t) { red[t][threadIdx.x] = acc[t]; red[4 + t][threadIdx.x] = bthin[t]; }
    red[8][threadIdx.x] = bwide;
    __syncthreads();
    if (grp != 0) return;
    for (int gi = 1; gi < ngrp; ++gi) {
#pragma unroll
        for (int t = 0; t < 4; ++t) { acc[t] += red[t][gi * W + col]; bthin[t] += red[4 + t][gi * W + col]; }
        bwide += red[8][gi * W + col];
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        if (t >= T) continue;
        const int n = tp.thin_is_k ? col : t, k = tp.thin_is_k ? t : col;
        const int tt = k / g.Cin, ci = k - tt * g.Cin;
        atomicAdd(p.dw + iins_w_index(g, n, ci, tt), acc[t]);
        if (p.db != nullptr && !tp.thin_is_k && col == 0) atomicAdd(p.db + t, bthin[t]);
    }
    if (p.db != nullptr && tp.thin_is_k) atomicAdd(p.db + col, bwide);
}
