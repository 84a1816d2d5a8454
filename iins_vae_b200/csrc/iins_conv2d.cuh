// Kernels of the 2-D variant (conv_type = 2, expand = True; models.py:179-215 RangeEncoder2d, :304-346 EnvEncoder2d,
// :474-539 Decoder2d, :1008-1025 ResidualBlock2d, :1082-1113 AdaptiveInstanceNorm2d).
//
// Activations are channels-last (B, H, W, C).  A Conv2d runs as   VERTICAL im2col  +  the 1-D implicit-GEMM kernels:
//     xs[b, ho, w, kh * C + c] = x_padH[b, ho * stride + kh, w, c]          (this file: the H direction incl. its padding)
//     y[(b, ho), wo, co]       = conv1d over w of xs with weights W'[co][kh * C + c][kw] = W[co][c][kh][kw]
// i.e. the rows (b, ho) are the 1-D path's "samples", W is its length, k * C (zero-padded to a power of two so that the
// tensor-core gathers apply) its channel count; the W direction's padding mode / stride / upsampling are the 1-D kernels'
// own (reflection, zero, nearest x2).  The data gradient comes back the same way (1-D dgrad -> dxs, then a gather over
// the taps that read a row: iins_v2c_bwd_kernel), the weight gradient is the 1-D wgrad against xs followed by the inverse
// weight permutation.  The norms reduce over H * W positions, which no GEMM tile holds: they are their own kernels here
// (one CTA per sample).
#pragma once
#include "iins_common.cuh"

struct IinsV2cParams {
    const float* x;        // (B, Hi, Wi, C); in_w_bcast: (B, Hi, C) broadcast along w (the expanded CIR, models.py:55)
    float* xs;             // (B, Ho, Wi, Cp)
    int B, Hi, Wi, C, Ho, k, stride, pad, mode, Cp, in_w_bcast;
};

// source row of output row ho, tap kh; -1 = zero padding
IINS_HD int iins_v2c_src_row(int ho, int kh, int Hi, int stride, int pad, int mode) {
    int u = ho * stride + kh - pad;
    if (mode == IINS_PAD_REFLECT) {
        if (u < 0) u = -u;
        else if (u >= Hi) u = 2 * (Hi - 1) - u;
        return u;
    }
    if (mode == IINS_PAD_UP2) return (u < 0 || u >= 2 * Hi) ? -1 : (u >> 1);
    return (u < 0 || u >= Hi) ? -1 : u;
}

static __global__ void __launch_bounds__(256) iins_v2c_fwd_kernel(const IinsV2cParams p) {
    iins_pdl_enter();
    const int KC = p.k * p.C;
    if ((p.C & 3) == 0 && !p.in_w_bcast) {
        // 16 bytes per thread: four consecutive channels of one tap (C % 4 == 0, so a float4 never straddles two taps)
        const int Cq = p.Cp >> 2;
        const long n = (long)p.B * p.Ho * p.Wi * Cq;
        for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
            const int q = (int)(e % Cq) * 4;
            long r = e / Cq;
            const int w = (int)(r % p.Wi);
            r /= p.Wi;
            const int ho = (int)(r % p.Ho), b = (int)(r / p.Ho);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q < KC) {
                const int kh = q / p.C, c = q - kh * p.C;
                const int h = iins_v2c_src_row(ho, kh, p.Hi, p.stride, p.pad, p.mode);
                if (h >= 0) v = __ldg(reinterpret_cast<const float4*>(p.x + (((long)b * p.Hi + h) * p.Wi + w) * p.C + c));
            }
            reinterpret_cast<float4*>(p.xs)[e] = v;
        }
        return;
    }
    const long n = (long)p.B * p.Ho * p.Wi * p.Cp;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        const int q = (int)(e % p.Cp);
        long r = e / p.Cp;
        const int w = (int)(r % p.Wi);
        r /= p.Wi;
        const int ho = (int)(r % p.Ho), b = (int)(r / p.Ho);
        float v = 0.f;
        if (q < KC) {
            const int kh = q / p.C, c = q - kh * p.C;
            const int h = iins_v2c_src_row(ho, kh, p.Hi, p.stride, p.pad, p.mode);
            if (h >= 0) v = p.in_w_bcast ? __ldg(p.x + ((long)b * p.Hi + h) * p.C + c) : __ldg(p.x + (((long)b * p.Hi + h) * p.Wi + w) * p.C + c);
        }
        p.xs[e] = v;
    }
}

// dx[b, h, w, c] = sum over (ho, kh) whose source row is h of dxs[b, ho, w, kh * C + c]   (+ add: a skip gradient)
struct IinsV2cBwdParams {
    const float* dxs;      // (B, Ho, Wi, Cp)
    const float* add;      // (B, Hi, Wi, C) or nullptr
    float* dx;             // (B, Hi, Wi, C)
    int B, Hi, Wi, C, Ho, k, stride, pad, mode, Cp;
};

template <int V>          // V = 4: float4 over channels (C % 4 == 0), V = 1: scalar
__device__ __forceinline__ void iins_v2c_bwd_body(const IinsV2cBwdParams& p) {
    const int Cv = p.C / V;
    const long n = (long)p.B * p.Hi * p.Wi * Cv;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        const int c = (int)(e % Cv) * V;
        long r = e / Cv;
        const int w = (int)(r % p.Wi);
        r /= p.Wi;
        const int h = (int)(r % p.Hi), b = (int)(r / p.Hi);
        // padded / upsampled coordinates u (before the tap offset) that read source row h
        int cand[3];
        int nc = 0;
        if (p.mode == IINS_PAD_REFLECT) {
            cand[nc++] = h;
            if (h >= 1 && h <= p.pad) cand[nc++] = -h;
            if (h <= p.Hi - 2 && h >= p.Hi - 1 - p.pad) cand[nc++] = 2 * (p.Hi - 1) - h;
        } else if (p.mode == IINS_PAD_UP2) {
            cand[nc++] = 2 * h;
            cand[nc++] = 2 * h + 1;
        } else {
            cand[nc++] = h;
        }
        float acc[V];
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = p.add != nullptr ? __ldg(p.add + e * V + j) : 0.f;
        for (int i = 0; i < nc; ++i) {
            for (int kh = 0; kh < p.k; ++kh) {
                const int t = cand[i] + p.pad - kh;            // = ho * stride
                if (t < 0) continue;
                const int ho = t / p.stride;
                if (ho * p.stride != t || ho >= p.Ho) continue;
                const float* src = p.dxs + (((long)b * p.Ho + ho) * p.Wi + w) * p.Cp + kh * p.C + c;
                if (V == 4) { const float4 v = __ldg(reinterpret_cast<const float4*>(src)); acc[0] += v.x; acc[1 % V] += v.y; acc[2 % V] += v.z; acc[3 % V] += v.w; }
                else acc[0] += __ldg(src);
            }
        }
        if (V == 4) reinterpret_cast<float4*>(p.dx)[e] = make_float4(acc[0], acc[1 % V], acc[2 % V], acc[3 % V]);
        else p.dx[e] = acc[0];
    }
}
static __global__ void __launch_bounds__(256) iins_v2c_bwd_kernel(const IinsV2cBwdParams p) {
    iins_pdl_enter();
    if ((p.C & 3) == 0) iins_v2c_bwd_body<4>(p);
    else iins_v2c_bwd_body<1>(p);
}

// W[co][ci][kh][kw] -> W'[co][kh * C + ci (zero-padded to Cp)][kw]  and the inverse accumulation for the gradient
static __global__ void __launch_bounds__(256) iins_wperm_fwd_kernel(const float* __restrict__ w, float* __restrict__ wp, int Cout, int C, int k, int Cp) {
    iins_pdl_enter();
    const long n = (long)Cout * Cp * k;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        const int kw = (int)(e % k);
        const int q = (int)((e / k) % Cp), co = (int)(e / ((long)k * Cp));
        float v = 0.f;
        if (q < k * C) { const int kh = q / C, ci = q - kh * C; v = __ldg(w + (((long)co * C + ci) * k + kh) * k + kw); }
        wp[e] = v;
    }
}
static __global__ void __launch_bounds__(256) iins_wperm_bwd_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int Cout, int C, int k, int Cp) {
    iins_pdl_enter();
    const long n = (long)Cout * C * k * k;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        const int kw = (int)(e % k), kh = (int)((e / k) % k);
        const int ci = (int)((e / ((long)k * k)) % C), co = (int)(e / ((long)k * k * C));
        dw[e] += __ldg(dwp + ((long)co * Cp + kh * C + ci) * k + kw);
    }
}

// ---- norms over H * W positions ------------------------------------------------------------------------------------------
// One CTA (256 threads) per sample; (L = H*W, C) channels-last, C a power of two in [4, 256], L * C a multiple of 1024.
// A thread's float4 covers 4 fixed channels (256 % (C/4) == 0), so the per-channel sums are per-thread partials combined
// by a shared-memory tree whose strides stay multiples of the column-group count; LayerNorm continues the tree to 1.
struct IinsNorm2dParams {
    int B, L, C, norm, act;
    const float* z;            // forward: pre-norm values (conv output incl. bias);  backward: dy (grad w.r.t. the layer output)
    const float* add;          // forward: residual operand added after norm (+ act) or nullptr
    float* y;                  // forward: output;  backward: dz
    float* xhat;               // saved normalised values (forward: written, backward: read)
    float* rstd;               // IN / AdaIN: [B * C];  LN: [B] = 1 / (std + eps)
    const float* gamma; const float* beta;       // LN
    float* dgamma; float* dbeta;                 // LN backward (atomically accumulated)
    const float* adain; float* dadain;           // AdaIN parameters (B, ld): bias at +off_b, weight at +off_w; their gradients
    int adain_ld, adain_off_b, adain_off_w;
};

// sums v[0..3] over the threads that share this thread's column group (IN / AdaIN) or over the whole CTA (all = true)
__device__ __forceinline__ void iins_n2d_reduce(float* v, float (*sm)[256], int CG, bool all) {
    const int t = threadIdx.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) sm[j][t] = v[j];
    __syncthreads();
    const int stop = all ? 1 : CG;
    for (int s = 128; s >= stop; s >>= 1) {
        if (t < s) {
#pragma unroll
            for (int j = 0; j < 4; ++j) sm[j][t] += sm[j][t + s];
        }
        __syncthreads();
    }
    if (all) {
        const float tot = (sm[0][0] + sm[1][0]) + (sm[2][0] + sm[3][0]);
        v[0] = v[1] = v[2] = v[3] = tot;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = sm[j][t & (CG - 1)];
    }
    __syncthreads();
}

static __global__ void __launch_bounds__(256) iins_norm2d_fwd_kernel(const IinsNorm2dParams p) {
    iins_pdl_enter();
    __shared__ float sm[4][256];
    const int t = threadIdx.x, CG = p.C >> 2, c0 = (t & (CG - 1)) * 4;
    const int nel = p.L * p.C, nstep = nel >> 10;
    const bool ln = p.norm == IINS_NORM_LN;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        const float4* z4 = reinterpret_cast<const float4*>(p.z + (long)b * nel);
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < nstep; ++i) { const float4 v = z4[t + 256 * i]; s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w; }
        iins_n2d_reduce(s, sm, CG, ln);
        const float cnt = ln ? (float)nel : (float)p.L;
        const float mean[4] = {s[0] / cnt, s[1] / cnt, s[2] / cnt, s[3] / cnt};
        float q[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < nstep; ++i) {
            const float4 v = z4[t + 256 * i];
            const float d0 = v.x - mean[0], d1 = v.y - mean[1], d2 = v.z - mean[2], d3 = v.w - mean[3];
            q[0] = fmaf(d0, d0, q[0]); q[1] = fmaf(d1, d1, q[1]); q[2] = fmaf(d2, d2, q[2]); q[3] = fmaf(d3, d3, q[3]);
        }
        iins_n2d_reduce(q, sm, CG, ln);
        float rs[4], sc[4], sh[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // LN: unbiased std, eps added to std (models.py:976-981);  IN / AdaIN: biased variance, eps inside the root (:152, :1072)
            rs[j] = ln ? 1.0f / (sqrtf(q[j] / (float)(nel - 1)) + IINS_EPS) : 1.0f / sqrtf(q[j] / (float)p.L + IINS_EPS);
            sc[j] = 1.f; sh[j] = 0.f;
            if (ln) { sc[j] = __ldg(p.gamma + c0 + j); sh[j] = __ldg(p.beta + c0 + j); }
            else if (p.norm == IINS_NORM_ADAIN) {
                sc[j] = __ldg(p.adain + (long)b * p.adain_ld + p.adain_off_w + c0 + j);
                sh[j] = __ldg(p.adain + (long)b * p.adain_ld + p.adain_off_b + c0 + j);
            }
        }
        if (ln) { if (t == 0) p.rstd[b] = rs[0]; }
        else if (t < CG) *reinterpret_cast<float4*>(p.rstd + (long)b * p.C + c0) = make_float4(rs[0], rs[1], rs[2], rs[3]);
        float4* xh4 = reinterpret_cast<float4*>(p.xhat + (long)b * nel);
        float4* y4 = reinterpret_cast<float4*>(p.y + (long)b * nel);
        const float4* a4 = p.add != nullptr ? reinterpret_cast<const float4*>(p.add + (long)b * nel) : nullptr;
        for (int i = 0; i < nstep; ++i) {
            const float4 v = z4[t + 256 * i];
            const float4 xh = make_float4((v.x - mean[0]) * rs[0], (v.y - mean[1]) * rs[1], (v.z - mean[2]) * rs[2], (v.w - mean[3]) * rs[3]);
            float4 o = make_float4(fmaf(xh.x, sc[0], sh[0]), fmaf(xh.y, sc[1], sh[1]), fmaf(xh.z, sc[2], sh[2]), fmaf(xh.w, sc[3], sh[3]));
            if (p.act == IINS_ACT_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            if (a4 != nullptr) { const float4 a = a4[t + 256 * i]; o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
            xh4[t + 256 * i] = xh;
            y4[t + 256 * i] = o;
        }
    }
}

// backward (same formulas as iins_norm_bwd_kernel, iins_misc.cuh):
//   raw = dy masked by the layer's ReLU (u = xhat * scale + shift > 0), g = raw * scale
//   IN / AdaIN: dz = rstd * scale * (raw - mean_l(raw) - xhat * mean_l(raw * xhat));  d bias = sum_l raw, d weight = sum_l raw * xhat
//   LN:         dz = r * (g - mean(g)) - xhat * sum(g * xhat) / ((n - 1) * std),  r = 1 / (std + eps);  dgamma / dbeta per channel
static __global__ void __launch_bounds__(256) iins_norm2d_bwd_kernel(const IinsNorm2dParams p) {
    iins_pdl_enter();
    __shared__ float sm[4][256];
    const int t = threadIdx.x, CG = p.C >> 2, c0 = (t & (CG - 1)) * 4;
    const int nel = p.L * p.C, nstep = nel >> 10;
    const bool ln = p.norm == IINS_NORM_LN, relu = p.act == IINS_ACT_RELU;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        float sc[4], sh[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            sc[j] = 1.f; sh[j] = 0.f;
            if (ln) { sc[j] = __ldg(p.gamma + c0 + j); sh[j] = __ldg(p.beta + c0 + j); }
            else if (p.norm == IINS_NORM_ADAIN) {
                sc[j] = __ldg(p.adain + (long)b * p.adain_ld + p.adain_off_w + c0 + j);
                sh[j] = __ldg(p.adain + (long)b * p.adain_ld + p.adain_off_b + c0 + j);
            }
        }
        const float4* dy4 = reinterpret_cast<const float4*>(p.z + (long)b * nel);
        const float4* xh4 = reinterpret_cast<const float4*>(p.xhat + (long)b * nel);
        float sr[4] = {0.f, 0.f, 0.f, 0.f}, srx[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < nstep; ++i) {
            const float4 d = dy4[t + 256 * i], x = xh4[t + 256 * i];
            const float dv[4] = {d.x, d.y, d.z, d.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float u = fmaf(xv[j], sc[j], sh[j]);
                const float raw = (relu && !(u > 0.f)) ? 0.f : dv[j];
                sr[j] += raw; srx[j] = fmaf(raw, xv[j], srx[j]);
            }
        }
        float tg[4], tgx[4];                       // LN: whole-sample sums of g and g * xhat (g = raw * gamma)
#pragma unroll
        for (int j = 0; j < 4; ++j) { tg[j] = sr[j] * sc[j]; tgx[j] = srx[j] * sc[j]; }
        iins_n2d_reduce(sr, sm, CG, false);
        iins_n2d_reduce(srx, sm, CG, false);
        if (ln) {
            iins_n2d_reduce(tg, sm, CG, true);
            iins_n2d_reduce(tgx, sm, CG, true);
            if (t < CG) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { atomicAdd(p.dgamma + c0 + j, srx[j]); atomicAdd(p.dbeta + c0 + j, sr[j]); }
            }
        } else if (p.norm == IINS_NORM_ADAIN && p.dadain != nullptr && t < CG) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                p.dadain[(long)b * p.adain_ld + p.adain_off_b + c0 + j] = sr[j];
                p.dadain[(long)b * p.adain_ld + p.adain_off_w + c0 + j] = srx[j];
            }
        }
        float rs[4], coef = 0.f, mean_g = 0.f;
        if (ln) {
            const float r = __ldg(p.rstd + b), sd = 1.0f / r - IINS_EPS;
            rs[0] = rs[1] = rs[2] = rs[3] = r;
            coef = tgx[0] / ((float)(nel - 1) * sd);
            mean_g = tg[0] / (float)nel;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) rs[j] = __ldg(p.rstd + (long)b * p.C + c0 + j);
        }
        const float invL = 1.0f / (float)p.L;
        float4* dz4 = reinterpret_cast<float4*>(p.y + (long)b * nel);
        for (int i = 0; i < nstep; ++i) {
            const float4 d = dy4[t + 256 * i], x = xh4[t + 256 * i];
            const float dv[4] = {d.x, d.y, d.z, d.w}, xv[4] = {x.x, x.y, x.z, x.w};
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float u = fmaf(xv[j], sc[j], sh[j]);
                const float raw = (relu && !(u > 0.f)) ? 0.f : dv[j];
                if (ln) o[j] = rs[j] * (raw * sc[j] - mean_g) - xv[j] * coef;
                else o[j] = rs[j] * sc[j] * (raw - sr[j] * invL - xv[j] * srx[j] * invL);
            }
            dz4[t + 256 * i] = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

// ---- small helpers -----------------------------------------------------------------------------------------------------------
// dst[b, i, c] = src[b, c] * scale  for i < n      (backward of the global average pool: models.py:333)
static __global__ void __launch_bounds__(256) iins_bcast_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int n, int C, float scale) {
    iins_pdl_enter();
    const long tot = (long)B * n * C;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (long)gridDim.x * blockDim.x) {
        const int c = (int)(e % C), b = (int)(e / ((long)n * C));
        dst[e] = __ldg(src + (long)b * C + c) * scale;
    }
}
// column 0 of a (B, H, W) map <-> a (B, H) vector (the decoder output keeps x_recon[:, :, :, 0] only, models.py:90)
static __global__ void __launch_bounds__(256) iins_col0_get_kernel(const float* __restrict__ y, float* __restrict__ col, int B, int H, int W) {
    iins_pdl_enter();
    const long tot = (long)B * H;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (long)gridDim.x * blockDim.x) col[e] = __ldg(y + e * W);
}
static __global__ void __launch_bounds__(256) iins_col0_put_kernel(const float* __restrict__ col, float* __restrict__ dy, int B, int H, int W) {
    iins_pdl_enter();
    const long tot = (long)B * H * W;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (long)gridDim.x * blockDim.x) dy[e] = (e % W) == 0 ? __ldg(col + e / W) : 0.f;
}
// elementwise tanh' : dy *= (1 - y^2)   (the output convolution's Tanh, models.py:505)
static __global__ void __launch_bounds__(256) iins_tanh_bwd_kernel(float* __restrict__ dy, const float* __restrict__ y, long n) {
    iins_pdl_enter();
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) { const float t = __ldg(y + e); dy[e] *= (1.f - t * t); }
}

// (n, C) -> (n, Cp) with zero channels appended: the weight-gradient GEMMs gather 8 channels of dz at a time, so a convolution
// with fewer than 8 output channels (the last up-sampling stage, the output convolution) runs them on a padded copy
static __global__ void __launch_bounds__(256) iins_pad_channels_kernel(const float* __restrict__ src, float* __restrict__ dst, long n, int C, int Cp) {
    iins_pdl_enter();
    const long tot = n * Cp;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (long)gridDim.x * blockDim.x) {
        const int c = (int)(e % Cp);
        dst[e] = c < C ? __ldg(src + (e / Cp) * C + c) : 0.f;
    }
}
static __global__ void __launch_bounds__(64) iins_add_n_kernel(float* __restrict__ dst, const float* __restrict__ src, int n) {
    iins_pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += __ldg(src + i);
}
