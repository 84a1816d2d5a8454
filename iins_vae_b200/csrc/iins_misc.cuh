// Non-GEMM kernels of the path: norm backward, adaptive pooling, KL / reparameterisation,
// fused loss (+ seed gradients + metrics), fused Adam.
#pragma once
#include "iins_common.cuh"

// ------------------------------------------------------------------------ norm backward
// Backward of  y = act(affine(xhat)) (+ residual handled by the caller)  for one tile of whole samples.
//   IN    : xhat = (z-mean)*rstd, no affine        dz = rstd*(g - mean_l(g) - xhat*mean_l(g*xhat))
//   AdaIN : same with per-(b,c) weight/bias        + d_adain
//   LN    : xhat = (z-mean)/(std+eps), per-channel gamma/beta (models.py:976-985)
//           dz = r*(g - mean(g)) - xhat * sum(g*xhat) / ((n-1)*std),  r = 1/(std+eps)
struct IinsNormBwdParams {
    int B, L, C;
    int norm, act;
    const float* dy;           // grad w.r.t. the layer output (NLC)
    const float* xhat;         // saved normalised values (NLC)
    const float* rstd;         // IN/AdaIN [B*C]; LN [B]
    const float* gamma;        // LN
    const float* beta;
    float* dgamma;             // LN, atomically accumulated
    float* dbeta;
    const float* adain;        // AdaIN params (B, ld)
    float* dadain;             // AdaIN grads  (B, ld), written (each entry owned by one (b,c))
    int adain_ld, adain_off_b, adain_off_w;
    float* dz;                 // out (NLC)
    // column blocks: the kernel handles C <= 128 channels per launch; a wider InstanceNorm / AdaIN layer (per-channel
    // statistics are independent) is processed as blocks of 128 channels: the host offsets dy / xhat / dz / rstd and the
    // AdaIN offsets by the block's first channel and passes the full row stride here
    int ldc;                   // floats between consecutive rows (positions) of dy / xhat / dz: the layer's full channel count
    int rstd_ld;               // floats between consecutive samples of rstd (IN / AdaIN)
};

// One WARP per sample (the per-sample tensors are 2 KB: L*C = 512 on this path): two passes over the sample's
// (L, C) block, 16 bytes per lane per access, reductions by warp shuffles -- no shared memory, no block barrier.
//   pass 1: accumulate sum_l g and sum_l g*xhat per channel (IN / AdaIN) or over the whole sample (LN)
//   pass 2: reload (L1-resident) and write dz
// Requires C a power of two, 4 <= C <= 128, and L*C a multiple of 128.
static __global__ void __launch_bounds__(256) iins_norm_bwd_kernel(const IinsNormBwdParams p) {
    iins_pdl_enter();
    // LayerNorm's dgamma / dbeta are sums over the whole batch: every warp keeps a running per-channel total over the
    // samples it visits (persistent grid), the CTA combines its 8 warps in shared memory and issues ONE atomic per
    // channel -- one atomic per (sample, channel) serialised 4096 deep on the same few L2 sectors.
    __shared__ float s_dg[8][128], s_db[8][128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = p.C, L = p.L;
    // C >= 4: a lane's float4 covers 4 channels of one row (CG column groups per row); C < 4 (dim < 4 configurations): a
    // float4 covers 4/C rows and element j has channel j & (C-1)
    const int CG = C >= 4 ? C >> 2 : 1;             // float4 column groups per row (<= 32)
    const int nstep = (L * C) >> 7;                 // float4 steps of 32 lanes
    const bool relu = p.act == IINS_ACT_RELU;
    const int cg = lane & (CG - 1), c0 = cg * 4;    // CG <= 32 and nstep*32 is a multiple of CG: the lane's channels are fixed
    // float4 index of this lane's element in step i: contiguous when the rows are dense (ldc == C); otherwise row * ldc/4 + cg
    const int cgs = 31 - __clz(CG), ld4 = p.ldc >> 2;
    const bool dense = p.ldc == C;
    auto idx4 = [&](int i) { const int f = lane + 32 * i; return dense ? f : (f >> cgs) * ld4 + (f & (CG - 1)); };
    int ch[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) ch[j] = C >= 4 ? c0 + j : (j & (C - 1));
    const int nch = C >= 4 ? 4 : C;                 // distinct channels held by a lane
    float scale[4] = {1.f, 1.f, 1.f, 1.f}, shift[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.norm == IINS_NORM_LN) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { scale[j] = __ldg(p.gamma + ch[j]); shift[j] = __ldg(p.beta + ch[j]); }
    }
    float tot_dg[4] = {0.f, 0.f, 0.f, 0.f}, tot_db[4] = {0.f, 0.f, 0.f, 0.f};
    const float invL = 1.0f / (float)L;
    for (int b = blockIdx.x * 8 + warp; b < p.B; b += gridDim.x * 8) {          // warp-uniform
        if (p.norm == IINS_NORM_ADAIN) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                scale[j] = __ldg(p.adain + (long)b * p.adain_ld + p.adain_off_w + ch[j]);
                shift[j] = __ldg(p.adain + (long)b * p.adain_ld + p.adain_off_b + ch[j]);
            }
        }
        const float4* dy4 = reinterpret_cast<const float4*>(p.dy + (long)b * L * p.ldc);
        const float4* xh4 = reinterpret_cast<const float4*>(p.xhat + (long)b * L * p.ldc);
        float sg[4] = {0.f, 0.f, 0.f, 0.f}, sgx[4] = {0.f, 0.f, 0.f, 0.f};        // grad wrt xhat
        float sr[4] = {0.f, 0.f, 0.f, 0.f}, srx[4] = {0.f, 0.f, 0.f, 0.f};        // grad wrt the affine output
        for (int i = 0; i < nstep; ++i) {
            const float4 d = __ldg(dy4 + idx4(i)), x = __ldg(xh4 + idx4(i));
            const float dv[4] = {d.x, d.y, d.z, d.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float u = fmaf(xv[j], scale[j], shift[j]);
                const float raw = (relu && !(u > 0.f)) ? 0.f : dv[j];
                const float gx = raw * scale[j];
                sr[j] += raw; srx[j] += raw * xv[j];
                sg[j] += gx; sgx[j] += gx * xv[j];
            }
        }
        float tot_g = 0.f, tot_gx = 0.f;
        if (p.norm == IINS_NORM_LN) {
            tot_g = iins_warp_sum(sg[0] + sg[1] + sg[2] + sg[3]);
            tot_gx = iins_warp_sum(sgx[0] + sgx[1] + sgx[2] + sgx[3]);
        }
        if (C < 4) {                                    // fold the elements of a lane that share a channel (uniform branch)
            // elements j < nch hold distinct channels and their index sets {i : ch[i] == ch[j]} are disjoint: in place is safe
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j >= nch) continue;
                float a = 0.f, b2 = 0.f, c2 = 0.f, d2 = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) if (ch[i] == ch[j]) { a += sg[i]; b2 += sgx[i]; c2 += sr[i]; d2 += srx[i]; }
                sg[j] = a; sgx[j] = b2; sr[j] = c2; srx[j] = d2;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < nch) continue;
                sg[j] = sg[j & (C - 1)]; sgx[j] = sgx[j & (C - 1)]; sr[j] = sr[j & (C - 1)]; srx[j] = srx[j & (C - 1)];
            }
        }
        // per-channel totals: combine the lanes that share this lane's column group (lane bits >= log2(CG))
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            for (int o = 16; o >= CG; o >>= 1) {
                sg[j] += __shfl_xor_sync(0xffffffffu, sg[j], o);
                sgx[j] += __shfl_xor_sync(0xffffffffu, sgx[j], o);
                sr[j] += __shfl_xor_sync(0xffffffffu, sr[j], o);
                srx[j] += __shfl_xor_sync(0xffffffffu, srx[j], o);
            }
        }
        if (p.norm == IINS_NORM_LN) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { tot_dg[j] += srx[j]; tot_db[j] += sr[j]; }
        } else if (p.norm == IINS_NORM_ADAIN && p.dadain != nullptr) {
            if (lane < CG) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (j >= nch) continue;
                    p.dadain[(long)b * p.adain_ld + p.adain_off_b + ch[j]] = sr[j];
                    p.dadain[(long)b * p.adain_ld + p.adain_off_w + ch[j]] = srx[j];
                }
            }
        }
        float4* dz4 = reinterpret_cast<float4*>(p.dz + (long)b * L * p.ldc);
        float rs[4];
        float coef = 0.f, mean_g = 0.f;
        if (p.norm == IINS_NORM_LN) {
            const float r = __ldg(p.rstd + b);
            const float sd = 1.0f / r - IINS_EPS;
            const int nel = L * C;
            rs[0] = rs[1] = rs[2] = rs[3] = r;
            coef = tot_gx / ((float)(nel - 1) * sd);
            mean_g = tot_g / (float)nel;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) rs[j] = __ldg(p.rstd + (long)b * p.rstd_ld + ch[j]);
        }
        for (int i = 0; i < nstep; ++i) {
            const float4 d = __ldg(dy4 + idx4(i)), x = __ldg(xh4 + idx4(i));
            const float dv[4] = {d.x, d.y, d.z, d.w}, xv[4] = {x.x, x.y, x.z, x.w};
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float u = fmaf(xv[j], scale[j], shift[j]);
                const float gx = ((relu && !(u > 0.f)) ? 0.f : dv[j]) * scale[j];
                if (p.norm == IINS_NORM_LN) o[j] = rs[j] * (gx - mean_g) - xv[j] * coef;
                else o[j] = rs[j] * (gx - sg[j] * invL - xv[j] * sgx[j] * invL);
            }
            dz4[idx4(i)] = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
    if (p.norm == IINS_NORM_LN) {               // uniform for the whole CTA
        if (lane < CG) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { if (j < nch) { s_dg[warp][ch[j]] = tot_dg[j]; s_db[warp][ch[j]] = tot_db[j]; } }
        }
        __syncthreads();
        if (threadIdx.x < C) {
            float a = 0.f, bsum = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) { a += s_dg[w][threadIdx.x]; bsum += s_db[w][threadIdx.x]; }
            atomicAdd(p.dgamma + threadIdx.x, a);
            atomicAdd(p.dbeta + threadIdx.x, bsum);
        }
    }
}

// ------------------------------------------------------------------------ LayerNorm forward (split path)
// The reference's custom LayerNorm (models.py:976-985) normalises over ALL (C, L) of a sample.  When the convolution that
// feeds it is wider than one GEMM column block (C > 64: dim >= 8) the statistics span several CTAs of the GEMM, so the
// GEMM writes z = conv + bias and this kernel finishes the layer: per-sample mean, UNBIASED std, xhat = (z - mean) / (std + eps),
// y = relu(xhat * gamma[c] + beta[c]).  One warp per sample, three passes over the (L2-resident) sample; z may alias xhat.
struct IinsLnFwdParams {
    int B, L, C;               // C a power of two in [4, 128], L * C a multiple of 128
    const float* z;            // (B, L, C) pre-norm values
    const float* gamma; const float* beta;
    float* y; float* xhat; float* rstd;     // rstd: [B] = 1 / (std + eps)
    int relu;
};

static __global__ void __launch_bounds__(256) iins_ln_fwd_kernel(const IinsLnFwdParams p) {
    iins_pdl_enter();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nel = p.L * p.C, nstep = nel >> 7;
    const int c0 = (lane & ((p.C >> 2) - 1)) * 4;            // C/4 <= 32 column groups: the lane's 4 channels are fixed
    float g[4], bt[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { g[j] = __ldg(p.gamma + c0 + j); bt[j] = __ldg(p.beta + c0 + j); }
    for (int b = blockIdx.x * 8 + warp; b < p.B; b += gridDim.x * 8) {
        const float4* z4 = reinterpret_cast<const float4*>(p.z + (long)b * nel);
        float sum = 0.f;
        for (int i = 0; i < nstep; ++i) { const float4 v = z4[lane + 32 * i]; sum += (v.x + v.y) + (v.z + v.w); }
        const float mean = iins_warp_sum(sum) / (float)nel;
        float sq = 0.f;
        for (int i = 0; i < nstep; ++i) {
            const float4 v = z4[lane + 32 * i];
            const float a = v.x - mean, c = v.y - mean, d = v.z - mean, e = v.w - mean;
            sq = fmaf(a, a, sq); sq = fmaf(c, c, sq); sq = fmaf(d, d, sq); sq = fmaf(e, e, sq);
        }
        const float rs = 1.0f / (sqrtf(iins_warp_sum(sq) / (float)(nel - 1)) + IINS_EPS);
        if (lane == 0) p.rstd[b] = rs;
        float4* xh4 = reinterpret_cast<float4*>(p.xhat + (long)b * nel);
        float4* y4 = reinterpret_cast<float4*>(p.y + (long)b * nel);
        for (int i = 0; i < nstep; ++i) {
            const float4 v = z4[lane + 32 * i];
            const float4 xh = make_float4((v.x - mean) * rs, (v.y - mean) * rs, (v.z - mean) * rs, (v.w - mean) * rs);
            float4 o = make_float4(fmaf(xh.x, g[0], bt[0]), fmaf(xh.y, g[1], bt[1]), fmaf(xh.z, g[2], bt[2]), fmaf(xh.w, g[3], bt[3]));
            if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            xh4[lane + 32 * i] = xh;
            y4[lane + 32 * i] = o;
        }
    }
}

// -------------------------------------------------------------------- adaptive avg pooling (C = 1)
// window i of AdaptiveAvgPool1d(Lin -> Lout): [floor(i*Lin/Lout), ceil((i+1)*Lin/Lout))  (models.py:146,436)
IINS_HD void iins_pool_window(int i, int Lin, int Lout, int& s, int& e) {
    s = (i * Lin) / Lout;
    e = ((i + 1) * Lin + Lout - 1) / Lout;
}

static __global__ void __launch_bounds__(256) iins_pool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                            int B, int Lin, int Lout) {
    iins_pdl_enter();
    long n = (long)B * Lout;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        int b = (int)(i / Lout), o = (int)(i - (long)b * Lout);
        int s, e;
        iins_pool_window(o, Lin, Lout, s, e);
        float acc = 0.f;
        for (int j = s; j < e; ++j) acc += __ldg(x + (long)b * Lin + j);
        y[i] = acc / (float)(e - s);
    }
}

// dx[b,j] = sum over windows containing j of dy[b,o]/len(o); optionally times (1 - t^2) with t = tanh output
static __global__ void __launch_bounds__(256) iins_pool_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ tanh_y,
                                                            float* __restrict__ dx, int B, int Lin, int Lout) {
    iins_pdl_enter();
    long n = (long)B * Lin;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        int b = (int)(i / Lin), j = (int)(i - (long)b * Lin);
        // candidate windows: o in [floor(j*Lout/Lin) - 1, ceil((j+1)*Lout/Lin)]
        int lo = (int)(((long)j * Lout) / Lin) - 1;
        int hi = (int)(((long)(j + 1) * Lout + Lin - 1) / Lin) + 1;
        if (lo < 0) lo = 0;
        if (hi > Lout) hi = Lout;
        float acc = 0.f;
        for (int o = lo; o < hi; ++o) {
            int s, e;
            iins_pool_window(o, Lin, Lout, s, e);
            if (j >= s && j < e) acc += __ldg(dy + (long)b * Lout + o) / (float)(e - s);
        }
        if (tanh_y != nullptr) { float t = __ldg(tanh_y + i); acc *= (1.f - t * t); }
        dx[i] = acc;
    }
}

// mean over L of an NLC tensor: (B,L,C) -> (B,C)   (AdaptiveAvgPool1d(1), models.py:279)
static __global__ void __launch_bounds__(256) iins_mean_l_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                          int B, int L, int C) {
    iins_pdl_enter();
    long n = (long)B * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        int b = (int)(i / C), c = (int)(i - (long)b * C);
        float acc = 0.f;
        for (int l = 0; l < L; ++l) acc += __ldg(x + ((long)b * L + l) * C + c);
        y[i] = acc / (float)L;
    }
}

// --------------------------------------------------------------------------- Philox4x32-10
struct IinsPhilox { unsigned c[4]; };
IINS_HD unsigned iins_mulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
IINS_HD IinsPhilox iins_philox(unsigned long long seed, unsigned long long idx, unsigned long long offset) {
    unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
    unsigned c0 = (unsigned)idx, c1 = (unsigned)(idx >> 32), c2 = (unsigned)offset, c3 = (unsigned)(offset >> 32);
    for (int r = 0; r < 10; ++r) {
        unsigned hi0 = iins_mulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        unsigned hi1 = iins_mulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    IinsPhilox o; o.c[0] = c0; o.c[1] = c1; o.c[2] = c2; o.c[3] = c3;
    return o;
}
// standard normal from two 32-bit words (Box-Muller)
IINS_HD float iins_normal(unsigned a, unsigned b) {
    float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
    return sqrtf(-2.0f * logf(u1)) * cosf(6.28318530717958647692f * u2);
}
IINS_HD float iins_noise_at(unsigned long long seed, unsigned long long offset, long b, int j) {
    // one Philox block (4 words = 2 normals) per (sample, latent pair); offset advances per step
    IinsPhilox r = iins_philox(seed, (unsigned long long)b, offset + (unsigned long long)(j >> 1));
    return (j & 1) ? iins_normal(r.c[2], r.c[3]) : iins_normal(r.c[0], r.c[1]);
}

// --------------------------------------------------------------- reparameterisation + KL
// cat (B,E) = [mu | log_sigma];  latent = noise*exp(ls)+mu;  kl = mean_b 0.5*sum(exp(2ls)+mu^2-1-2ls)
// (models.py:285-298).  noise: explicit (B,E/2) tensor if given, else Philox(seed, offset).
static __global__ void __launch_bounds__(256) iins_reparam_kl_kernel(const float* __restrict__ cat, const float* __restrict__ noise,
                                                              float* __restrict__ latent, float* __restrict__ kl,
                                                              int B, int E, unsigned long long seed, unsigned long long offset) {
    iins_pdl_enter();
    const int H = E / 2;
    float part = 0.f;
    long n = (long)B * H;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        long b = i / H;
        int j = (int)(i - b * H);
        float mu = __ldg(cat + b * E + j), ls = __ldg(cat + b * E + H + j);
        if (latent != nullptr) {
            float nz = noise != nullptr ? __ldg(noise + i) : iins_noise_at(seed, offset, b, j);
            latent[i] = fmaf(nz, expf(ls), mu);
        }
        part += 0.5f * (expf(2.f * ls) + mu * mu - 1.f - 2.f * ls);
    }
    part = iins_warp_sum(part);
    if ((threadIdx.x & 31) == 0) atomicAdd(kl, part / (float)B);
}

// dcat = d_cat_in + d_kl * dkl/dcat + latent-path terms
static __global__ void __launch_bounds__(256) iins_reparam_kl_bwd_kernel(const float* __restrict__ cat, const float* __restrict__ noise,
                                                                  const float* __restrict__ d_cat_in, const float* __restrict__ d_latent,
                                                                  const float* __restrict__ d_kl, float* __restrict__ dcat,
                                                                  int B, int E, unsigned long long seed, unsigned long long offset) {
    iins_pdl_enter();
    const int H = E / 2;
    const float gk = d_kl != nullptr ? __ldg(d_kl) / (float)B : 0.f;
    long n = (long)B * H;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        long b = i / H;
        int j = (int)(i - b * H);
        float mu = __ldg(cat + b * E + j), ls = __ldg(cat + b * E + H + j);
        float dmu = gk * mu;
        float dls = gk * (expf(2.f * ls) - 1.f);
        if (d_latent != nullptr) {
            float nz = noise != nullptr ? __ldg(noise + i) : iins_noise_at(seed, offset, b, j);
            float dl = __ldg(d_latent + i);
            dmu += dl;
            dls += dl * nz * expf(ls);
        }
        if (d_cat_in != nullptr) { dmu += __ldg(d_cat_in + b * E + j); dls += __ldg(d_cat_in + b * E + H + j); }
        dcat[b * E + j] = dmu;
        dcat[b * E + H + j] = dls;
    }
}

// ---------------------------------------------------------------------------------- fused loss
// One pass over the batch computing the terms of train_semi.py:199-225 (or train.py:87-91), the
// seed gradients of every head output and the metrics of train.py:104-115.
//   out[0] = mean|x - xrec|      (L1Loss over B*L)          out[4] = sum (err_est-err)^2 / B  (MSE; rmse = sqrt)
//   out[1] = mean|err - err_est|                            out[5] = number of correct argmax predictions
//   out[2] = mean CE(logits, label)                         out[6] = number of labels outside [0, NC) after the offset
//                                                           (torch's CrossEntropyLoss device-asserts on those); out[7] reserved
//   out[3] = lam_ae*out[0] + lam_res*out[1] + lam_env*out[2]   (KL is added by the caller: it lives in the encoder)
struct IinsLossParams {
    int B, L, NC;
    const float* x;            // (B,L) or nullptr (no reconstruction term)
    const float* xrec;
    const float* err;          // (B,) or nullptr (unsupervised batch)
    const float* err_est;
    const float* logits;       // (B,NC)
    const float* label;        // (B,) float32 holding integers (dataset.py:122) ...
    const long long* label_i64;   // ... or int64 (train.py:72); exactly one of the two
    int label_offset;          // class index = label - label_offset (train_semi.py:217-222: 1 for every dataset_env but room_full)
    float lam_ae, lam_res, lam_env;
    float* out;                // 8 floats, zeroed by the caller
    float* d_xrec;             // (B,L)
    float* d_err_est;          // (B,)
    float* d_logits;           // (B,NC)
    int* pred;                 // (B,) argmax or nullptr
};

static __global__ void __launch_bounds__(256) iins_loss_kernel(const IinsLossParams p) {
    iins_pdl_enter();
    __shared__ float s_part[8][8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float acc_ae = 0.f, acc_res = 0.f, acc_ce = 0.f, acc_sq = 0.f, acc_ok = 0.f, acc_bad = 0.f;
    const long stride = (long)gridDim.x * blockDim.x;
    const long gtid = (long)blockIdx.x * blockDim.x + tid;
    if (p.x != nullptr) {
        const long n = (long)p.B * p.L;
        const float ginv = p.lam_ae / (float)n;
        // 128-bit accesses over the flattened (B*L) stream whenever the three buffers are 16-byte aligned
        const bool vec = ((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.xrec) | reinterpret_cast<uintptr_t>(p.d_xrec)) & 15) == 0;
        const long n4 = vec ? n >> 2 : 0;
#pragma unroll 4
        for (long i = gtid; i < n4; i += stride) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p.xrec) + i), b = __ldg(reinterpret_cast<const float4*>(p.x) + i);
            const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
            acc_ae += (fabsf(d0) + fabsf(d1)) + (fabsf(d2) + fabsf(d3));
            if (p.d_xrec != nullptr)
                reinterpret_cast<float4*>(p.d_xrec)[i] = make_float4(d0 > 0.f ? ginv : (d0 < 0.f ? -ginv : 0.f), d1 > 0.f ? ginv : (d1 < 0.f ? -ginv : 0.f),
                                                                     d2 > 0.f ? ginv : (d2 < 0.f ? -ginv : 0.f), d3 > 0.f ? ginv : (d3 < 0.f ? -ginv : 0.f));
        }
        for (long i = 4 * n4 + gtid; i < n; i += stride) {
            float d = __ldg(p.xrec + i) - __ldg(p.x + i);
            acc_ae += fabsf(d);
            if (p.d_xrec != nullptr) p.d_xrec[i] = d > 0.f ? ginv : (d < 0.f ? -ginv : 0.f);
        }
    }
    if (p.err != nullptr) {
        for (long b = gtid; b < p.B; b += stride) {
            float d = __ldg(p.err_est + b) - __ldg(p.err + b);
            acc_res += fabsf(d);
            acc_sq += d * d;
            if (p.d_err_est != nullptr) {
                float gi = p.lam_res / (float)p.B;
                p.d_err_est[b] = d > 0.f ? gi : (d < 0.f ? -gi : 0.f);
            }
            int tgt = (p.label_i64 != nullptr ? (int)p.label_i64[b] : (int)__ldg(p.label + b)) - p.label_offset;
            if (tgt < 0 || tgt >= p.NC) { acc_bad += 1.f; tgt = tgt < 0 ? 0 : p.NC - 1; }      // reported in out[6]; never read outside the row
            const float* z = p.logits + b * p.NC;
            float mx = __ldg(z);
            int am = 0;
            for (int c = 1; c < p.NC; ++c) { float v = __ldg(z + c); if (v > mx) { mx = v; am = c; } }
            float se = 0.f;
            for (int c = 0; c < p.NC; ++c) se += expf(__ldg(z + c) - mx);
            float lse = mx + logf(se);
            acc_ce += lse - __ldg(z + tgt);
            acc_ok += (am == tgt) ? 1.f : 0.f;
            if (p.pred != nullptr) p.pred[b] = am;
            if (p.d_logits != nullptr) {
                float gi = p.lam_env / (float)p.B;
                for (int c = 0; c < p.NC; ++c) {
                    float sm = expf(__ldg(z + c) - lse);
                    p.d_logits[b * p.NC + c] = gi * (sm - (c == tgt ? 1.f : 0.f));
                }
            }
        }
    }
    float vals[6] = {acc_ae, acc_res, acc_ce, acc_sq, acc_ok, acc_bad};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        float v = iins_warp_sum(vals[k]);
        if (lane == 0) s_part[k][warp] = v;
    }
    __syncthreads();
    if (tid < 6) {
        float v = 0.f;
        for (int w = 0; w < 8; ++w) v += s_part[tid][w];
        float nae = p.x != nullptr ? (float)((long)p.B * p.L) : 1.f;
        if (tid == 0) { v /= nae; atomicAdd(p.out + 0, v); atomicAdd(p.out + 3, p.lam_ae * v); }
        if (tid == 1) { v /= (float)p.B; atomicAdd(p.out + 1, v); atomicAdd(p.out + 3, p.lam_res * v); }
        if (tid == 2) { v /= (float)p.B; atomicAdd(p.out + 2, v); atomicAdd(p.out + 3, p.lam_env * v); }
        if (tid == 3) { v /= (float)p.B; atomicAdd(p.out + 4, v); }
        if (tid == 4) { atomicAdd(p.out + 5, v); }
        if (tid == 5 && v != 0.f) { atomicAdd(p.out + 6, v); }
    }
}

// dst1 += src1, dst2 += src2 in one launch: sums the head gradients (Restorer -> d range_code, Classifier -> d env_code)
// into the decoder's when the engine runs the heads concurrently with the decoder (engine.py)
static __global__ void __launch_bounds__(256) iins_accumulate2_kernel(float* __restrict__ dst1, const float* __restrict__ src1, long n1,
                                                               float* __restrict__ dst2, const float* __restrict__ src2, long n2) {
    iins_pdl_enter();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2; i += (long)gridDim.x * blockDim.x) {
        if (i < n1) dst1[i] += __ldg(src1 + i);
        else dst2[i - n1] += __ldg(src2 + i - n1);
    }
}

// ---------------------------------------------------------------------------------- fused Adam
// torch.optim.Adam semantics (train_semi.py:118-122): per-group step counters live on the device so a
// captured CUDA graph can replay the update; a group whose gradients are "None" this step (Res/Cls on
// an unsupervised batch, restorer.linear_layer2 always) is skipped entirely: no moment decay, no step++.
struct IinsAdamGroup { long begin, end; int active; };
struct IinsAdamParams {
    float* p; float* g; float* m; float* v;
    float grad_scale;          // g is multiplied by this first (1/world after a SUM all-reduce: the mean of the per-rank gradients)
    int zero_grads;            // write 0 to every gradient element that was consumed (the next step accumulates into it)
    int* steps;                // [n_groups + 1] device counters of COMPLETED updates (+ a ticket word, zero between launches): this
                               // launch uses steps[g] + 1 and its last CTA to finish advances the active groups
    const float* lr;           // device scalar (LambdaLR changes it per epoch)
    double beta1, beta2;       // bias corrections are formed in double like torch's Python scalars
    float eps;
    int n_groups;
    IinsAdamGroup groups[8];
};

static __global__ void __launch_bounds__(256) iins_adam_kernel(const IinsAdamParams a) {
    iins_pdl_enter();
    // the bias corrections are double-precision pow() like torch's Python scalars: evaluated by ONE thread per CTA
    // and group (they cost hundreds of fp64 instructions), then shared
    __shared__ float s_step[8], s_isb2[8];
    const float lr = __ldg(a.lr);
    if (threadIdx.x < a.n_groups && a.groups[threadIdx.x].active) {
        const int t = a.steps[threadIdx.x] + 1;
        const double bc1 = 1.0 - pow(a.beta1, (double)t);
        const double bc2 = 1.0 - pow(a.beta2, (double)t);
        s_step[threadIdx.x] = (float)((double)lr / bc1);
        s_isb2[threadIdx.x] = (float)(1.0 / sqrt(bc2));
    }
    __syncthreads();
    const long stride = (long)gridDim.x * blockDim.x;
    const long gtid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const float b1 = (float)a.beta1, b2 = (float)a.beta2;
    for (int gi = 0; gi < a.n_groups; ++gi) {
        if (!a.groups[gi].active) continue;
        const float step_size = s_step[gi], inv_sqrt_bc2 = s_isb2[gi];
        const long begin = a.groups[gi].begin, end = a.groups[gi].end;
        const float gs = a.grad_scale;
        auto upd = [&](float g, float& m, float& v, float& w) {
            g *= gs;
            m = b1 * m + (1.f - b1) * g;
            v = b2 * v + (1.f - b2) * g * g;
            const float denom = sqrtf(v) * inv_sqrt_bc2 + a.eps;
            w -= step_size * (m / denom);
        };
        // head up to the first 16-byte boundary, 128-bit body, scalar tail (the four flat buffers share their alignment
        // only if they were allocated alike: checked)
        const bool vec = ((reinterpret_cast<uintptr_t>(a.p) | reinterpret_cast<uintptr_t>(a.g) | reinterpret_cast<uintptr_t>(a.m) |
                           reinterpret_cast<uintptr_t>(a.v)) & 15) == 0;
        long vb = vec ? (begin + 3) & ~3L : end, ve = vec ? end & ~3L : end;
        if (vb > ve) { vb = end; ve = end; }
        const bool zg = a.zero_grads != 0;
        for (long i = begin + gtid; i < vb; i += stride) { float m = a.m[i], v = a.v[i], w = a.p[i]; upd(a.g[i], m, v, w); a.m[i] = m; a.v[i] = v; a.p[i] = w; if (zg) a.g[i] = 0.f; }
        for (long i = (vb >> 2) + gtid; i < (ve >> 2); i += stride) {
            const float4 g4 = reinterpret_cast<const float4*>(a.g)[i];
            float4 m4 = reinterpret_cast<float4*>(a.m)[i], v4 = reinterpret_cast<float4*>(a.v)[i], w4 = reinterpret_cast<float4*>(a.p)[i];
            upd(g4.x, m4.x, v4.x, w4.x); upd(g4.y, m4.y, v4.y, w4.y); upd(g4.z, m4.z, v4.z, w4.z); upd(g4.w, m4.w, v4.w, w4.w);
            reinterpret_cast<float4*>(a.m)[i] = m4; reinterpret_cast<float4*>(a.v)[i] = v4; reinterpret_cast<float4*>(a.p)[i] = w4;
            if (zg) reinterpret_cast<float4*>(a.g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (long i = ve + gtid; i < end; i += stride) { float m = a.m[i], v = a.v[i], w = a.p[i]; upd(a.g[i], m, v, w); a.m[i] = m; a.v[i] = v; a.p[i] = w; if (zg) a.g[i] = 0.f; }
    }
    // every CTA has read the step counters above; the last one to get here advances them (one launch instead of two)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int ticket = atomicAdd(a.steps + a.n_groups, 1);
        if (ticket == (int)gridDim.x - 1) {
            for (int gi = 0; gi < a.n_groups; ++gi) if (a.groups[gi].active) a.steps[gi] += 1;
            a.steps[a.n_groups] = 0;
        }
    }
}
