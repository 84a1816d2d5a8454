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
};

__global__ void __launch_bounds__(256) iins_norm_bwd_kernel(const IinsNormBwdParams p) {
    __shared__ float s_g[1024];        // sum_l g        per (s,c)   | LN: per sample [0..7]
    __shared__ float s_gx[1024];       // sum_l g*xhat
    const int tid = threadIdx.x;
    const int L = p.L, C = p.C;
    const int S = 128 / L;             // samples per tile
    const int b0 = blockIdx.x * S;
    const bool relu = p.act == IINS_ACT_RELU;

    // g(b,l,c) = dy * mask * scale   and the value u whose sign is the ReLU mask
    auto load_g = [&](int b, int l, int c, float& xh) -> float {
        long i = ((long)b * L + l) * C + c;
        xh = __ldg(p.xhat + i);
        float g = __ldg(p.dy + i);
        float u = xh, scale = 1.f;
        if (p.norm == IINS_NORM_ADAIN) {
            scale = __ldg(p.adain + (long)b * p.adain_ld + p.adain_off_w + c);
            u = fmaf(xh, scale, __ldg(p.adain + (long)b * p.adain_ld + p.adain_off_b + c));
        } else if (p.norm == IINS_NORM_LN) {
            scale = __ldg(p.gamma + c);
            u = fmaf(xh, scale, __ldg(p.beta + c));
        }
        if (relu && !(u > 0.f)) g = 0.f;
        return g * scale;              // gradient w.r.t. xhat; (g w.r.t. affine output) = value / scale
    };

    if (p.norm == IINS_NORM_IN || p.norm == IINS_NORM_ADAIN) {
        const int pairs = S * C;
        int G = 1;
        while (G < 32 && pairs * G * 2 <= 256) G <<= 1;
        const int per_iter = 256 / G;
        for (int p0 = 0; p0 < pairs; p0 += per_iter) {
            int pr = p0 + tid / G, sub = tid % G;
            int s = pr / C, c = pr - s * C;
            int b = b0 + s;
            bool ok = pr < pairs && b < p.B;
            float sg = 0.f, sgx = 0.f, sraw = 0.f, srawx = 0.f;
            if (ok) {
                float scale = p.norm == IINS_NORM_ADAIN ? __ldg(p.adain + (long)b * p.adain_ld + p.adain_off_w + c) : 1.f;
                for (int l = sub; l < L; l += G) {
                    float xh;
                    float g = load_g(b, l, c, xh);
                    sg += g;
                    sgx += g * xh;
                    if (p.norm == IINS_NORM_ADAIN) {
                        // raw = dy*mask (gradient w.r.t. the affine output): recompute without the scale
                        long i = ((long)b * L + l) * C + c;
                        float u = fmaf(xh, scale, __ldg(p.adain + (long)b * p.adain_ld + p.adain_off_b + c));
                        float raw = (relu && !(u > 0.f)) ? 0.f : __ldg(p.dy + i);
                        sraw += raw;
                        srawx += raw * xh;
                    }
                }
            }
            sg = iins_group_sum(sg, G);
            sgx = iins_group_sum(sgx, G);
            if (p.norm == IINS_NORM_ADAIN) {
                sraw = iins_group_sum(sraw, G);
                srawx = iins_group_sum(srawx, G);
            }
            if (ok && sub == 0) {
                s_g[s * C + c] = sg;
                s_gx[s * C + c] = sgx;
                if (p.norm == IINS_NORM_ADAIN && p.dadain != nullptr) {
                    p.dadain[(long)b * p.adain_ld + p.adain_off_b + c] = sraw;
                    p.dadain[(long)b * p.adain_ld + p.adain_off_w + c] = srawx;
                }
            }
        }
        __syncthreads();
        const float invL = 1.0f / (float)L;
        for (int e = tid; e < 128 * C; e += 256) {
            int r = e / C, c = e - r * C;
            int s = r / L, l = r - s * L;
            int b = b0 + s;
            if (b >= p.B) continue;
            float xh;
            float g = load_g(b, l, c, xh);
            float rs = __ldg(p.rstd + (long)b * C + c);
            p.dz[((long)b * L + l) * C + c] = rs * (g - s_g[s * C + c] * invL - xh * s_gx[s * C + c] * invL);
        }
    } else {   // LN
        const int warp = tid >> 5, lane = tid & 31;
        const int nel = L * C;
        for (int s0 = 0; s0 < S; s0 += 8) {
            int s = s0 + warp;
            int b = b0 + s;
            bool ok = s < S && b < p.B;
            float sg = 0.f, sgx = 0.f;
            if (ok) for (int e = lane; e < nel; e += 32) {
                float xh;
                float g = load_g(b, e / C, e % C, xh);
                sg += g;
                sgx += g * xh;
            }
            sg = iins_warp_sum(sg);
            sgx = iins_warp_sum(sgx);
            if (ok && lane == 0) { s_g[s] = sg; s_gx[s] = sgx; }
        }
        __syncthreads();
        // per-channel dgamma / dbeta over every row of the tile: thread c (< C) loops the rows
        // (C <= 64; rows 128) -- small next to the conv work of the layer.
        for (int c = tid; c < C; c += 256) {
            float dg = 0.f, dbt = 0.f;
            float gm = __ldg(p.gamma + c), bt = __ldg(p.beta + c);
            for (int r = 0; r < 128; ++r) {
                int s = r / L, l = r - s * L;
                int b = b0 + s;
                if (b >= p.B) break;
                long i = ((long)b * L + l) * C + c;
                float xh = __ldg(p.xhat + i);
                float u = fmaf(xh, gm, bt);
                float raw = (relu && !(u > 0.f)) ? 0.f : __ldg(p.dy + i);
                dg += raw * xh;
                dbt += raw;
            }
            atomicAdd(p.dgamma + c, dg);
            atomicAdd(p.dbeta + c, dbt);
        }
        const float inv_n = 1.0f / (float)nel;
        for (int e = tid; e < 128 * C; e += 256) {
            int r = e / C, c = e - r * C;
            int s = r / L, l = r - s * L;
            int b = b0 + s;
            if (b >= p.B) continue;
            float xh;
            float g = load_g(b, l, c, xh);
            float rs = __ldg(p.rstd + b);
            float sd = 1.0f / rs - IINS_EPS;
            p.dz[((long)b * L + l) * C + c] = rs * (g - s_g[s] * inv_n) - xh * s_gx[s] / ((float)(nel - 1) * sd);
        }
    }
}

// -------------------------------------------------------------------- adaptive avg pooling (C = 1)
// window i of AdaptiveAvgPool1d(Lin -> Lout): [floor(i*Lin/Lout), ceil((i+1)*Lin/Lout))  (models.py:146,436)
IINS_HD void iins_pool_window(int i, int Lin, int Lout, int& s, int& e) {
    s = (i * Lin) / Lout;
    e = ((i + 1) * Lin + Lout - 1) / Lout;
}

__global__ void __launch_bounds__(256) iins_pool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                            int B, int Lin, int Lout) {
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
__global__ void __launch_bounds__(256) iins_pool_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ tanh_y,
                                                            float* __restrict__ dx, int B, int Lin, int Lout) {
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
__global__ void __launch_bounds__(256) iins_mean_l_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                          int B, int L, int C) {
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
__global__ void __launch_bounds__(256) iins_reparam_kl_kernel(const float* __restrict__ cat, const float* __restrict__ noise,
                                                              float* __restrict__ latent, float* __restrict__ kl,
                                                              int B, int E, unsigned long long seed, unsigned long long offset) {
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
__global__ void __launch_bounds__(256) iins_reparam_kl_bwd_kernel(const float* __restrict__ cat, const float* __restrict__ noise,
                                                                  const float* __restrict__ d_cat_in, const float* __restrict__ d_latent,
                                                                  const float* __restrict__ d_kl, float* __restrict__ dcat,
                                                                  int B, int E, unsigned long long seed, unsigned long long offset) {
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
//   out[2] = mean CE(logits, label)                         out[6], out[7] reserved
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
    float lam_ae, lam_res, lam_env;
    float* out;                // 8 floats, zeroed by the caller
    float* d_xrec;             // (B,L)
    float* d_err_est;          // (B,)
    float* d_logits;           // (B,NC)
    int* pred;                 // (B,) argmax or nullptr
};

__global__ void __launch_bounds__(256) iins_loss_kernel(const IinsLossParams p) {
    __shared__ float s_part[8][8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float acc_ae = 0.f, acc_res = 0.f, acc_ce = 0.f, acc_sq = 0.f, acc_ok = 0.f;
    const long stride = (long)gridDim.x * blockDim.x;
    const long gtid = (long)blockIdx.x * blockDim.x + tid;
    if (p.x != nullptr) {
        const long n = (long)p.B * p.L;
        const float ginv = p.lam_ae / (float)n;
        for (long i = gtid; i < n; i += stride) {
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
            int tgt = p.label_i64 != nullptr ? (int)p.label_i64[b] : (int)__ldg(p.label + b);
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
    float vals[5] = {acc_ae, acc_res, acc_ce, acc_sq, acc_ok};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        float v = iins_warp_sum(vals[k]);
        if (lane == 0) s_part[k][warp] = v;
    }
    __syncthreads();
    if (tid < 5) {
        float v = 0.f;
        for (int w = 0; w < 8; ++w) v += s_part[tid][w];
        float nae = p.x != nullptr ? (float)((long)p.B * p.L) : 1.f;
        if (tid == 0) { v /= nae; atomicAdd(p.out + 0, v); atomicAdd(p.out + 3, p.lam_ae * v); }
        if (tid == 1) { v /= (float)p.B; atomicAdd(p.out + 1, v); atomicAdd(p.out + 3, p.lam_res * v); }
        if (tid == 2) { v /= (float)p.B; atomicAdd(p.out + 2, v); atomicAdd(p.out + 3, p.lam_env * v); }
        if (tid == 3) { v /= (float)p.B; atomicAdd(p.out + 4, v); }
        if (tid == 4) { atomicAdd(p.out + 5, v); }
    }
}

// ---------------------------------------------------------------------------------- fused Adam
// torch.optim.Adam semantics (train_semi.py:118-122): per-group step counters live on the device so a
// captured CUDA graph can replay the update; a group whose gradients are "None" this step (Res/Cls on
// an unsupervised batch, restorer.linear_layer2 always) is skipped entirely: no moment decay, no step++.
struct IinsAdamGroup { long begin, end; int active; };
struct IinsAdamParams {
    float* p; const float* g; float* m; float* v;
    int* steps;                // [n_groups] device counters (already incremented for this step)
    const float* lr;           // device scalar (LambdaLR changes it per epoch)
    double beta1, beta2;       // bias corrections are formed in double like torch's Python scalars
    float eps;
    int n_groups;
    IinsAdamGroup groups[8];
};

__global__ void iins_adam_tick_kernel(int* steps, int n_groups, unsigned active_mask) {
    int i = threadIdx.x;
    if (i < n_groups && ((active_mask >> i) & 1u)) steps[i] += 1;
}

__global__ void __launch_bounds__(256) iins_adam_kernel(const IinsAdamParams a) {
    const long stride = (long)gridDim.x * blockDim.x;
    const float lr = __ldg(a.lr);
    for (int gi = 0; gi < a.n_groups; ++gi) {
        if (!a.groups[gi].active) continue;
        const int t = a.steps[gi];
        const double bc1 = 1.0 - pow(a.beta1, (double)t);
        const double bc2 = 1.0 - pow(a.beta2, (double)t);
        const float step_size = (float)((double)lr / bc1);
        const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
        const float b1 = (float)a.beta1, b2 = (float)a.beta2;
        for (long i = a.groups[gi].begin + (long)blockIdx.x * blockDim.x + threadIdx.x; i < a.groups[gi].end; i += stride) {
            float g = __ldg(a.g + i);
            float m = a.m[i] = b1 * a.m[i] + (1.f - b1) * g;
            float v = a.v[i] = b2 * a.v[i] + (1.f - b2) * g * g;
            float denom = sqrtf(v) * inv_sqrt_bc2 + a.eps;
            a.p[i] -= step_size * (m / denom);
        }
    }
}
