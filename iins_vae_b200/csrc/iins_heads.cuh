// Kernels of the Conv1d heads (models.py:661-716 RestorerConv1d, :865-902 ClassifierConv1d) that are not convolutions:
// Dropout(0.25) with an explicit or Philox keep-mask, BatchNorm1d(eps = 0.8) forward / backward over the batch.
// The convolutions / the Linear output go through the generic implicit-GEMM layer kernels (iins_runtime.cu).
// Activations are channels-last (B, L, C); masks and Philox counters follow the REFERENCE's (B, C, L) element order so that
// a mask recorded from the reference (or torch's layout in general) replays unchanged.
#pragma once
#include "iins_misc.cuh"

struct IinsDropoutParams {
    const float* x; float* y;          // (B, L, C) channels-last; y may alias x
    const float* mask;                 // (B, C, L) keep-mask of 0 / 1 or nullptr -> Philox4x32-10(seed, offset + element index)
    int B, L, C;
    float p;                           // drop probability; kept values are scaled by 1 / (1 - p)
    unsigned long long seed, offset;
    const int* offset_dev;             // optional device counter added to offset (CUDA-graph replays)
    long long sample_offset;           // global index of sample 0 (Philox counters only)
};

IINS_HD float iins_uniform01(unsigned a) { return ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f); }

static __global__ void __launch_bounds__(256) iins_dropout_kernel(const IinsDropoutParams p) {
    iins_pdl_enter();
    const long n = (long)p.B * p.L * p.C;
    const float scale = 1.0f / (1.0f - p.p);
    const unsigned long long off = p.offset + (p.offset_dev != nullptr ? 2ull * (unsigned long long)__ldg(p.offset_dev) : 0ull);
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        const int c = (int)(e % p.C);
        const long r = e / p.C;
        const int l = (int)(r % p.L);
        const long b = r / p.L;
        const long ncl = (b * p.C + c) * p.L + l;
        float keep;
        if (p.mask != nullptr) keep = __ldg(p.mask + ncl);
        else {
            const long gl = ncl + p.sample_offset * p.C * p.L;
            const IinsPhilox rr = iins_philox(p.seed, (unsigned long long)(gl >> 2), off);
            keep = iins_uniform01(rr.c[gl & 3]) >= p.p ? 1.f : 0.f;
        }
        p.y[e] = __ldg(p.x + e) * keep * scale;
    }
}

// ---- BatchNorm1d over the rows (B * L) of a channels-last (rows, C) tensor -------------------------------------------
// Two kernels per direction with a double-precision sum buffer in between: the caller may all-reduce that buffer across
// data-parallel ranks between the two (SyncBN, 2 * C values) -- the only place where the Conv1d heads leave the
// "every normalisation is per sample" property of the path.
struct IinsBnStatsParams {
    const float* a; const float* b;    // sums of a[r][c] and of a[r][c] * (b ? b[r][c] : a[r][c])
    long rows; int C;
    double* sums;                      // [2 * C], zeroed by the caller: [0,C) sum a, [C,2C) sum a * (b or a)
    float* g_first; float* g_second;   // optional float gradient accumulators (+= the LOCAL sums): d beta, d gamma
};

static __global__ void __launch_bounds__(256) iins_bn_stats_kernel(const IinsBnStatsParams p) {
    iins_pdl_enter();
    __shared__ double s1[256], s2[256];
    const int tid = threadIdx.x;
    const int lanes = 256 / p.C;                       // rows handled in parallel by this CTA (C <= 256)
    const int c = tid % p.C, rl = tid / p.C;
    double a1 = 0.0, a2 = 0.0;
    if (rl < lanes) {
        for (long r = (long)blockIdx.x * lanes + rl; r < p.rows; r += (long)gridDim.x * lanes) {
            const float v = __ldg(p.a + r * p.C + c);
            const float w = p.b != nullptr ? __ldg(p.b + r * p.C + c) : v;
            a1 += (double)v;
            a2 += (double)v * (double)w;
        }
    }
    s1[tid] = a1; s2[tid] = a2;
    __syncthreads();
    if (tid < p.C) {
        double t1 = 0.0, t2 = 0.0;
        for (int k = 0; k < lanes; ++k) { t1 += s1[k * p.C + tid]; t2 += s2[k * p.C + tid]; }
        atomicAdd(p.sums + tid, t1);
        atomicAdd(p.sums + p.C + tid, t2);
        if (p.g_first != nullptr) atomicAdd(p.g_first + tid, (float)t1);
        if (p.g_second != nullptr) atomicAdd(p.g_second + tid, (float)t2);
    }
}

struct IinsBnApplyParams {
    const float* x; float* xhat; float* y;   // (rows, C)
    long rows; int C;
    const double* sums;                      // forward sums (training) or nullptr (eval: running statistics)
    double count;                            // number of rows the sums cover (all ranks)
    const float* gamma; const float* beta;
    float eps, momentum;
    float* running_mean; float* running_var; long long* num_batches_tracked;   // updated in training (by CTA 0) when non-null
    float* saved;                            // [2 * C]: mean, rstd (for the backward pass)
};

static __global__ void __launch_bounds__(256) iins_bn_apply_kernel(const IinsBnApplyParams p) {
    iins_pdl_enter();
    __shared__ float s_mean[256], s_rstd[256];
    const int tid = threadIdx.x;
    if (tid < p.C) {
        double mean, var;
        if (p.sums != nullptr) {
            mean = p.sums[tid] / p.count;
            var = p.sums[p.C + tid] / p.count - mean * mean;          // biased; double precision: no cancellation issue at fp32 data
            if (var < 0.0) var = 0.0;
        } else {
            mean = (double)p.running_mean[tid];
            var = (double)p.running_var[tid];
        }
        const float rstd = (float)(1.0 / sqrt(var + (double)p.eps));
        s_mean[tid] = (float)mean; s_rstd[tid] = rstd;
        if (blockIdx.x == 0) {
            p.saved[tid] = (float)mean; p.saved[p.C + tid] = rstd;
            if (p.sums != nullptr && p.running_mean != nullptr) {
                const double unb = p.count > 1.0 ? var * p.count / (p.count - 1.0) : var;
                p.running_mean[tid] = (1.f - p.momentum) * p.running_mean[tid] + p.momentum * (float)mean;
                p.running_var[tid] = (1.f - p.momentum) * p.running_var[tid] + p.momentum * (float)unb;
                if (tid == 0 && p.num_batches_tracked != nullptr) *p.num_batches_tracked += 1;
            }
        }
    }
    __syncthreads();
    const long n = p.rows * p.C;
    for (long e = (long)blockIdx.x * blockDim.x + tid; e < n; e += (long)gridDim.x * blockDim.x) {
        const int c = (int)(e % p.C);
        const float xh = (__ldg(p.x + e) - s_mean[c]) * s_rstd[c];
        p.xhat[e] = xh;
        p.y[e] = fmaf(xh, __ldg(p.gamma + c), __ldg(p.beta + c));
    }
}

struct IinsBnBwdParams {
    const float* dy; const float* xhat; float* dx;   // (rows, C)
    long rows; int C;
    const double* sums;                      // backward sums [sum dy | sum dy * xhat] (training) or nullptr (eval)
    double count;
    const float* gamma; const float* saved;  // saved: [mean | rstd]
};

static __global__ void __launch_bounds__(256) iins_bn_bwd_kernel(const IinsBnBwdParams p) {
    iins_pdl_enter();
    __shared__ float s_m1[256], s_m2[256], s_sc[256];
    const int tid = threadIdx.x;
    if (tid < p.C) {
        s_m1[tid] = p.sums != nullptr ? (float)(p.sums[tid] / p.count) : 0.f;
        s_m2[tid] = p.sums != nullptr ? (float)(p.sums[p.C + tid] / p.count) : 0.f;
        s_sc[tid] = __ldg(p.gamma + tid) * p.saved[p.C + tid];
    }
    __syncthreads();
    const long n = p.rows * p.C;
    for (long e = (long)blockIdx.x * blockDim.x + tid; e < n; e += (long)gridDim.x * blockDim.x) {
        const int c = (int)(e % p.C);
        p.dx[e] = s_sc[c] * (__ldg(p.dy + e) - s_m1[c] - __ldg(p.xhat + e) * s_m2[c]);
    }
}

// ---- soft Restorer (RestorerLinear with soft=True, models.py:634-655) --------------------------------------------------
// ml (B, 2) = linear_layer2 output: mu = ml[:, 0], logvar = ml[:, 1];  std = exp(logvar / 2);  noise (B, 1) ~ N(0, 1).
// The reference computes ``sampled_z * std + mu`` with sampled_z of shape (B, 1) and std / mu of shape (B,): torch broadcasts
// that to (B, B):  z[i][j] = noise[i] * std[j] + mu[j].  Reproduced as is (the drop-in contract is the reference's behaviour).
static __global__ void __launch_bounds__(256) iins_soft_reparam_kernel(const float* __restrict__ ml, const float* __restrict__ noise,
                                                                      float* __restrict__ z, int B) {
    iins_pdl_enter();
    const long n = (long)B * B;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        const int i = (int)(e / B), j = (int)(e - (long)i * B);
        z[e] = fmaf(__ldg(noise + i), expf(0.5f * __ldg(ml + 2 * j + 1)), __ldg(ml + 2 * j));
    }
}
// d_ml[j][0] = sum_i dz[i][j];   d_ml[j][1] = sum_i dz[i][j] * noise[i] * 0.5 * std[j].   One warp per column j.
static __global__ void __launch_bounds__(256) iins_soft_reparam_bwd_kernel(const float* __restrict__ ml, const float* __restrict__ noise,
                                                                          const float* __restrict__ dz, float* __restrict__ d_ml, int B) {
    iins_pdl_enter();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = blockIdx.x * 8 + warp; j < B; j += gridDim.x * 8) {
        float s0 = 0.f, s1 = 0.f;
        for (int i = lane; i < B; i += 32) {
            const float g = __ldg(dz + (long)i * B + j);
            s0 += g;
            s1 = fmaf(g, __ldg(noise + i), s1);
        }
        s0 = iins_warp_sum(s0); s1 = iins_warp_sum(s1);
        if (lane == 0) {
            d_ml[2 * j] = s0;
            d_ml[2 * j + 1] = s1 * 0.5f * expf(0.5f * __ldg(ml + 2 * j + 1));
        }
    }
}
