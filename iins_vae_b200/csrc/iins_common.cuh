// Common definitions for the IIns-VAE sm_100a kernels.
#pragma once

#ifdef IINS_CPUSIM
#include "cuda_sim.h"          // tests/cpusim: logic simulator, never part of the product build
#else
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <string.h>
#include <stdlib.h>
// Launch bookkeeping: a launch counter (bench.py reports it as gpu_launches) and an optional per-launch
// CUDA-event profile (iins_profile_begin / iins_profile_collect) used to time individual kernels inside
// a real step without an external profiler.
struct IinsProfState {
    unsigned long long launches;
    int enabled;
    int n;                               // recorded launches
    const char* names[4096];
    double flops[4096];
    double cur_flops;                    // set by the GEMM launchers just before IINS_LAUNCH
    double bytes[4096];                  // algorithmic HBM bytes of the launch (operands read once + results written once)
    double cur_bytes;
    int shape[4096][3];
    int cur_shape[3];
    cudaEvent_t ev0[4096], ev1[4096];
    int ev_ready;
};
// one instance for the whole library (defined in iins_runtime.cu; the kernel-instance translation units share it)
#ifdef IINS_DEFINE_GLOBALS
IinsProfState g_iins_prof;
int g_iins_pdl = -1;
thread_local int t_iins_pdl_force = -1;
#else
extern IinsProfState g_iins_prof;
extern int g_iins_pdl;
extern thread_local int t_iins_pdl_force;
#endif
static inline void iins_prof_pre(const char* name, cudaStream_t st) {
    g_iins_prof.launches++;
    if (g_iins_prof.enabled && g_iins_prof.n < 4096) {
        g_iins_prof.names[g_iins_prof.n] = name;
        g_iins_prof.flops[g_iins_prof.n] = g_iins_prof.cur_flops;
        g_iins_prof.bytes[g_iins_prof.n] = g_iins_prof.cur_bytes;
        for (int i = 0; i < 3; ++i) g_iins_prof.shape[g_iins_prof.n][i] = g_iins_prof.cur_shape[i];
        cudaEventRecord(g_iins_prof.ev0[g_iins_prof.n], st);
    }
}
static inline void iins_prof_post(cudaStream_t st) {
    if (g_iins_prof.enabled && g_iins_prof.n < 4096) {
        cudaEventRecord(g_iins_prof.ev1[g_iins_prof.n], st);
        g_iins_prof.n++;
    }
    g_iins_prof.cur_flops = 0.0;
    g_iins_prof.cur_bytes = 0.0;
    g_iins_prof.cur_shape[0] = g_iins_prof.cur_shape[1] = g_iins_prof.cur_shape[2] = 0;
}
// Every kernel of the library CAN be launched with programmatic dependent launch allowed (IINS_PDL=1): it then becomes
// resident while its predecessor in the stream is still draining, runs its prologue (index math, barrier / TMEM set-up,
// weight staging address arithmetic) and blocks in iins_pdl_wait() (griddepcontrol.wait = the predecessor grid has
// completed and its memory is visible) before it touches global memory.  In round 1 (a chain of ~124 dependent 10-30 us
// kernels on few streams) that hid the launch gap and the prologue of each (+3 %).  Since the persistent one-CTA-per-SM
// kernels and the wider stream concurrency of round 2 the early-resident grids cost more than they hide: measured
// (profiles/r02e_pdl_switches.log) the step is 1.3 % faster at B = 4096 and within +/- 1 % elsewhere with the attribute
// OFF, which is now the default.
static inline int iins_pdl_enabled() {
    // the environment is read ONCE (first launch of the process); a single-stream launch plan (the 2-D variant: ~330 small
    // dependent launches, no helper streams) turns the attribute on for its duration unless IINS_PDL was set explicitly: there
    // the overlapped prologues are a net gain (8.4 vs 9.1 ms per step)
    if (g_iins_pdl < 0) { const char* e = getenv("IINS_PDL"); g_iins_pdl = e ? (atoi(e) != 0 ? 3 : 2) : 0; }   // bit 1: explicit
    if (t_iins_pdl_force >= 0 && (g_iins_pdl & 2) == 0) return t_iins_pdl_force;
    return g_iins_pdl & 1;
}
#define IINS_LAUNCH(kernel, grid_, block_, smem_, stream_, ...)                                    \
    do {                                                                                            \
        iins_prof_pre(#kernel, (stream_));                                                          \
        cudaLaunchConfig_t cfg_;                                                                    \
        memset(&cfg_, 0, sizeof(cfg_));                                                             \
        cfg_.gridDim = dim3(grid_); cfg_.blockDim = dim3(block_);                                   \
        cfg_.dynamicSmemBytes = (size_t)(smem_); cfg_.stream = (stream_);                           \
        cudaLaunchAttribute attr_[1];                                                               \
        attr_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                           \
        attr_[0].val.programmaticStreamSerializationAllowed = iins_pdl_enabled();                   \
        cfg_.attrs = attr_; cfg_.numAttrs = 1;                                                      \
        cudaLaunchKernelEx(&cfg_, kernel, __VA_ARGS__);                                             \
        iins_prof_post((stream_));                                                                  \
    } while (0)
#define IINS_DYN_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
#define IINS_SET_FLOPS(f) (g_iins_prof.cur_flops = (f))
#define IINS_SET_BYTES(b) (g_iins_prof.cur_bytes = (b))
#define IINS_SET_SHAPE(m, n, k) (g_iins_prof.cur_shape[0] = (m), g_iins_prof.cur_shape[1] = (n), g_iins_prof.cur_shape[2] = (k))
#endif

#define IINS_HD __host__ __device__ __forceinline__
#define IINS_D __device__ __forceinline__

// Programmatic dependent launch, device side (see IINS_LAUNCH): let the next kernel in the stream start launching, then
// wait until everything before this kernel has completed.  Called at the top of EVERY kernel (before the first global
// access); the tensor-core kernels call iins_pdl_wait() after their prologue instead.
IINS_D void iins_pdl_launch_dependents() {
#ifndef IINS_CPUSIM
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
IINS_D void iins_pdl_wait() {
#ifndef IINS_CPUSIM
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
IINS_D void iins_pdl_enter() { iins_pdl_launch_dependents(); iins_pdl_wait(); }
// Persistent kernels (one CTA per SM for the kernel's whole life: the window and fused trunk kernels) do NOT trigger their
// dependents early: a dependent grid that becomes resident while they run only holds registers / shared memory / TMEM columns of
// the SMs it lands on (taken from the kernels of the OTHER streams) and spins in griddepcontrol.wait.  Measured on the B200
// (profiles/r02e_pdl_switches.log): the step is 3-5 % faster without the early trigger.  -DIINS_PERSISTENT_PDL=1 restores it.
#if defined(IINS_PERSISTENT_PDL) && IINS_PERSISTENT_PDL
#define IINS_PERSISTENT_PDL_TRIGGER() iins_pdl_launch_dependents()
#else
#define IINS_PERSISTENT_PDL_TRIGGER() ((void)0)
#endif

enum { IINS_PAD_ZERO = 0, IINS_PAD_REFLECT = 1, IINS_PAD_UP2 = 2 };
enum { IINS_ACT_NONE = 0, IINS_ACT_RELU = 1, IINS_ACT_LRELU = 2, IINS_ACT_TANH = 3 };
enum { IINS_NORM_NONE = 0, IINS_NORM_IN = 1, IINS_NORM_ADAIN = 2, IINS_NORM_LN = 3 };
enum { IINS_NLC = 0, IINS_NCL = 1 };

#define IINS_EPS 1e-5f          // InstanceNorm1d / AdaIN / custom LayerNorm eps (models.py:152,965,1049)

IINS_HD float iins_act(float v, int act, float slope) {
    if (act == IINS_ACT_RELU) return v > 0.f ? v : 0.f;
    if (act == IINS_ACT_LRELU) return v > 0.f ? v : v * slope;
    if (act == IINS_ACT_TANH) return tanhf(v);
    return v;
}

// derivative of the activation expressed through its OUTPUT y (sign-preserving activations)
IINS_HD float iins_dact_from_y(float y, int act, float slope) {
    if (act == IINS_ACT_RELU) return y > 0.f ? 1.f : 0.f;
    if (act == IINS_ACT_LRELU) return y > 0.f ? 1.f : slope;
    if (act == IINS_ACT_TANH) return 1.f - y * y;
    return 1.f;
}

IINS_D float iins_warp_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// Warp sums of N per-lane values (N = 4, 8, 16) with N - 1 + (5 - log2 N) shuffles instead of 5 N: at each of the first log2 N
// butterfly steps a lane keeps the half of the values its lane bit selects and sends the other half, so the number of values
// halves while the distance halves; the remaining steps are plain butterflies.  Every value is added in exactly the order
// iins_warp_sum uses (own + partner at distance 16, 8, 4, 2, 1), so the totals are bit-identical to N calls of it.
// Returns the total of value index iins_warp_sums_slot<N>(lane).
template <int N> IINS_D int iins_warp_sums_slot(int lane) {
    return N == 16 ? (lane >> 1) & 15 : N == 8 ? (lane >> 2) & 7 : (lane >> 3) & 3;
}
template <int N> IINS_D float iins_warp_sums(const float* v, int lane) {
    static_assert(N == 4 || N == 8 || N == 16, "N");
    float r[N];
#pragma unroll
    for (int j = 0; j < N; ++j) r[j] = v[j];
    int d = 16;
#pragma unroll
    for (int n = N; n > 1; n >>= 1, d >>= 1) {
        const bool hi = (lane & d) != 0;
#pragma unroll
        for (int j = 0; j < n / 2; ++j) {
            const float keep = hi ? r[j + n / 2] : r[j];
            const float send = hi ? r[j] : r[j + n / 2];
            r[j] = keep + __shfl_xor_sync(0xffffffffu, send, d);
        }
    }
#pragma unroll
    for (; d >= 1; d >>= 1) r[0] += __shfl_xor_sync(0xffffffffu, r[0], d);
    return r[0];
}

// sum over aligned groups of G consecutive lanes (G power of two <= 32)
IINS_D float iins_group_sum(float v, int G) {
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Geometry of one conv1d / linear layer.  Activations are channels-last (NLC) in HBM unless a
// layout flag says NCL (the reference's layout, used only at the module boundary).
struct IinsGeom {
    int B;                 // samples
    int Lin, Lout;         // positions per sample, input / output (1 for a Linear layer)
    int Cin, Cout;
    int ks, stride, pad;   // kernel taps, stride, padding amount
    int mode;              // IINS_PAD_ZERO | IINS_PAD_REFLECT | IINS_PAD_UP2 (nearest x2 then zero pad)
    int in_layout;         // layout of the layer input  (IINS_NLC / IINS_NCL)
    int out_layout;        // layout of the layer output
};

// How a gradient w.r.t. the conv OUTPUT (pre-norm pre-activation "dz") is read.
//   dz[b,l,co] = dy[...] * dy_scale * act'(y[b,l,co])
// dy is indexed like the layer output unless dy_bcast (then dy is (B,Cout), broadcast over l:
// the backward of AdaptiveAvgPool1d(1), models.py:279).
struct IinsDz {
    const float* dy;
    const float* y;        // layer output (post activation) for act', or nullptr
    int act;
    float slope;
    int dy_bcast;
    float dy_scale;
};

// source position inside the (un-padded, un-upsampled) input for output row l, tap t; -1 = zero
IINS_HD int iins_src_pos(const IinsGeom& g, int l, int t) {
    int u = l * g.stride + t - g.pad;
    if (g.mode == IINS_PAD_REFLECT) {
        if (u < 0) u = -u;
        else if (u >= g.Lin) u = 2 * (g.Lin - 1) - u;
        return u;
    }
    if (g.mode == IINS_PAD_UP2) {
        if (u < 0 || u >= 2 * g.Lin) return -1;
        return u >> 1;
    }
    if (u < 0 || u >= g.Lin) return -1;
    return u;
}

IINS_HD long iins_in_index(const IinsGeom& g, int b, int pos, int ci) {
    return g.in_layout == IINS_NCL ? ((long)b * g.Cin + ci) * g.Lin + pos : ((long)b * g.Lin + pos) * g.Cin + ci;
}
IINS_HD long iins_out_index(const IinsGeom& g, int b, int l, int co) {
    return g.out_layout == IINS_NCL ? ((long)b * g.Cout + co) * g.Lout + l : ((long)b * g.Lout + l) * g.Cout + co;
}

// forward implicit-GEMM A operand: row = (b,l), k = t*Cin + ci
IINS_D float iins_a_fwd(const IinsGeom& g, const float* __restrict__ x, int b, int l, int t, int ci) {
    int pos = iins_src_pos(g, l, t);
    if (pos < 0) return 0.f;
    return __ldg(x + iins_in_index(g, b, pos, ci));
}

IINS_D float iins_dz_at(const IinsGeom& g, const IinsDz& d, int b, int l, int co) {
    long idx = iins_out_index(g, b, l, co);
    float v = d.dy_bcast ? __ldg(d.dy + (long)b * g.Cout + co) : __ldg(d.dy + idx);
    v *= d.dy_scale;
    if (d.y != nullptr && d.act != IINS_ACT_NONE) v *= iins_dact_from_y(__ldg(d.y + idx), d.act, d.slope);
    return v;
}

// data-gradient implicit-GEMM A operand: row = (b,pos) of the layer INPUT, k = t*Cout + co.
// Gathers every output row l whose tap t reads input position pos (<= 3 candidates: the direct
// one plus the two reflected images, or the two upsampled copies).
IINS_D float iins_a_dgrad(const IinsGeom& g, const IinsDz& d, int b, int pos, int t, int co) {
    int q[3];
    int nq;
    if (g.mode == IINS_PAD_REFLECT) {
        q[0] = pos + g.pad;
        nq = 1;
        if (pos >= 1 && pos <= g.pad) q[nq++] = g.pad - pos;
        if (pos <= g.Lin - 2 && pos >= g.Lin - 1 - g.pad) q[nq++] = g.pad + 2 * (g.Lin - 1) - pos;
    } else if (g.mode == IINS_PAD_UP2) {
        q[0] = 2 * pos + g.pad;
        q[1] = 2 * pos + 1 + g.pad;
        nq = 2;
    } else {
        q[0] = pos + g.pad;
        nq = 1;
    }
    float acc = 0.f;
    for (int j = 0; j < nq; ++j) {
        int r = q[j] - t;
        if (r < 0) continue;
        int l = r / g.stride;
        if (l * g.stride != r || l >= g.Lout) continue;
        acc += iins_dz_at(g, d, b, l, co);
    }
    return acc;
}

// conv weight W[co][ci][t]  (torch Conv1d layout; Linear is ks == 1)
IINS_HD long iins_w_index(const IinsGeom& g, int co, int ci, int t) {
    return ((long)co * g.Cin + ci) * g.ks + t;
}
