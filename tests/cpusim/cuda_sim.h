// cuda_sim.h -- TEST INFRASTRUCTURE ONLY.  A tiny single-process SIMT simulator that lets the plain
// CUDA C++ kernels under iins_vae_b200/csrc be compiled with g++ (-DIINS_CPUSIM) and executed on
// the CPU of the authoring container (which has no GPU), so kernel LOGIC (indexing, barriers,
// reductions, host sequencing) can be unit-tested before a GPU box is spent on it.
//
// It is NOT a fallback: the product package only ever loads the nvcc-built sm_100a library and
// raises if that is missing; this header is included solely by tests/cpusim/build.sh.
//
// Model: one CTA at a time; every CUDA thread is a ucontext fiber; __syncthreads / __syncwarp /
// warp shuffles are cooperative yield points with CUDA's "all live threads must arrive" rule
// (a deadlock -- i.e. a divergent barrier -- aborts with a message).  Kernels using inline PTX
// (tcgen05 / TMA) are excluded from the simulator build.
#pragma once
#include <ucontext.h>
#include <stdint.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { float4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static

typedef void* cudaStream_t;
typedef void* cudaEvent_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };

namespace cudasim {

enum State { RUNNABLE = 0, WAIT_CTA = 1, WAIT_WARP = 2, DONE = 3 };

struct Thread {
    ucontext_t ctx;
    dim3 tid;
    int lin;
    int state;
    char* stack;
};

constexpr int kMaxThreads = 1024;
constexpr size_t kStack = 256 * 1024;

struct World {
    Thread th[kMaxThreads];
    ucontext_t sched;
    Thread* cur = nullptr;
    dim3 blockIdx_, blockDim_, gridDim_;
    int nthreads = 0, ndone = 0, cta_wait = 0;
    int warp_wait[kMaxThreads / 32];
    int warp_live[kMaxThreads / 32];
    uint64_t warp_buf[kMaxThreads / 32][32];
    unsigned char* dyn_smem = nullptr;
    std::function<void()> body;
    bool stacks = false;
};

inline World& W() { static World w; return w; }

inline void trampoline() {
    World& w = W();
    w.body();
    w.cur->state = DONE;
    w.ndone++;
    w.warp_live[w.cur->lin / 32]--;
    swapcontext(&w.cur->ctx, &w.sched);
}

inline void release_checks() {
    World& w = W();
    int live = w.nthreads - w.ndone;
    if (live > 0 && w.cta_wait == live) {
        for (int i = 0; i < w.nthreads; ++i) if (w.th[i].state == WAIT_CTA) w.th[i].state = RUNNABLE;
        w.cta_wait = 0;
    }
    int nw = (w.nthreads + 31) / 32;
    for (int g = 0; g < nw; ++g) {
        if (w.warp_live[g] > 0 && w.warp_wait[g] == w.warp_live[g]) {
            for (int i = g * 32; i < g * 32 + 32 && i < w.nthreads; ++i)
                if (w.th[i].state == WAIT_WARP) w.th[i].state = RUNNABLE;
            w.warp_wait[g] = 0;
        }
    }
}

inline void yield_to_sched() {
    World& w = W();
    Thread* me = w.cur;
    swapcontext(&me->ctx, &w.sched);
}

inline void sync_cta() {
    World& w = W();
    w.cur->state = WAIT_CTA;
    w.cta_wait++;
    yield_to_sched();
}

inline void sync_warp() {
    World& w = W();
    w.cur->state = WAIT_WARP;
    w.warp_wait[w.cur->lin / 32]++;
    yield_to_sched();
}

template <typename T>
inline T shfl_idx(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    World& w = W();
    int warp = w.cur->lin / 32, lane = w.cur->lin % 32;
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    w.warp_buf[warp][lane] = bits;
    sync_warp();
    int nl = w.nthreads - warp * 32;
    if (nl > 32) nl = 32;
    uint64_t got = (src_lane >= 0 && src_lane < nl) ? w.warp_buf[warp][src_lane] : bits;
    sync_warp();
    T r;
    memcpy(&r, &got, sizeof(T));
    return r;
}

inline void launch(dim3 grid, dim3 block, size_t smem, std::function<void()> body) {
    World& w = W();
    int nt = (int)(block.x * block.y * block.z);
    if (nt > kMaxThreads) { fprintf(stderr, "cudasim: block too large\n"); abort(); }
    if (!w.stacks) {
        for (int i = 0; i < kMaxThreads; ++i) w.th[i].stack = (char*)malloc(kStack);
        w.stacks = true;
    }
    std::vector<unsigned char> dyn(smem + 1024);
    w.dyn_smem = (unsigned char*)(((uintptr_t)dyn.data() + 1023) & ~(uintptr_t)1023);
    w.body = body;
    w.blockDim_ = block;
    w.gridDim_ = grid;
    w.nthreads = nt;
    for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
        w.blockIdx_ = dim3(bx, by, bz);
        w.ndone = 0; w.cta_wait = 0;
        memset(w.warp_wait, 0, sizeof(w.warp_wait));
        memset(w.warp_live, 0, sizeof(w.warp_live));
        for (int i = 0; i < nt; ++i) {
            Thread& t = w.th[i];
            t.lin = i;
            t.tid = dim3(i % block.x, (i / block.x) % block.y, i / (block.x * block.y));
            t.state = RUNNABLE;
            w.warp_live[i / 32]++;
            getcontext(&t.ctx);
            t.ctx.uc_stack.ss_sp = t.stack;
            t.ctx.uc_stack.ss_size = kStack;
            t.ctx.uc_link = &w.sched;
            makecontext(&t.ctx, (void (*)())trampoline, 0);
        }
        while (w.ndone < nt) {
            bool progressed = false;
            for (int i = 0; i < nt; ++i) {
                if (w.th[i].state != RUNNABLE) continue;
                w.cur = &w.th[i];
                swapcontext(&w.sched, &w.th[i].ctx);
                progressed = true;
                release_checks();
            }
            if (!progressed) {
                release_checks();
                bool any = false;
                for (int i = 0; i < nt; ++i) any |= (w.th[i].state == RUNNABLE);
                if (!any) {
                    fprintf(stderr, "cudasim: DEADLOCK in block (%u,%u,%u): divergent barrier (cta_wait=%d live=%d)\n",
                            bx, by, bz, w.cta_wait, nt - w.ndone);
                    abort();
                }
            }
        }
    }
    w.cur = nullptr;
}

}  // namespace cudasim

#define threadIdx (cudasim::W().cur->tid)
#define blockIdx (cudasim::W().blockIdx_)
#define blockDim (cudasim::W().blockDim_)
#define gridDim (cudasim::W().gridDim_)

static inline void __syncthreads() { cudasim::sync_cta(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { cudasim::sync_warp(); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int lane_mask, int = 32) {
    return cudasim::shfl_idx(v, (cudasim::W().cur->lin % 32) ^ lane_mask);
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int delta, int = 32) {
    return cudasim::shfl_idx(v, (cudasim::W().cur->lin % 32) + delta);
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) {
    return cudasim::shfl_idx(v, src);
}
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
static inline double atomicAdd(double* p, double v) { double o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
static inline void __threadfence() {}
static inline long long clock64() { return 0; }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }

static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "cudasim"; }

#define IINS_LAUNCH(kernel, grid, block, smem, stream, ...) \
    cudasim::launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { kernel(__VA_ARGS__); })
#define IINS_DYN_SMEM(name) unsigned char* name = cudasim::W().dyn_smem
#define IINS_SET_FLOPS(f) ((void)(f))
#define IINS_SET_BYTES(b) ((void)(b))
#define IINS_SET_SHAPE(m, n, k) ((void)0)
