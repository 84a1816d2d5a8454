// Micro-benchmark: issue-to-completion cost of tcgen05.mma (M=128, K=16, bf16) as a function of N, the number
// of independent TMEM accumulators the MMAs rotate over, and the smem operand footprint.  One CTA.
#include <cstdio>
#include <cstdlib>
#include "../iins_vae_b200/csrc/iins_umma.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void __launch_bounds__(128) bench_kernel(long long* out, int N, int nacc, int nmma, int distinct_a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 49152 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // bf16 ~0.0078
    if (tid == 0) { umma::mbar_init(umma::smem_u32(&mbar), 1); umma::fence_mbar_init(); }
    if (warp == 0) umma::tmem_alloc(umma::smem_u32(&tmem_slot), 512);
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        const uint32_t idesc = umma::make_idesc_bf16(128, N, 0, 0);
        const uint32_t a0 = umma::smem_u32(smem), b0 = umma::smem_u32(smem + 32768);
        const bool leader = umma::elect_one();
        long long t0 = clock64();
        const uint64_t ad0 = umma::make_desc(a0, 2048, 128), bd0 = umma::make_desc(b0, N * 16, 128);
        const uint32_t amask = distinct_a ? 7u : 0u, cmask = (uint32_t)(nacc - 1);
#pragma unroll 8
        for (int i = 0; i < nmma; ++i) {
            const uint64_t ad = ad0 + (uint64_t)(((uint32_t)i & amask) * 256u);          // +4096 B per step, >>4
            const uint64_t bd = bd0 + (uint64_t)(((uint32_t)i & 3u) * 128u);
            if (leader) umma::mma_bf16_ss(tmem + ((uint32_t)i & cmask) * (uint32_t)N, ad, bd, idesc, i >= nacc ? 1u : 0u);
        }
        long long t1 = clock64();
        if (leader) umma::commit(umma::smem_u32(&mbar));
        __syncwarp();
        umma::mbar_wait(umma::smem_u32(&mbar), 0);
        long long t2 = clock64();
        if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

int main() {
    long long* d; CK(cudaMalloc(&d, 16));
    CK(cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024));
    int Ns[] = {16, 64, 128, 256};
    for (int N : Ns) for (int nacc : {1, 2, 4}) for (int da : {0, 1}) {
        if (nacc * N > 512) continue;
        long long h[2];
        for (int rep = 0; rep < 2; ++rep) {
            bench_kernel<<<1, 128, 49152 + 1024>>>(d, N, nacc, 64, da);
            CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
        printf("N=%3d accumulators=%d distinctA=%d : issue %6.1f cyc/MMA, complete %6.1f cyc/MMA  (ideal %4.1f)\n", N, nacc, da, h[0] / 64.0, h[1] / 64.0, 128.0 * N / 256.0);
    }
    return 0;
}
