// Standalone bring-up test of the tcgen05 building blocks (descriptor layouts, TMEM, mbarrier) used by
// the tensor-core kernels.   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_test umma_test.cu
//   test 1: K-major A (128xK) x K-major B (NxK)^T, single bf16 piece        (conv/linear forward + dgrad)
//   test 2: same with the 3-piece bf16 split (6 MMAs per k-step)            -> fp32-grade accuracy
//   test 3: MN-major A (Rx128) and MN-major B (RxN): D = A^T B               (weight gradient)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../iins_vae_b200/csrc/iins_umma.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

// mode 0: K-major operands.  A [128][K] row-major, B [N][K] row-major.  pieces = 1 or 3.
// mode 1: MN-major operands. A [K][128] row-major (reduction over rows), B [K][N] row-major.
__global__ void __launch_bounds__(256) umma_test_kernel(const float* A, const float* B, float* D, int N, int K, int mode, int pieces, int swap) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int nchunk = K / 8;                      // 16-byte chunks along the reduction dim (K-major) ...
    // region sizes per piece
    const uint32_t a_bytes = 128 * K * 2, b_bytes = N * K * 2;
    unsigned char* sa = smem;                       // pieces x a_bytes
    unsigned char* sb = smem + 3 * a_bytes;
    if (tid == 0) { umma::mbar_init(umma::smem_u32(&mbar), 1); umma::fence_mbar_init(); }
    if (warp == 0) umma::tmem_alloc(umma::smem_u32(&tmem_slot), 64);
    // ---- fill operands
    if (mode == 0) {
        // [piece][kchunk][row][8 elems]
        for (int e = tid; e < 128 * nchunk; e += 256) {
            int row = e % 128, j = e / 128;
            uint32_t w[3][4];
            for (int q = 0; q < 4; ++q) {
                uint32_t p0a, p1a, p2a, p0b, p1b, p2b;
                umma::split3(A[row * K + j * 8 + 2 * q], p0a, p1a, p2a);
                umma::split3(A[row * K + j * 8 + 2 * q + 1], p0b, p1b, p2b);
                w[0][q] = umma::pack2(p0a, p0b); w[1][q] = umma::pack2(p1a, p1b); w[2][q] = umma::pack2(p2a, p2b);
            }
            for (int p = 0; p < pieces; ++p)
                *reinterpret_cast<uint4*>(sa + p * a_bytes + (j * 128 + row) * 16) = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
        }
        for (int e = tid; e < N * nchunk; e += 256) {
            int row = e % N, j = e / N;
            uint32_t w[3][4];
            for (int q = 0; q < 4; ++q) {
                uint32_t p0a, p1a, p2a, p0b, p1b, p2b;
                umma::split3(B[row * K + j * 8 + 2 * q], p0a, p1a, p2a);
                umma::split3(B[row * K + j * 8 + 2 * q + 1], p0b, p1b, p2b);
                w[0][q] = umma::pack2(p0a, p0b); w[1][q] = umma::pack2(p1a, p1b); w[2][q] = umma::pack2(p2a, p2b);
            }
            for (int p = 0; p < pieces; ++p)
                *reinterpret_cast<uint4*>(sb + p * b_bytes + (j * N + row) * 16) = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
        }
    } else {
        // MN-major: [piece][mn group of 8][k row][8 mn elems]   (k row = reduction index)
        for (int e = tid; e < (128 / 8) * K; e += 256) {
            int r = e % K, g = e / K;
            uint32_t w[3][4];
            for (int q = 0; q < 4; ++q) {
                uint32_t p0a, p1a, p2a, p0b, p1b, p2b;
                umma::split3(A[r * 128 + g * 8 + 2 * q], p0a, p1a, p2a);
                umma::split3(A[r * 128 + g * 8 + 2 * q + 1], p0b, p1b, p2b);
                w[0][q] = umma::pack2(p0a, p0b); w[1][q] = umma::pack2(p1a, p1b); w[2][q] = umma::pack2(p2a, p2b);
            }
            for (int p = 0; p < pieces; ++p)
                *reinterpret_cast<uint4*>(sa + p * a_bytes + (g * K + r) * 16) = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
        }
        for (int e = tid; e < (N / 8) * K; e += 256) {
            int r = e % K, g = e / K;
            uint32_t w[3][4];
            for (int q = 0; q < 4; ++q) {
                uint32_t p0a, p1a, p2a, p0b, p1b, p2b;
                umma::split3(B[r * N + g * 8 + 2 * q], p0a, p1a, p2a);
                umma::split3(B[r * N + g * 8 + 2 * q + 1], p0b, p1b, p2b);
                w[0][q] = umma::pack2(p0a, p0b); w[1][q] = umma::pack2(p1a, p1b); w[2][q] = umma::pack2(p2a, p2b);
            }
            for (int p = 0; p < pieces; ++p)
                *reinterpret_cast<uint4*>(sb + p * b_bytes + (g * K + r) * 16) = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
        }
    }
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        uint32_t idesc = umma::make_idesc_bf16(128, N, mode, mode);
        uint32_t acc = 0;
        // piece pairs (i,j) with i+j <= 2 when pieces == 3
        for (int ks = 0; ks < K / 16; ++ks) {
            for (int pa = 0; pa < pieces; ++pa)
                for (int pb = 0; pb < pieces; ++pb) {
                    if (pa + pb > 2) continue;
                    uint64_t ad, bd;
                    if (mode == 0) {
                        uint32_t a_lbo = 128 * 16, a_sbo = 128, b_lbo = N * 16, b_sbo = 128;
                        if (swap) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
                        ad = umma::make_desc(umma::smem_u32(sa + pa * a_bytes + ks * 2 * 128 * 16), a_lbo, a_sbo);
                        bd = umma::make_desc(umma::smem_u32(sb + pb * b_bytes + ks * 2 * N * 16), b_lbo, b_sbo);
                    } else {
                        // one MMA covers 16 reduction rows = 2 groups of 8 k-rows (LBO = 128 B); MN groups K*16 B apart
                        uint32_t a_lbo = 128, a_sbo = K * 16, b_lbo = 128, b_sbo = K * 16;
                        if (swap) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
                        ad = umma::make_desc(umma::smem_u32(sa + pa * a_bytes + ks * 16 * 16), a_lbo, a_sbo);
                        bd = umma::make_desc(umma::smem_u32(sb + pb * b_bytes + ks * 16 * 16), b_lbo, b_sbo);
                    }
                    umma::mma_bf16_ss(tmem, ad, bd, idesc, acc);
                    acc = 1;
                }
        }
        umma::commit(umma::smem_u32(&mbar));
    }
    umma::mbar_wait(umma::smem_u32(&mbar), 0);
    umma::tc_fence_after();
    if (warp < 4) {
        int row = warp * 32 + (tid & 31);
        for (int c0 = 0; c0 < N; c0 += 16) {
            float v[16];
            umma::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
            for (int i = 0; i < 16; ++i) D[row * N + c0 + i] = v[i];
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 64);
}

static float bf16r(float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u; float r; memcpy(&r, &u, 4); return r; }

int run(int N, int K, int mode, int pieces, int swap) {
    std::vector<float> A(128 * K), B(N * K), D(128 * N), R(128 * N);
    srand(1 + N + K + mode);
    for (auto& v : A) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (auto& v : B) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) {
                float a = mode == 0 ? A[m * K + k] : A[k * 128 + m];
                float b = mode == 0 ? B[n * K + k] : B[k * N + n];
                if (pieces == 1) { a = bf16r(a); b = bf16r(b); }
                s += (double)a * b;
            }
            R[m * N + n] = (float)s;
        }
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, D.size() * 4));
    size_t smem = 3 * (128 * K * 2) + 3 * (N * K * 2) + 1024;
    CK(cudaFuncSetAttribute(umma_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_test_kernel<<<1, 256, smem>>>(dA, dB, dD, N, K, mode, pieces, swap);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (size_t i = 0; i < D.size(); ++i) { maxerr = fmax(maxerr, fabs((double)D[i] - R[i])); maxref = fmax(maxref, fabs((double)R[i])); }
    printf("N=%d K=%d mode=%d pieces=%d swap=%d : max|err| %.3e (max|ref| %.3e) %s\n", N, K, mode, pieces, swap, maxerr, maxref,
           maxerr < (pieces == 1 ? 1e-4 : 2e-6) * maxref + 1e-7 ? "OK" : "MISMATCH");
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return 0;
}

int main(int argc, char** argv) {
    int swap = argc > 1 ? atoi(argv[1]) : 0;
    run(64, 64, 0, 1, swap);
    run(64, 64, 0, 3, swap);
    run(32, 32, 0, 3, swap);
    run(16, 48, 0, 3, swap);
    run(64, 64, 1, 1, swap);
    run(64, 32, 1, 3, swap);
    run(16, 32, 1, 3, swap);
    return 0;
}
