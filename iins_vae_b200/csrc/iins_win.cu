// Persistent window kernels for the k4 / stride-2 / zero-pad-1 convolutions (see iins_win.h).
//
// Resident operand.  A 128-row GEMM tile is 16 groups of 8 consecutive output rows (8 <= Lout, a power of two, so a group
// never straddles a sample).  Per bf16 piece the operand is stored as   A[chunk of 8 channels][group (16)][R rows][16 B]:
// the R = 8 + halo rows of a group are 16 bytes apart, i.e. rows r .. r+7 of a group are one UMMA core matrix (K-major, no
// swizzle) for EVERY r -- a tap is the same descriptor with its start address advanced by r * 16 bytes; the stride between
// groups (SBO = R * 16) is uniform across sample boundaries because every group carries its own halo rows (the first / last
// row of a group is stored twice: once as row 8 / 9 of the previous group or row 0 of the next one; the halo rows at the
// edges of a sample are the zero padding of the convolution and are never written).
//
//   forward (S2F),  out[l] = W0 x[2l-1] + W1 x[2l] + W2 x[2l+1] + W3 x[2l+2]:   two parity planes with R = 9,
//       plane O': rows r = 0..8 of group g hold x[2(8g+r-1)+1]   -> tap 0 = shift 0, tap 2 = shift 1
//       plane E : rows r = 0..8 hold x[2(8g+r)]                  -> tap 1 = shift 0, tap 3 = shift 1
//   data gradient split by the parity of the input position (S2D; iins_tc.cuh AKIND 2),
//       dx[2j+par] = sum_u W[2u+1-par]^T dz[j-u+par]:   ONE plane of dz rows with R = 10, rows r = 0..9 hold dz[8g+r-1],
//       (u, par) -> shift 1-u+par; the two parity classes are two TMEM accumulators fed from the same resident tile.
//
// Every input element is loaded (coalesced: the tile's input rows are one contiguous range), split into bf16 pieces and
// stored ONCE (the per-layer kernels do that once per tap that reads it).  The layer's packed weights (the same packs the
// per-layer kernels use) are loaded once per CTA.  Roles (544 threads, 1 CTA / SM, persistent over tiles): warps 0-7
// epilogue (the register epilogues of iins_tc.cuh), warps 8-15 producers, warp 16 weights + MMA issue.  Two operand stages
// and two TMEM accumulators: the producers fill tile i+1 while the MMAs of tile i run and the epilogue of tile i-1 drains.
#include "iins_win.h"
#include <stdio.h>
#include <type_traits>

namespace {

constexpr int WIN_THREADS = 544;
constexpr uint32_t WIN_SMEM_MAX = 227 * 1024;

template <int WK>
struct WinGeo {
    static constexpr int R = WK == IINS_WIN_S2F ? 9 : 10;          // rows per group incl. halo
    static constexpr int NPL = WK == IINS_WIN_S2F ? 2 : 1;         // planes
    static constexpr int NACC = WK == IINS_WIN_S2D ? 2 : 1;        // accumulators (parity classes) / weight packs
    static constexpr uint32_t SBO = R * 16;                        // bytes between groups
    static constexpr uint32_t LBO = 16 * SBO + 16;                 // bytes between 8-channel chunks (+16: the 8-byte stores of the
                                                                   // chunks of one row fall into different banks)
};

__host__ __device__ constexpr uint32_t align128(uint32_t v) { return (v + 127u) & ~127u; }
__host__ __device__ constexpr int win_tmem_cols(int need) { return need <= 32 ? 32 : (need <= 64 ? 64 : (need <= 128 ? 128 : (need <= 256 ? 256 : 512))); }


// split 4 floats into bf16 pieces ONCE (registers), then store them to one or two places
template <int PIECES>
__device__ __forceinline__ void win_split4(const float4 v, uint2* w) {
    if (PIECES == 3) {
        uint32_t a0, a1, a2, b0, b1, b2;
        umma::split3_pair(v.x, v.y, a0, a1, a2);
        umma::split3_pair(v.z, v.w, b0, b1, b2);
        w[0] = make_uint2(a0, b0); w[1] = make_uint2(a1, b1); w[2] = make_uint2(a2, b2);
    } else {
        w[0] = make_uint2(umma::cvt_bf16x2(v.x, v.y), umma::cvt_bf16x2(v.z, v.w));
    }
}
template <int PIECES>
__device__ __forceinline__ void win_store(const uint2* w, unsigned char* d, uint32_t piece_stride) {
#pragma unroll
    for (int pc = 0; pc < PIECES; ++pc) *reinterpret_cast<uint2*>(d + pc * piece_stride) = w[pc];
}
template <int NT, int PIECES, int WK, int EPI, int LL>
__global__ void __launch_bounds__(WIN_THREADS, 1) iins_win_nt_kernel(const IinsWinParams wp) {
    using G = WinGeo<WK>;
    constexpr int NACC = G::NACC;
    constexpr uint32_t ACC_COLS = PIECES * NT * NACC;
    constexpr int TCOLS = win_tmem_cols(2 * ACC_COLS);
    constexpr uint32_t B_TILE = 4 * PIECES * NT * 16;              // one (k block of 32) weight tile: [chunk][piece * NT + n][16 B]
    extern __shared__ __align__(1024) unsigned char dsm[];
    __shared__ __align__(8) unsigned long long a_full[2], a_empty[2], acc_full[2], acc_empty[2], w_ready;
    __shared__ uint32_t tmem_slot;
    __shared__ float xch[512];
    const IinsNTParams& p = wp.nt;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nch = wp.ca >> 3;                                    // 8-channel chunks of the resident operand
    const uint32_t PS = nch * G::LBO;                              // bytes between pieces
    const uint32_t PLS = PIECES * PS;                              // bytes between planes
    const uint32_t STAGE = align128(G::NPL * PLS);
    const uint32_t wbytes = NACC * wp.nkb * B_TILE;
    unsigned char* sW = dsm;
    unsigned char* sA = dsm + align128(wbytes);
    const int ntiles = (p.M + 127) >> 7;

    IINS_PERSISTENT_PDL_TRIGGER();
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            umma::mbar_init(umma::smem_u32(&a_full[i]), 256);
            umma::mbar_init(umma::smem_u32(&a_empty[i]), 1);
            umma::mbar_init(umma::smem_u32(&acc_full[i]), 1);
            umma::mbar_init(umma::smem_u32(&acc_empty[i]), 256);
        }
        umma::mbar_init(umma::smem_u32(&w_ready), 1);
        umma::fence_mbar_init();
    }
    if (warp == 16) umma::tmem_alloc(umma::smem_u32(&tmem_slot), TCOLS);
    // the halo rows at the edges of a sample are never written: they must read as the convolution's zero padding
    for (uint32_t i = (uint32_t)tid * 16u; i < 2 * STAGE; i += WIN_THREADS * 16u) *reinterpret_cast<uint4*>(sA + i) = make_uint4(0u, 0u, 0u, 0u);
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    iins_pdl_wait();                       // everything above overlapped the tail of the previous kernel in the stream

    if (warp == 16) {
        // ------------------------------------------------------------------ weights (once), then MMA issue (warp-uniform code)
        if (umma::elect_one()) {
            const uint32_t bar = umma::smem_u32(&w_ready);
            umma::mbar_arrive_expect_tx(bar, wbytes);
            for (int kb = 0; kb < wp.nkb; ++kb)
                umma::tma_bulk_g2s(umma::smem_u32(sW + kb * B_TILE), reinterpret_cast<const unsigned char*>(wp.wpack) + (size_t)kb * B_TILE, B_TILE, bar);
            if (NACC == 2)
                for (int kb = 0; kb < wp.nkb; ++kb)
                    umma::tma_bulk_g2s(umma::smem_u32(sW + (wp.nkb + kb) * B_TILE), reinterpret_cast<const unsigned char*>(wp.wpack_odd) + (size_t)kb * B_TILE, B_TILE, bar);
        }
        __syncwarp();
        umma::mbar_wait(umma::smem_u32(&w_ready), 0);
        const int ksteps = wp.ca >> 4;                             // k-steps of 16 channels per tap
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int s = it & 1;
            const uint32_t ph = (uint32_t)(it >> 1) & 1u;
            umma::mbar_wait(umma::smem_u32(&a_full[s]), ph);
            if (it >= 2) umma::mbar_wait(umma::smem_u32(&acc_empty[s]), ph ^ 1u);
            umma::tc_fence_after();
            const bool leader = umma::elect_one();
            const uint32_t a0 = umma::smem_u32(sA + s * STAGE);
            const uint32_t w0 = umma::smem_u32(sW);
            const uint32_t acc = tmem + (uint32_t)s * ACC_COLS;
            if (WK == IINS_WIN_S2F) {
                for (int t = 0; t < 4; ++t) {
                    const uint32_t ap = a0 + (uint32_t)(t & 1) * PLS + (uint32_t)(t >> 1) * 16u;
                    for (int kk = 0; kk < ksteps; ++kk) {
                        const int k0 = t * wp.ca + kk * 16;
                        const uint64_t ad = umma::make_desc(ap + (uint32_t)kk * 2u * G::LBO, G::LBO, G::SBO);
                        const uint64_t bd = umma::make_desc(w0 + (uint32_t)(k0 >> 5) * B_TILE + (uint32_t)((k0 >> 4) & 1) * (2u * PIECES * NT * 16u),
                                                            PIECES * NT * 16, 128);
                        iins_issue_kstep<NT, PIECES, 0, 0>(acc, ad, bd, PS >> 4, leader, (t | kk) ? 1u : 0u);
                    }
                }
            } else {
                for (int par = 0; par < 2; ++par) {
                    for (int u = 0; u < 2; ++u) {
                        const uint32_t ap = a0 + (uint32_t)(1 - u + par) * 16u;
                        for (int kk = 0; kk < ksteps; ++kk) {
                            const int k0 = u * wp.ca + kk * 16;
                            const uint64_t ad = umma::make_desc(ap + (uint32_t)kk * 2u * G::LBO, G::LBO, G::SBO);
                            const uint64_t bd = umma::make_desc(w0 + (uint32_t)(par * wp.nkb + (k0 >> 5)) * B_TILE +
                                                                    (uint32_t)((k0 >> 4) & 1) * (2u * PIECES * NT * 16u), PIECES * NT * 16, 128);
                            iins_issue_kstep<NT, PIECES, 0, 0>(acc + (uint32_t)par * (PIECES * NT), ad, bd, PS >> 4, leader, (u | kk) ? 1u : 0u);
                        }
                    }
                }
            }
            if (leader) {
                umma::commit(umma::smem_u32(&a_empty[s]));         // the stage may be refilled once these MMAs have read it
                umma::commit(umma::smem_u32(&acc_full[s]));
            }
            __syncwarp();
        }
    } else if (warp >= 8) {
        // ------------------------------------------------------------------ producers
        const int pt = tid - 256;
        const int c4n = wp.ca >> 2;                                // float4 units per operand row
        int c4sh = 0;
        while ((1 << c4sh) < c4n) ++c4sh;
        const int c4 = pt & (c4n - 1), r0 = pt >> c4sh, rstep = 256 >> c4sh;
        const uint32_t coff = (uint32_t)(c4 >> 1) * G::LBO + (uint32_t)(c4 & 1) * 8u;
        const int Lr = 1 << wp.lsh_in;                             // rows per sample of the resident operand
        constexpr int TROWS = WK == IINS_WIN_S2F ? 256 : 128;      // operand rows per tile
        const int nun = TROWS / rstep;                             // units per thread and tile: 4 or 8 (a power of two <= 8)
        int nsh = 0, cash = 0;
        while ((1 << nsh) < nun) ++nsh;
        while ((1 << cash) < wp.ca) ++cash;
        const int my_tiles = blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
        const int total = my_tiles << nsh;                         // this thread's units over all of the CTA's tiles
        // Unit q = (tile number q >> nsh of this CTA, unit q & (nun - 1) of that tile).  Everything about a unit except its tile is
        // a per-thread constant: element offset inside the tile, destination inside the stage, the offset of its second copy
        // (0 = none; a row is the halo row of at most one neighbour group).  Eight units = one unrolled round (nun = 4: two tiles).
        int goff[8], dup[8];
        uint32_t doff[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int row = r0 + (u & (nun - 1)) * rstep;
            const int pos = row & (Lr - 1);                        // position inside the sample
            goff[u] = row * wp.ca + c4 * 4;
            if (WK == IINS_WIN_S2F) {
                // odd position  -> plane O' (offset 0):   O[j] is row i8 + 1 of its group and row 0 of the next one
                // even position -> plane E (offset PLS):  E[j] is row i8 of its group and row 8 of the previous one
                const int orow = row >> 1, g = orow >> 3, i8 = orow & 7, j = pos >> 1, odd = pos & 1;
                doff[u] = (odd ? 0u : PLS) + coff + (uint32_t)g * G::SBO + (uint32_t)(i8 + odd) * 16u;
                dup[u] = odd ? ((i8 == 7 && j < (Lr >> 1) - 1) ? (int)G::SBO - 128 : 0) : ((i8 == 0 && j > 0) ? 128 - (int)G::SBO : 0);
            } else {
                // dz[l] is row i8 + 1 of its group, row 9 of the previous one (i8 == 0) or row 0 of the next one (i8 == 7)
                const int g = row >> 3, i8 = row & 7;
                doff[u] = coff + (uint32_t)g * G::SBO + (uint32_t)(i8 + 1) * 16u;
                dup[u] = (i8 == 0 && pos > 0) ? 128 - (int)G::SBO : ((i8 == 7 && pos < Lr - 1) ? (int)G::SBO - 128 : 0);
            }
        }
        const long lime = ((long)p.g.B << wp.lsh_in) << cash;      // operand elements that exist
        const bool has_y = WK == IINS_WIN_S2D && p.dz.y != nullptr && p.dz.act != IINS_ACT_NONE;
        const bool bcast = WK == IINS_WIN_S2D && p.dz.dy_bcast != 0;
        const float* src = WK == IINS_WIN_S2F ? p.x : p.dz.dy;
        // RAW operand data rides in a register ring (8 loads in flight per thread: 8 units, or 4 units of a gradient with an
        // activation mask = dy and y); nothing touches a loaded value before its unit is consumed, so the global-load latency of
        // a tile overlaps the split / store work of the units before it
        auto run = [&](auto tag) {
            constexpr bool HAS_Y = decltype(tag)::value;
            constexpr int D = HAS_Y ? 4 : 8;
            float4 vd[D], vy[HAS_Y ? D : 1];
            auto issue = [&](int slot, int u, int q) {
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f), y = a;
                if (q < total) {
                    const int tile = (int)blockIdx.x + (q >> nsh) * (int)gridDim.x;
                    const long ge = (((long)tile * TROWS) << cash) + goff[u];
                    if (ge < lime) {
                        a = __ldg(reinterpret_cast<const float4*>(src + (bcast ? ((ge >> (wp.lsh_in + cash)) << cash) + c4 * 4 : ge)));
                        if (HAS_Y) y = __ldg(reinterpret_cast<const float4*>(p.dz.y + ge));
                    }
                }
                vd[slot] = a;
                if (HAS_Y) vy[slot] = y;
            };
            auto consume = [&](int slot, int u, int q) {
                const int it = q >> nsh, ii = q & (nun - 1), s = it & 1;
                if (ii == 0 && it >= 2) umma::mbar_wait_suspend(umma::smem_u32(&a_empty[s]), ((uint32_t)(it >> 1) & 1u) ^ 1u);
                float4 v = vd[slot];
                if (WK == IINS_WIN_S2D) {
                    if (HAS_Y) {
                        const float4 y = vy[slot];
                        v.x *= iins_dact_from_y(y.x, p.dz.act, p.dz.slope); v.y *= iins_dact_from_y(y.y, p.dz.act, p.dz.slope);
                        v.z *= iins_dact_from_y(y.z, p.dz.act, p.dz.slope); v.w *= iins_dact_from_y(y.w, p.dz.act, p.dz.slope);
                    }
                    if (p.dz.dy_scale != 1.f) { v.x *= p.dz.dy_scale; v.y *= p.dz.dy_scale; v.z *= p.dz.dy_scale; v.w *= p.dz.dy_scale; }
                }
                uint2 w[3];
                win_split4<PIECES>(v, w);
                unsigned char* d = sA + s * STAGE + doff[u];
                win_store<PIECES>(w, d, PS);
                if (dup[u] != 0) win_store<PIECES>(w, d + dup[u], PS);
                if (ii == nun - 1) {
                    umma::fence_async_smem();
                    umma::mbar_arrive(umma::smem_u32(&a_full[s]));
                }
            };
#pragma unroll
            for (int u = 0; u < D; ++u) issue(u, u, u);
            for (int q0 = 0; q0 < total; q0 += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (q0 + u >= total) break;
                    consume(u % D, u, q0 + u);
                    issue(u % D, (u + D) & 7, q0 + u + D);
                }
            }
        };
        if (has_y) run(std::true_type{}); else run(std::false_type{});
    } else {
        // ------------------------------------------------------------------ epilogue warps
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int s = it & 1;
            umma::mbar_wait_suspend(umma::smem_u32(&acc_full[s]), (uint32_t)(it >> 1) & 1u);
            umma::tc_fence_after();
            const uint32_t acc = tmem + (uint32_t)s * ACC_COLS;
            if (WK == IINS_WIN_S2F) {
                iins_tc_epilogue_regs<NT, PIECES, EPI, LL>(p, acc, tile * 128, 0, warp, lane, xch);
            } else {
                // GEMM row (b, j) of parity class `par` is input position 2 j + par
                iins_tc_epilogue_regs<NT, PIECES, EPI, LL>(p, acc, tile * 128, 0, warp, lane, xch, 2, 0);
                iins_tc_epilogue_regs<NT, PIECES, EPI, LL>(p, acc + PIECES * NT, tile * 128, 0, warp, lane, xch, 2, 1);
            }
            umma::tc_fence_before();
            umma::mbar_arrive(umma::smem_u32(&acc_empty[s]));
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 16) umma::tmem_dealloc(tmem, TCOLS);
}

uint32_t win_smem_bytes(int nt, int pieces, int wk, int ca) {
    const uint32_t lbo = wk == IINS_WIN_S2F ? WinGeo<IINS_WIN_S2F>::LBO : WinGeo<IINS_WIN_S2D>::LBO;
    const int npl = wk == IINS_WIN_S2F ? 2 : 1, nacc = wk == IINS_WIN_S2D ? 2 : 1;
    const int K = (wk == IINS_WIN_S2F ? 4 : 2) * ca, nkb = (K + 31) / 32;
    const uint32_t stage = align128((uint32_t)npl * pieces * (ca >> 3) * lbo);
    const uint32_t wbytes = (uint32_t)nacc * nkb * 4u * pieces * nt * 16u;
    return align128(wbytes) + 2 * stage;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
    }
    return n;
}

// ======================================================================================================= weight gradient
// dW[co][ci][t] = sum_(b,l) dz[b,l,co] x[b, 2l+t-1, ci] of the same convolutions: the x planes of the forward kernel and the
// dz rows of the tile are resident as bf16 pieces, and BOTH are read MN-major by the MMA (the reduction runs over the rows:
// a K group of the instruction is one 8-row group of the layouts above, 16 rows per instruction).  A = dz with the pieces
// STACKED ALONG M (the pieces are consecutive blocks of 8-channel chunks with one chunk stride, so M = 128 lanes cover
// [dz0; dz1; dz2; -] for 32 output channels and [dz0; dz1] for 64), B = the x plane at the tap's row shift with its pieces
// stacked along N: one instruction forms all piece products of a (tap, 16 rows) pair (64 output channels: a second one for
// dz2 x [x0 | x1]).  The accumulators stay in TMEM for the CTA's whole life (persistent over tiles); at the end the piece
// blocks are summed through shared memory and flushed with one atomic per weight.  NROLE = 2 (64 output channels: the
// accumulators of four taps do not fit 512 TMEM columns): blockIdx.y selects ONE parity plane = two taps.
struct IinsWinTNParams {
    IinsTNParams tn;             // geometry, x, dz, dw, db; M = B * Lout
    int lsh_out;                 // log2(Lout)
};

template <int CIN, int COUT, int PIECES>
struct WinTN {
    static constexpr int NROLE = (PIECES == 3 && COUT > 32) ? 2 : 1;
    static constexpr bool TWO = PIECES * COUT > 128;                       // second instruction for the third dz piece
    static constexpr int CPT = PIECES * CIN + (TWO ? 2 * CIN : 0);         // accumulator columns per tap
    static constexpr int NTAP = 4 / NROLE;
    static constexpr int TCOLS = win_tmem_cols(NTAP * CPT);
    static constexpr int NCHX = CIN / 8, NCHZ = COUT / 8;
    static constexpr uint32_t LBOX = WinGeo<IINS_WIN_S2F>::LBO, SBOX = WinGeo<IINS_WIN_S2F>::SBO;
    static constexpr uint32_t CSZ = 16 * 128 + 16;                         // dz: bytes between 8-channel chunks
    static constexpr uint32_t PSX = NCHX * LBOX, PLSX = PIECES * PSX;      // x: bytes between pieces / planes
    static constexpr uint32_t PSZ = NCHZ * CSZ;
    static constexpr uint32_t DZ_BYTES = PIECES * PSZ;
    static constexpr uint32_t X_BYTES = (2 / NROLE) * PLSX;
    // M = 128 lanes read 16 chunks from the A start address: the chunks beyond the last dz piece must lie inside the stage
    static constexpr uint32_t TAIL = 16 * CSZ;
    static constexpr uint32_t STAGE = align128((DZ_BYTES + X_BYTES) > ((TWO ? 2 * PSZ : 0) + TAIL) ? (DZ_BYTES + X_BYTES) : ((TWO ? 2 * PSZ : 0) + TAIL));
    static constexpr uint32_t SMEM = 2 * STAGE;
};

template <int CIN, int COUT, int PIECES>
__global__ void __launch_bounds__(WIN_THREADS, 1) iins_win_tn_kernel(const IinsWinTNParams wp) {
    using T = WinTN<CIN, COUT, PIECES>;
    constexpr int NROLE = T::NROLE, NTAP = T::NTAP, CPT = T::CPT;
    extern __shared__ __align__(1024) unsigned char dsm[];
    __shared__ __align__(8) unsigned long long a_full[2], a_empty[2], done;
    __shared__ uint32_t tmem_slot;
    __shared__ double sbias[COUT];
    const IinsTNParams& p = wp.tn;
    const IinsGeom& g = p.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = (p.M + 127) >> 7;
    const int role = NROLE == 2 ? (int)blockIdx.y : 0;             // NROLE 2: role 0 = plane O' (taps 0, 2), role 1 = plane E (taps 1, 3)
    const int my_tiles = (int)blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    IINS_PERSISTENT_PDL_TRIGGER();
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            umma::mbar_init(umma::smem_u32(&a_full[i]), 512);
            umma::mbar_init(umma::smem_u32(&a_empty[i]), 1);
        }
        umma::mbar_init(umma::smem_u32(&done), 1);
        umma::fence_mbar_init();
    }
    if (tid < COUT) sbias[tid] = 0.0;
    if (warp == 16) umma::tmem_alloc(umma::smem_u32(&tmem_slot), T::TCOLS);
    for (uint32_t i = (uint32_t)tid * 16u; i < T::SMEM; i += WIN_THREADS * 16u) *reinterpret_cast<uint4*>(dsm + i) = make_uint4(0u, 0u, 0u, 0u);
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    iins_pdl_wait();

    // bias gradient = column sums of dz: accumulated in DOUBLE per thread and per CTA (a thread's rows belong to many samples:
    // for a convolution in front of an InstanceNorm the true sum is 0 and only the whole-sample totals cancel)
    double bsum[4] = {0.0, 0.0, 0.0, 0.0};
    if (warp == 16) {
        // ------------------------------------------------------------------ MMA issue (warp-uniform code)
        for (int it = 0; it < my_tiles; ++it) {
            const int s = it & 1;
            umma::mbar_wait(umma::smem_u32(&a_full[s]), (uint32_t)(it >> 1) & 1u);
            umma::tc_fence_after();
            const bool leader = umma::elect_one();
            const uint32_t zb = umma::smem_u32(dsm + s * T::STAGE);
            const uint32_t xb = zb + T::DZ_BYTES;
            for (int ks = 0; ks < 8; ++ks) {
                const uint32_t acc_on = (it | ks) ? 1u : 0u;
                // MN-major: LBO = bytes between 8-row (K) groups, SBO = bytes between 8-channel chunks
                const uint64_t ad = umma::make_desc(zb + (uint32_t)ks * 256u, 128, T::CSZ);
                for (int tt = 0; tt < NTAP; ++tt) {
                    const int plane = NROLE == 2 ? 0 : (tt & 1), shift = NROLE == 2 ? tt : (tt >> 1);
                    const uint64_t bd = umma::make_desc(xb + (uint32_t)plane * T::PLSX + (uint32_t)shift * 16u + (uint32_t)ks * 2u * T::SBOX, T::SBOX, T::LBOX);
                    const uint32_t col = tmem + (uint32_t)(tt * CPT);
                    if (leader) umma::mma_bf16_ss(col, ad, bd, umma::make_idesc_bf16(128, PIECES * CIN, 1, 1), acc_on);
                    if (T::TWO) { if (leader) umma::mma_bf16_ss(col + PIECES * CIN, ad + ((2 * T::PSZ) >> 4), bd, umma::make_idesc_bf16(128, 2 * CIN, 1, 1), acc_on); }
                }
            }
            if (leader) {
                umma::commit(umma::smem_u32(&a_empty[s]));
                if (it == my_tiles - 1) umma::commit(umma::smem_u32(&done));
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ producers: ALL 16 other warps (this kernel has no per-tile
        // epilogue; the producers run at IPC ~1 on global-load and conversion latency, so twice the warps is close to twice the rate)
        const int pt = tid;
        constexpr int C4X = CIN / 4, RSX = 512 / C4X, NX = (NROLE == 1 ? 256 : 128) / RSX;
        constexpr int C4Z = COUT / 4, RSZ = 512 / C4Z, NZ = 128 / RSZ;
        constexpr int UPT = NX + NZ;
        const int c4x = pt % C4X, r0x = pt / C4X, c4z = pt % C4Z, r0z = pt / C4Z;
        const uint32_t coffx = (uint32_t)(c4x >> 1) * T::LBOX + (uint32_t)(c4x & 1) * 8u;
        const uint32_t coffz = (uint32_t)(c4z >> 1) * T::CSZ + (uint32_t)(c4z & 1) * 8u;
        const int Lout = 1 << wp.lsh_out;
        const bool do_bias = p.db != nullptr && role == 0;
        const long xlim = (long)g.B << (wp.lsh_out + 1), zlim = (long)g.B << wp.lsh_out;       // operand rows that exist
        const bool has_y = p.dz.y != nullptr && p.dz.act != IINS_ACT_NONE;
        const bool bcast = p.dz.dy_bcast != 0;
        // RAW operand data rides in a register ring (see the forward kernel); a tile is UPT units per thread: NX of x, then NZ of dz
        auto run = [&](auto tag) {
            constexpr bool HAS_Y = decltype(tag)::value;
            constexpr int D = HAS_Y ? (UPT % 4 == 0 ? 4 : 3) : UPT;   // ring depth: divides UPT (<= 8), so a unit's slot is the same in every tile
            static_assert(UPT <= 8, "units per tile");
            static_assert(UPT % D == 0, "ring depth must divide the units per tile");
            float4 vd[D], vy[HAS_Y ? D : 1];
            auto issue = [&](int slot, int u, int it) {            // unit u (compile-time after unrolling) of this CTA's tile number `it`
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f), y = a;
                if (it < my_tiles) {
                    const int tile = (int)blockIdx.x + it * (int)gridDim.x;
                    if (u < NX) {
                        const int row = r0x + u * RSX;
                        // input row of the tile: NROLE 1 -> row itself (both parities); NROLE 2 -> position 2 * orow + parity of this plane
                        const long grow = (long)tile * 256 + (NROLE == 1 ? row : 2 * row + (role == 0 ? 1 : 0));
                        if (grow < xlim) a = __ldg(reinterpret_cast<const float4*>(p.x + grow * CIN + c4x * 4));
                    } else {
                        const long grow = (long)tile * 128 + r0z + (u - NX) * RSZ;
                        if (grow < zlim) {
                            const long ge = grow * COUT + c4z * 4;
                            a = __ldg(reinterpret_cast<const float4*>(p.dz.dy + (bcast ? (grow >> wp.lsh_out) * COUT + c4z * 4 : ge)));
                            if (HAS_Y) y = __ldg(reinterpret_cast<const float4*>(p.dz.y + ge));
                        }
                    }
                }
                vd[slot] = a;
                if (HAS_Y) vy[slot] = y;
            };
            auto consume = [&](int slot, int u, int it) {
                const int s = it & 1;
                if (u == 0 && it >= 2) umma::mbar_wait_suspend(umma::smem_u32(&a_empty[s]), ((uint32_t)(it >> 1) & 1u) ^ 1u);
                unsigned char* zst = dsm + s * T::STAGE;
                float4 v = vd[slot];
                uint2 w[3];
                if (u < NX) {
                    // odd position  -> plane O' (offset 0):   O[j] is row i8 + 1 of its group and row 0 of the next one
                    // even position -> plane E (offset PLSX): E[j] is row i8 of its group and row 8 of the previous one
                    const int row = r0x + u * RSX;
                    const int orow = NROLE == 1 ? (row >> 1) : row;
                    const int odd = NROLE == 1 ? (row & 1) : (role == 0 ? 1 : 0);
                    const int gq = orow >> 3, i8 = orow & 7, j = orow & (Lout - 1);
                    win_split4<PIECES>(v, w);
                    unsigned char* d = zst + T::DZ_BYTES + ((NROLE == 1 && !odd) ? T::PLSX : 0u) + coffx + (uint32_t)gq * T::SBOX + (uint32_t)(i8 + odd) * 16u;
                    win_store<PIECES>(w, d, T::PSX);
                    const int dup = odd ? ((i8 == 7 && j < Lout - 1) ? (int)T::SBOX - 128 : 0) : ((i8 == 0 && j > 0) ? 128 - (int)T::SBOX : 0);
                    if (dup != 0) win_store<PIECES>(w, d + dup, T::PSX);
                } else {
                    if (HAS_Y) {
                        const float4 y = vy[slot];
                        v.x *= iins_dact_from_y(y.x, p.dz.act, p.dz.slope); v.y *= iins_dact_from_y(y.y, p.dz.act, p.dz.slope);
                        v.z *= iins_dact_from_y(y.z, p.dz.act, p.dz.slope); v.w *= iins_dact_from_y(y.w, p.dz.act, p.dz.slope);
                    }
                    if (p.dz.dy_scale != 1.f) { v.x *= p.dz.dy_scale; v.y *= p.dz.dy_scale; v.z *= p.dz.dy_scale; v.w *= p.dz.dy_scale; }
                    win_split4<PIECES>(v, w);
                    win_store<PIECES>(w, zst + coffz + (uint32_t)(r0z + (u - NX) * RSZ) * 16u, T::PSZ);
                    if (do_bias) { bsum[0] += (double)v.x; bsum[1] += (double)v.y; bsum[2] += (double)v.z; bsum[3] += (double)v.w; }
                }
                if (u == UPT - 1) {
                    umma::fence_async_smem();
                    umma::mbar_arrive(umma::smem_u32(&a_full[s]));
                }
            };
#pragma unroll
            for (int u = 0; u < D; ++u) issue(u, u, 0);
            for (int it = 0; it < my_tiles; ++it) {
#pragma unroll
                for (int u = 0; u < UPT; ++u) {
                    consume(u % D, u, it);
                    issue(u % D, (u + D) % UPT, it + (u + D >= UPT ? 1 : 0));
                }
            }
        };
        if (has_y) run(std::true_type{}); else run(std::false_type{});
    }
    // ---- flush: every MMA has completed -> the stages are free (the reduction buffer aliases stage 0), TMEM is final
    umma::mbar_wait_suspend(umma::smem_u32(&done), 0);
    umma::tc_fence_after();
    float* red = reinterpret_cast<float*>(dsm);                    // [tap of this role][ci][co]: the lanes of a warp (= co) hit different banks
    constexpr int NRED = COUT * NTAP * CIN;
    for (int e = tid; e < NRED; e += WIN_THREADS) red[e] = 0.f;
    __syncthreads();
    if (warp < 8) {
        // TMEM lane = piece_a * COUT + co; the two warps that share a lane quarter split the taps
        const int q = warp & 3, hf = warp >> 2;
        const int r = q * 32 + lane;
        const int pa = r / COUT, co = r % COUT;
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
        if (pa < PIECES || T::TWO) {                               // warp-uniform for COUT >= 32 (a piece block is a multiple of 32 lanes)
            for (int tt = hf; tt < NTAP; tt += 2) {
#pragma unroll
                for (int c0 = 0; c0 < CIN; c0 += 16) {
                    float acc16[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc16[i] = 0.f;
                    // first instruction: lanes [dz_pa] (pa < 2 when TWO, else pa < PIECES), columns [x0 | x1 | x2]
#pragma unroll
                    for (int pb = 0; pb < PIECES; ++pb) {
                        float u[16];
                        umma::tmem_ld16(tl + (uint32_t)(tt * CPT + pb * CIN + c0), u);
#pragma unroll
                        for (int i = 0; i < 16; ++i) acc16[i] += u[i];
                    }
                    if (T::TWO) {                                  // second instruction: lanes 0..COUT-1 hold dz2 x [x0 | x1]
                        float u[16], w[16];
                        umma::tmem_ld16(tl + (uint32_t)(tt * CPT + PIECES * CIN + c0), u);
                        umma::tmem_ld16(tl + (uint32_t)(tt * CPT + PIECES * CIN + CIN + c0), w);
                        if (pa == 0) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) acc16[i] += u[i] + w[i];
                        }
                    }
                    if (pa < (T::TWO ? 2 : PIECES)) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) atomicAdd(red + (tt * CIN + c0 + i) * COUT + co, acc16[i]);
                    }
                }
            }
        }
    }
    if (warp < 16 && p.db != nullptr && role == 0) {
        const int c4z = tid % (COUT / 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) atomicAdd(&sbias[c4z * 4 + k], bsum[k]);
    }
    umma::tc_fence_before();
    __syncthreads();
    for (int e = tid; e < NRED; e += WIN_THREADS) {
        const int co = e % COUT, ci = (e / COUT) % CIN, tt = e / (COUT * CIN);
        const int t = NROLE == 2 ? 2 * tt + role : tt;             // role 0 (plane O'): taps 0, 2;  role 1 (plane E): taps 1, 3
        atomicAdd(p.dw + ((long)co * CIN + ci) * 4 + t, red[e]);
    }
    if (p.db != nullptr && role == 0 && tid < COUT) atomicAdd(p.db + tid, (float)sbias[tid]);
    if (warp == 16) umma::tmem_dealloc(tmem, T::TCOLS);
}

template <int CIN, int COUT, int PIECES>
void launch_win_tn_v(cudaStream_t st, const IinsWinTNParams& p) {
    using T = WinTN<CIN, COUT, PIECES>;
    static bool attr = false;
    auto iins_win_tn_kernel_ = iins_win_tn_kernel<CIN, COUT, PIECES>;
    if (!attr) { cudaFuncSetAttribute(iins_win_tn_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM); attr = true; }
    const int ntiles = (p.tn.M + 127) / 128;
    int gx = sm_count() / T::NROLE;
    if (gx > ntiles) gx = ntiles;
    IINS_LAUNCH(iins_win_tn_kernel_, dim3(gx, T::NROLE, 1), WIN_THREADS, T::SMEM, st, p);
}

// ======================================================================================== trunk weight gradients (k3, 64 -> 64, L = 8)
// The weight gradients of the residual trunk's k3 reflect-pad convolutions (up to 8 convolutions of identical geometry in one
// launch: blockIdx.y), in the same resident / MN-major form.  A tile is 16 samples = 128 rows; x is stored as in the fused
// trunk kernels -- 10 rows per sample, the reflected edge rows materialised, tap t = row shift t -- and dz as 8-row groups.
// 64 output channels x 64 input channels x 3 taps with every piece product in its own accumulator would need 960 TMEM columns;
// instead the products are grouped by magnitude class (fp32-grade mode, pieces d0 > d1 > d2 of dz and x0 > x1 > x2 of x):
//     A = [d0; d1] (M-stacked)  x  x0  -> block 0:  lanes 0-63  d0 x0,            lanes 64-127  d1 x0
//     A = [d0; d1]              x  x1  -> block 1:  lanes 0-63  d0 x1,            lanes 64-127  d1 x1
//     A = [d0; d1]              x  x2  -> block 1:  lanes 0-63  + d0 x2,          lanes 64-127  + d1 x2
//     A = [d2; 0 ]              x  x0  -> block 1:  lanes 0-63  + d2 x0           (the zero block follows d2 in shared memory)
// i.e. 128 columns per tap, 384 in all; seven of the nine products (d2 x1 and d2 x2, <= 2^-24 of the result, are dropped).
struct IinsWinK3Params {
    int B, nconv, db_on;
    const float* x[8]; const float* dz[8]; float* dw[8]; float* db[8];
};

template <int PIECES>
struct WinK3 {
    static constexpr int TS = 8;                                           // samples per tile (64 rows): two stages fit shared memory
    static constexpr int NBLK = PIECES == 3 ? 4 : 1;                       // dz blocks of 8 chunks: [d0][d1][d2][zero] / [d]
    static constexpr uint32_t CSZ = TS * 128 + 16;                         // dz: bytes between 8-channel chunks
    static constexpr uint32_t PSZ = 8 * CSZ;
    static constexpr uint32_t XL = TS * 160 + 16;                          // x: bytes between 8-channel chunks (TS samples x 10 rows)
    static constexpr uint32_t PSX = 8 * XL;
    static constexpr uint32_t DZ_BYTES = (NBLK > 2 ? NBLK : 2) * PSZ;      // M = 128 lanes always read 16 chunks from the A start
    static constexpr uint32_t X_BYTES = PIECES * PSX;
    static constexpr uint32_t STAGE = align128(DZ_BYTES + X_BYTES);
    static constexpr uint32_t RED_BYTES = 3 * 64 * 64 * 4;                 // flush buffer (aliases the stages)
    static constexpr uint32_t SMEM = 2 * STAGE > RED_BYTES ? 2 * STAGE : RED_BYTES;
    static constexpr int CPT = PIECES == 3 ? 128 : 64;                     // accumulator columns per tap
    static constexpr int TCOLS = win_tmem_cols(3 * CPT);
};

template <int PIECES>
__global__ void __launch_bounds__(WIN_THREADS, 1) iins_win_k3_tn_kernel(const IinsWinK3Params wp) {
    using T = WinK3<PIECES>;
    constexpr int TS = T::TS, TROWS = TS * 8;
    extern __shared__ __align__(1024) unsigned char dsm[];
    __shared__ __align__(8) unsigned long long a_full[2], a_empty[2], done;
    __shared__ uint32_t tmem_slot;
    __shared__ double sbias[64];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int conv = blockIdx.y;
    const float* __restrict__ px = wp.x[conv];
    const float* __restrict__ pz = wp.dz[conv];
    const int ntiles = (wp.B + TS - 1) / TS;
    const int my_tiles = (int)blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    IINS_PERSISTENT_PDL_TRIGGER();
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            umma::mbar_init(umma::smem_u32(&a_full[i]), 512);
            umma::mbar_init(umma::smem_u32(&a_empty[i]), 1);
        }
        umma::mbar_init(umma::smem_u32(&done), 1);
        umma::fence_mbar_init();
    }
    if (tid < 64) sbias[tid] = 0.0;
    if (warp == 16) umma::tmem_alloc(umma::smem_u32(&tmem_slot), T::TCOLS);
    for (uint32_t i = (uint32_t)tid * 16u; i < T::SMEM; i += WIN_THREADS * 16u) *reinterpret_cast<uint4*>(dsm + i) = make_uint4(0u, 0u, 0u, 0u);
    umma::fence_async_smem();
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    iins_pdl_wait();

    double bsum[4] = {0.0, 0.0, 0.0, 0.0};
    if (warp == 16) {
        // ------------------------------------------------------------------ MMA issue (warp-uniform code)
        for (int it = 0; it < my_tiles; ++it) {
            const int s = it & 1;
            umma::mbar_wait(umma::smem_u32(&a_full[s]), (uint32_t)(it >> 1) & 1u);
            umma::tc_fence_after();
            const bool leader = umma::elect_one();
            const uint32_t z0 = umma::smem_u32(dsm + s * T::STAGE), x0 = z0 + T::DZ_BYTES;
            for (int ks = 0; ks < TS / 2; ++ks) {                  // 16 rows = 2 samples per instruction
                const uint32_t acc_on = (it | ks) ? 1u : 0u;
                // MN-major: LBO = bytes between 8-row (K) groups, SBO = bytes between 8-channel chunks
                const uint64_t ad = umma::make_desc(z0 + (uint32_t)ks * 256u, 128, T::CSZ);
                for (int t = 0; t < 3; ++t) {
                    const uint32_t col = tmem + (uint32_t)(t * T::CPT);
                    const uint32_t xa = x0 + (uint32_t)t * 16u + (uint32_t)ks * 320u;
                    const uint64_t b0 = umma::make_desc(xa, 160, T::XL);
                    if (leader) umma::mma_bf16_ss(col, ad, b0, umma::make_idesc_bf16(128, 64, 1, 1), acc_on);
                    if (PIECES == 3) {
                        const uint64_t b1 = umma::make_desc(xa + T::PSX, 160, T::XL), b2 = umma::make_desc(xa + 2 * T::PSX, 160, T::XL);
                        if (leader) {
                            umma::mma_bf16_ss(col + 64, ad, b1, umma::make_idesc_bf16(128, 64, 1, 1), acc_on);
                            umma::mma_bf16_ss(col + 64, ad, b2, umma::make_idesc_bf16(128, 64, 1, 1), 1u);
                            umma::mma_bf16_ss(col + 64, ad + ((2 * T::PSZ) >> 4), b0, umma::make_idesc_bf16(128, 64, 1, 1), 1u);
                        }
                    }
                }
            }
            if (leader) {
                umma::commit(umma::smem_u32(&a_empty[s]));
                if (it == my_tiles - 1) umma::commit(umma::smem_u32(&done));
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ producers: ALL 16 other warps, 2 x units + 2 dz units per
        // thread and tile; the raw data of the NEXT tile rides in the register ring while this one is split and stored
        const int pt = tid, c4 = pt & 15, r0 = pt >> 4;            // r0 = 0 .. 31
        const uint32_t coffx = (uint32_t)(c4 >> 1) * T::XL + (uint32_t)(c4 & 1) * 8u;
        const uint32_t coffz = (uint32_t)(c4 >> 1) * T::CSZ + (uint32_t)(c4 & 1) * 8u;
        const long lim = (long)wp.B * 8;                           // rows that exist
        const bool do_bias = wp.db_on != 0;
        float4 v[4];
        auto issue = [&](int u, int it) {                          // units 0-1: x, 2-3: dz; row = r0 + 32 * (u & 1)
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (it < my_tiles) {
                const long grow = (long)((int)blockIdx.x + it * (int)gridDim.x) * TROWS + r0 + 32 * (u & 1);
                if (grow < lim) a = __ldg(reinterpret_cast<const float4*>((u < 2 ? px : pz) + grow * 64 + c4 * 4));
            }
            v[u] = a;
        };
#pragma unroll
        for (int u = 0; u < 4; ++u) issue(u, 0);
        for (int it = 0; it < my_tiles; ++it) {
            const int s = it & 1;
            unsigned char* zb = dsm + s * T::STAGE;
            unsigned char* xb = zb + T::DZ_BYTES;
            if (it >= 2) umma::mbar_wait_suspend(umma::smem_u32(&a_empty[s]), ((uint32_t)(it >> 1) & 1u) ^ 1u);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int row = r0 + 32 * (u & 1);
                uint2 w[3];
                win_split4<PIECES>(v[u], w);
                if (u < 2) {
                    // sample s occupies rows 10 s .. 10 s + 9: [x(1) | x(0) .. x(7) | x(6)] (ReflectionPad1d(1), models.py:993)
                    const int sm = row >> 3, pos = row & 7;
                    unsigned char* d = xb + coffx + (uint32_t)(sm * 160 + (pos + 1) * 16);
                    win_store<PIECES>(w, d, T::PSX);
                    if (pos == 1) win_store<PIECES>(w, d - 32, T::PSX);
                    if (pos == 6) win_store<PIECES>(w, d + 32, T::PSX);
                } else {
                    win_store<PIECES>(w, zb + coffz + (uint32_t)row * 16u, T::PSZ);
                    if (do_bias) { const float4 q = v[u]; bsum[0] += (double)q.x; bsum[1] += (double)q.y; bsum[2] += (double)q.z; bsum[3] += (double)q.w; }
                }
                issue(u, it + 1);
            }
            umma::fence_async_smem();
            umma::mbar_arrive(umma::smem_u32(&a_full[s]));
        }
    }
    // ---- flush
    umma::mbar_wait_suspend(umma::smem_u32(&done), 0);
    umma::tc_fence_after();
    float* red = reinterpret_cast<float*>(dsm);                    // [tap][ci][co]
    constexpr int NRED = 3 * 64 * 64;
    for (int e = tid; e < NRED; e += WIN_THREADS) red[e] = 0.f;
    __syncthreads();
    if (warp < 8) {
        // TMEM lane = piece block * 64 + co; the two warps that share a lane quarter split the 64 input channels of every tap
        const int q = warp & 3, hf = warp >> 2;
        const int r = q * 32 + lane, co = r & 63;
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
        if (PIECES == 3 || r < 64) {                               // warp-uniform
            for (int t = 0; t < 3; ++t) {
#pragma unroll
                for (int c0 = 0; c0 < 32; c0 += 16) {
                    float a[16];
                    umma::tmem_ld16(tl + (uint32_t)(t * T::CPT + hf * 32 + c0), a);
                    if (PIECES == 3) {
                        float b[16];
                        umma::tmem_ld16(tl + (uint32_t)(t * T::CPT + 64 + hf * 32 + c0), b);
#pragma unroll
                        for (int i = 0; i < 16; ++i) a[i] += b[i];
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) atomicAdd(red + (t * 64 + hf * 32 + c0 + i) * 64 + co, a[i]);
                }
            }
        }
    }
    if (warp < 16 && wp.db_on) {
        const int c4 = tid & 15;
#pragma unroll
        for (int k = 0; k < 4; ++k) atomicAdd(&sbias[c4 * 4 + k], bsum[k]);
    }
    umma::tc_fence_before();
    __syncthreads();
    float* pdw = wp.dw[conv];
    for (int e = tid; e < NRED; e += WIN_THREADS) {
        const int co = e & 63, ci = (e >> 6) & 63, t = e >> 12;
        atomicAdd(pdw + ((long)co * 64 + ci) * 3 + t, red[e]);
    }
    if (wp.db_on && wp.db[conv] != nullptr && tid < 64) atomicAdd(wp.db[conv] + tid, (float)sbias[tid]);
    if (warp == 16) umma::tmem_dealloc(tmem, T::TCOLS);
}

template <int PIECES>
void launch_win_k3_v(cudaStream_t st, const IinsWinK3Params& p) {
    using T = WinK3<PIECES>;
    static bool attr = false;
    auto iins_win_k3_tn_kernel_ = iins_win_k3_tn_kernel<PIECES>;
    if (!attr) { cudaFuncSetAttribute(iins_win_k3_tn_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM); attr = true; }
    const int ntiles = (p.B + T::TS - 1) / T::TS;
    int gx = sm_count() / p.nconv;
    if (gx < 1) gx = 1;
    if (gx > ntiles) gx = ntiles;
    IINS_LAUNCH(iins_win_k3_tn_kernel_, dim3(gx, p.nconv, 1), WIN_THREADS, T::SMEM, st, p);
}

template <int NT, int PIECES, int WK, int EPI, int LL>
void launch_win_v(cudaStream_t st, const IinsWinParams& p) {
    const int smem = (int)win_smem_bytes(NT, PIECES, WK, p.ca);
    static int attr_smem = 0;
    auto iins_win_nt_kernel_ = iins_win_nt_kernel<NT, PIECES, WK, EPI, LL>;
    if (smem > attr_smem) { cudaFuncSetAttribute(iins_win_nt_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr_smem = smem; }
    const int ntiles = (p.nt.M + 127) / 128;
    const int grid = ntiles < sm_count() ? ntiles : sm_count();
    IINS_LAUNCH(iins_win_nt_kernel_, grid, WIN_THREADS, smem, st, p);
}

// the instances that are built: (tile width, kind, epilogue, rows per sample)
#define IINS_WIN_INSTANCES(V) \
    V(32, IINS_WIN_S2F, IINS_EPI_PLAIN, 1) V(64, IINS_WIN_S2F, IINS_EPI_PLAIN, 1) \
    V(32, IINS_WIN_S2F, IINS_EPI_IN, 16) V(64, IINS_WIN_S2F, IINS_EPI_IN, 8) \
    V(16, IINS_WIN_S2D, IINS_EPI_PLAIN, 1) V(32, IINS_WIN_S2D, IINS_EPI_PLAIN, 1)

template <int PIECES>
bool launch_win_p(cudaStream_t st, const IinsWinParams& p, int nt, int wk, int epi, int ll) {
#define IINS_WV(NT_, WK_, EPI_, LL_) \
    if (nt == NT_ && wk == WK_ && epi == EPI_ && ll == LL_) { launch_win_v<NT_, PIECES, WK_, EPI_, LL_>(st, p); return true; }
    IINS_WIN_INSTANCES(IINS_WV)
#undef IINS_WV
    return false;
}

}  // namespace

bool iins_win_nt_supported(int nt, int pieces, int wk, int epi, int ll, int ca) {
    if (pieces != 1 && pieces != 3) return false;
    if (ca < 16 || (ca & (ca - 1)) != 0) return false;             // a power of two >= 16
    // the producers hold one unrolled round of 8 units per thread: a tile is ca / 4 (forward) or ca / 8 (data gradient) units
    if (ca > (wk == IINS_WIN_S2F ? 32 : 64)) return false;
    bool inst = false;
#define IINS_WS(NT_, WK_, EPI_, LL_) if (nt == NT_ && wk == WK_ && epi == EPI_ && ll == LL_) inst = true;
    IINS_WIN_INSTANCES(IINS_WS)
#undef IINS_WS
    if (!inst) return false;
    const int nacc = wk == IINS_WIN_S2D ? 2 : 1;
    if (2 * pieces * nt * nacc > 512) return false;                 // two TMEM accumulator sets
    return win_smem_bytes(nt, pieces, wk, ca) <= WIN_SMEM_MAX;
}

bool iins_win_nt_launch(cudaStream_t st, const IinsWinParams& p, int nt, int wk, int epi, int ll) {
    if (!iins_win_nt_supported(nt, p.pieces, wk, epi, ll, p.ca)) return false;
    return p.pieces == 3 ? launch_win_p<3>(st, p, nt, wk, epi, ll) : launch_win_p<1>(st, p, nt, wk, epi, ll);
}

bool iins_win_tn_launch(cudaStream_t st, const IinsTNParams& tn, int pieces) {
    const IinsGeom& g = tn.g;
    if (!(g.stride == 2 && g.ks == 4 && g.pad == 1 && g.mode == IINS_PAD_ZERO && g.Lin == 2 * g.Lout && g.in_layout == IINS_NLC &&
          g.out_layout == IINS_NLC) || g.Lout < 8 || g.Lout > 128 || (g.Lout & (g.Lout - 1)) != 0) return false;
    if ((reinterpret_cast<uintptr_t>(tn.x) & 15) != 0 || (reinterpret_cast<uintptr_t>(tn.dz.dy) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(tn.dz.y) & 15) != 0) return false;
    IinsWinTNParams p;
    memset(&p, 0, sizeof(p));
    p.tn = tn;
    while ((1 << p.lsh_out) < g.Lout) ++p.lsh_out;
#define IINS_WTN(CI_, CO_) \
    if (g.Cin == CI_ && g.Cout == CO_) { if (pieces == 3) launch_win_tn_v<CI_, CO_, 3>(st, p); else launch_win_tn_v<CI_, CO_, 1>(st, p); return true; }
    IINS_WTN(16, 32) IINS_WTN(32, 64)
#undef IINS_WTN
    return false;
}

bool iins_win_k3_tn_launch(cudaStream_t st, int B, int nconv, const float* const* xs, const float* const* dzs, float* const* dws,
                           float* const* dbs, int pieces) {
    if (B < 1 || nconv < 1 || nconv > 8 || (pieces != 1 && pieces != 3)) return false;
    IinsWinK3Params p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.nconv = nconv;
    for (int i = 0; i < nconv; ++i) {
        if ((reinterpret_cast<uintptr_t>(xs[i]) & 15) != 0 || (reinterpret_cast<uintptr_t>(dzs[i]) & 15) != 0 || dws[i] == nullptr) return false;
        p.x[i] = xs[i]; p.dz[i] = dzs[i]; p.dw[i] = dws[i]; p.db[i] = dbs[i];
        if (dbs[i] != nullptr) p.db_on = 1;
    }
    if (pieces == 3) launch_win_k3_v<3>(st, p); else launch_win_k3_v<1>(st, p);
    return true;
}
