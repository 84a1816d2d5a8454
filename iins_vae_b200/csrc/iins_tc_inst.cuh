// Shared body of iins_tc_inst_p{1,3}.cu: every (tile width, operand kind, epilogue kind, rows per sample) instance of the
// forward / data-gradient tensor-core kernel for ONE piece count.
#pragma once
#include "iins_launchers.h"

template <int NT, int PIECES, int AKIND, int EPI, int LL>
static void launch_tc_nt_v(cudaStream_t st, const IinsTCParams& tp, dim3 grid) {
    constexpr int smem = 2 * (3 * 4 * (128 * 16 + 64) + 3 * 4 * NT * 16) + 8192;      // A stages use the padded chunk stride
    static bool attr = false;
    auto iins_tc_nt_kernel_ = iins_tc_nt_kernel<NT, PIECES, AKIND, EPI, LL>;
    if (!attr) { cudaFuncSetAttribute(iins_tc_nt_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
    IINS_LAUNCH(iins_tc_nt_kernel_, grid, 288, smem, st, tp);
}

template <int PIECES>
static bool launch_tc_nt_variant(cudaStream_t st, const IinsTCParams& tp, dim3 grid, int nt, int akind, int epi, int ll) {
#define IINS_V(NT_, AK_, EPI_, LL_) \
    if (nt == NT_ && akind == AK_ && epi == EPI_ && ll == LL_) { launch_tc_nt_v<NT_, PIECES, AK_, EPI_, LL_>(st, tp, grid); return true; }
    IINS_V(16, 0, IINS_EPI_PLAIN, 1) IINS_V(32, 0, IINS_EPI_PLAIN, 1) IINS_V(64, 0, IINS_EPI_PLAIN, 1)
    IINS_V(16, 1, IINS_EPI_PLAIN, 1) IINS_V(32, 1, IINS_EPI_PLAIN, 1) IINS_V(64, 1, IINS_EPI_PLAIN, 1)
    IINS_V(16, 2, IINS_EPI_PLAIN, 1) IINS_V(32, 2, IINS_EPI_PLAIN, 1) IINS_V(64, 2, IINS_EPI_PLAIN, 1)
    IINS_V(64, 0, IINS_EPI_IN, 8) IINS_V(64, 0, IINS_EPI_IN, 16) IINS_V(32, 0, IINS_EPI_IN, 8) IINS_V(32, 0, IINS_EPI_IN, 16)
    IINS_V(32, 0, IINS_EPI_LN, 16) IINS_V(16, 0, IINS_EPI_LN, 32)
    IINS_V(64, 1, IINS_EPI_NBWD, 8) IINS_V(32, 1, IINS_EPI_NBWD, 16) IINS_V(16, 1, IINS_EPI_NBWD, 32)
    IINS_V(16, 0, IINS_EPI_SMEM, 1) IINS_V(32, 0, IINS_EPI_SMEM, 1) IINS_V(64, 0, IINS_EPI_SMEM, 1)
    IINS_V(16, 1, IINS_EPI_SMEM, 1) IINS_V(32, 1, IINS_EPI_SMEM, 1) IINS_V(64, 1, IINS_EPI_SMEM, 1)
#undef IINS_V
    return false;
}
