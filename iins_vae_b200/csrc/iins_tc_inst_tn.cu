// weight-gradient tensor-core kernel instances (iins_tc_tn_kernel)
#include "iins_launchers.h"

template <int NT, int PIECES>
static void launch_tc_tn_tp(cudaStream_t st, const IinsTCTNParams& tp, dim3 grid) {
    constexpr int smem = 2 * (3 * 8192 + 3 * (NT / 8) * 32 * 16);
    static bool attr = false;
    auto iins_tc_tn_kernel_ = iins_tc_tn_kernel<NT, PIECES>;
    if (!attr) { cudaFuncSetAttribute(iins_tc_tn_kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
    IINS_LAUNCH(iins_tc_tn_kernel_, grid, 288, smem, st, tp);
}
template <int NT>
static void launch_tc_tn_t(cudaStream_t st, const IinsTCTNParams& tp, dim3 grid) {
    if (tp.pieces == 3) launch_tc_tn_tp<NT, 3>(st, tp, grid);
    else launch_tc_tn_tp<NT, 1>(st, tp, grid);
}
void iins_launch_tc_tn(cudaStream_t st, const IinsTCTNParams& tp, dim3 grid, int nt) {
    if (nt == 16) launch_tc_tn_t<16>(st, tp, grid);
    else if (nt == 32) launch_tc_tn_t<32>(st, tp, grid);
    else launch_tc_tn_t<64>(st, tp, grid);
}
