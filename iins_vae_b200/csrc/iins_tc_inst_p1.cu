// plain bf16 (one piece) instances of iins_tc_nt_kernel
#include "iins_tc_inst.cuh"
bool iins_launch_tc_nt_p1(cudaStream_t st, const IinsTCParams& tp, dim3 grid, int nt, int akind, int epi, int ll) {
    return launch_tc_nt_variant<1>(st, tp, grid, nt, akind, epi, ll);
}
