// fp32-grade (three bf16 pieces) instances of iins_tc_nt_kernel
#include "iins_tc_inst.cuh"
bool iins_launch_tc_nt_p3(cudaStream_t st, const IinsTCParams& tp, dim3 grid, int nt, int akind, int epi, int ll) {
    return launch_tc_nt_variant<3>(st, tp, grid, nt, akind, epi, ll);
}
