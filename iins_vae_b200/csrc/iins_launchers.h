// Host-side launchers of the tensor-core kernel instances.  The instances live in their own translation units
// (iins_tc_inst_*.cu) so that the library builds in parallel and a change to the launch plans does not recompile them.
#pragma once
#ifndef IINS_CPUSIM
#include "iins_tc.cuh"
// (tile width, operand kind, epilogue kind, rows per sample) -> kernel instance; false if that instance is not built
bool iins_launch_tc_nt_p3(cudaStream_t st, const IinsTCParams& tp, dim3 grid, int nt, int akind, int epi, int ll);
bool iins_launch_tc_nt_p1(cudaStream_t st, const IinsTCParams& tp, dim3 grid, int nt, int akind, int epi, int ll);
void iins_launch_tc_tn(cudaStream_t st, const IinsTCTNParams& tp, dim3 grid, int nt);
#endif
