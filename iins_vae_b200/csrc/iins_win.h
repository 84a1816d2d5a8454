// Persistent "window" kernels for the stride-2 convolutions (host interface, see iins_win.cu).
//
// The k4 / stride-2 / zero-pad-1 convolutions of the two encoders (models.py:156-162, 268-274) are memory-shaped GEMMs
// (16..64 output channels, K = 64..128): the per-layer kernels of iins_tc.cuh spend their time building the im2col tile
// (every input element is gathered, split into bf16 pieces and stored once per tap that reads it) and pay one CTA
// prologue / epilogue latency per 128-row tile.  Here a persistent CTA keeps the tile's input rows RESIDENT in shared
// memory, split ONCE, in a layout where every tap is the same UMMA descriptor with a shifted start address (no im2col
// copy), the layer's packed weights stay resident for the CTA's lifetime, and producer warps, the MMA warp and the
// epilogue warps overlap across tiles (two operand stages, two TMEM accumulators).
#pragma once
#ifndef IINS_CPUSIM
#include "iins_tc.cuh"

enum { IINS_WIN_S2F = 0,      // forward of a k4 / s2 / p1 convolution: two parity planes of the input rows
       IINS_WIN_S2D = 1 };    // its data gradient split by the parity of the input position: one plane of dz rows, two accumulators

struct IinsWinParams {
    IinsNTParams nt;             // S2F: the forward problem; S2D: the PARITY problem (M = B * Lout rows, K = 2 * Cout, Lrow = Lout)
    const uint16_t* wpack;       // S2F: forward pack (kind 0); S2D: even-position pack
    const uint16_t* wpack_odd;   // S2D: odd-position pack
    int pieces;                  // 3 (fp32-grade) or 1 (bf16)
    int nkb;                     // K blocks of 32 of one pack
    int ca;                      // channels of the resident operand (S2F: Cin, S2D: Cout), a multiple of 16
    int lsh_in;                  // S2F: log2(Lin);  S2D: log2(Lout) (rows per sample of the resident operand)
};

// false: no instance for this (tile width, kind, epilogue, rows per sample) or it does not fit shared memory / TMEM
bool iins_win_nt_supported(int nt, int pieces, int wk, int epi, int ll, int ca);
bool iins_win_nt_launch(cudaStream_t st, const IinsWinParams& p, int nt, int wk, int epi, int ll);
// weight gradient of a k4 / s2 / p1 convolution (tn.M = B * Lout set by the caller); false: no instance for (Cin, Cout)
bool iins_win_tn_launch(cudaStream_t st, const IinsTNParams& tn, int pieces);
// weight (+ bias) gradients of up to 8 k3 / reflect-pad / 64 -> 64 channel convolutions over (B, 8, 64) tensors (the residual trunk)
bool iins_win_k3_tn_launch(cudaStream_t st, int B, int nconv, const float* const* xs, const float* const* dzs, float* const* dws,
                           float* const* dbs, int pieces);
#endif
