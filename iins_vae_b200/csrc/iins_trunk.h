// Fused residual-trunk kernels (host interface).  The L = 8 trunk of the range encoder (models.py:166-167: three
// ResidualBlock1d with InstanceNorm) and of the decoder (models.py:416-417: three ResidualBlock1d with AdaIN) is a chain of
// 2 * n_residual k3 reflect-pad convolutions over (B, 8, 64) activations.  One persistent CTA keeps a 16-sample tile
// resident in shared memory as bf16 pieces and walks the whole chain: no activation round trip through HBM between the
// convolutions, one launch instead of 2 * n_residual.
#pragma once
#ifndef IINS_CPUSIM
#include <cuda_runtime.h>
#include <stdint.h>

#define IINS_TRUNK_MAX_CONVS 32

struct IinsTrunkLayer {
    const float* bias;         // [64]
    float* y;                  // (B, 8, 64) output (post norm / activation / residual)
    float* xhat;               // (B, 8, 64) normalised pre-affine values (saved for backward)
    float* rstd;               // (B, 64)
    int adain_off_b, adain_off_w;   // AdaIN: offsets of this layer's bias / weight inside a sample's parameter row
};

struct IinsTrunkFwdParams {
    int B;                     // samples
    int nconv;                 // 2 * n_residual
    int pieces;                // 3 (fp32-grade) or 1 (bf16)
    int use_tmap;              // weight ring by tensor-map TMA (cp.async.bulk.tensor); 0: plain bulk copies
    const float* x;            // (B, 8, 64) trunk input
    const float* adain;        // (B, adain_ld) AdaIN parameters, or nullptr: plain InstanceNorm
    int adain_ld;
    const uint16_t* wpack;     // packed weights of conv 0 (the nconv packs are contiguous, 64 * 192 * pieces bf16 each)
    IinsTrunkLayer layer[IINS_TRUNK_MAX_CONVS];
};

// returns false when the kernel cannot take this configuration (the caller then runs the layer-by-layer path)
bool iins_trunk_forward_launch(cudaStream_t st, const IinsTrunkFwdParams& p);

// ---- backward: the data-gradient chain of the same trunk, each convolution's data gradient followed (in the same epilogue)
// by the InstanceNorm / AdaIN backward of the layer below it.  The weight gradients are separate kernels: they read the
// dz tensors this kernel leaves in HBM.
struct IinsTrunkBwdLayer {
    const float* xhat;         // (B, 8, 64) saved normalised values of this convolution's norm
    const float* rstd;         // (B, 64)
    float* dz;                 // OUT (B, 8, 64): gradient w.r.t. the convolution's pre-norm output
    int adain_off_b, adain_off_w;
};

struct IinsTrunkBwdParams {
    int B, nconv, pieces;
    int use_tmap;
    const float* dh;           // (B, 8, 64) gradient w.r.t. the trunk output
    float* dx;                 // OUT (B, 8, 64) gradient w.r.t. the trunk input (may be nullptr when `pre` is given)
    float* dh_scratch;         // (B, 8, 64) scratch: gradient w.r.t. the output of the residual block being left
    const float* adain;        // (B, adain_ld) AdaIN parameters or nullptr
    float* dadain;             // OUT (B, adain_ld): AdaIN parameter gradients (bias grad at off_b, weight grad at off_w)
    int adain_ld;
    const uint16_t* wpack;     // data-gradient packs, contiguous, in DESCENDING convolution order (conv nconv-1 first)
    IinsTrunkBwdLayer layer[IINS_TRUNK_MAX_CONVS];
    // optional: the InstanceNorm (+ReLU) layer that produced the trunk input; its backward then runs in the last epilogue
    IinsTrunkBwdLayer pre;
    int pre_relu;
};
bool iins_trunk_backward_launch(cudaStream_t st, const IinsTrunkBwdParams& p);
#endif
