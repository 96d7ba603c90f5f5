// Internal C++ interface of the tcgen05 convolution GEMMs (tf_conv_gemm.cu), used by tf_model.cu.
#pragma once
#include <cuda_runtime.h>

namespace tfg {
struct ConvArgs {
    const float* x; const float* x_lo;     // NHWC input (+ low-order split part in 3xTF32 mode)
    int B, H, W, Cin;
    const float* w; const float* w_lo;     // packed [Cout][k*k][Cin]
    int Cout, ksize;                       // padding ksize/2
    int stride;                            // 0/1 or 2 (strided TMA traversal); output is ceil(H/s) x ceil(W/s)
    const float* scale; const float* shift;// optional per-channel epilogue (shift alone = bias)
    int relu, round_out, accumulate;
    const float* res; const unsigned int* res_mask;   // optional (flat 1x1 only): y = conv + (mask bit ? res : 0), 1 bit / element
    float* y;                              // NHWC output
    float* stats_partial;                  // optional: per-channel (sum, sum^2) partials of the raw output,
    int* stats_blocks;                     //   [*stats_blocks][2][Cout] (HOST out: number of partial rows written)
};
int conv_fprop(const ConvArgs& a, cudaStream_t st);
// input gradient of a stride-2 conv as 4 parity-class GEMMs (no zero insertion): x = dy at ceil(H/2) x ceil(W/2), y = dx at H x W
int conv_dgrad_s2(const ConvArgs& a, cudaStream_t st);

struct WgradArgs {
    const float* x; const float* x_lo;
    const float* dy; const float* dy_lo;
    int B, H, W, Cin, Cout, ksize;
    int stride;                            // x is [B,H,W,Cin]; dy is [B,ceil(H/s),ceil(W/s),Cout]
    float* dw;                             // packed [Cout][k*k][Cin], accumulated into
};
int conv_wgrad(const WgradArgs& a, cudaStream_t st);
int debug_flag(int key);
int debug_epoch();                         // changes whenever a switch is set (invalidates cached plans)                  // tf_debug_set(key, value) experiment switches (0 = default behaviour)
}  // namespace tfg
