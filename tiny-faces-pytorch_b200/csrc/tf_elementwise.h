// Internal C++ interface of the elementwise layer (tf_elementwise.cu), used by tf_model.cu.
#pragma once
#include <cuda_runtime.h>

namespace tfe {
constexpr int BN_BWD_SLOTS = 32;             // accumulation rows of the BN-backward column reduction
constexpr int COLREDUCE_MAX_BLOCKS = 592;     // partial buffers hold COLREDUCE_MAX_BLOCKS * 2 * C floats
// `mode` of every activation producer: 0 = store as is, 1 = round to TF32 (fast mode operands),
// 2 = store the (hi, lo) TF32 split into (out, out_lo) (3xTF32 parity mode)
int bn_finalize_train(float* partial, int nblk, long long M, int C, const float* gamma, const float* beta, float eps,
                      float momentum, float* run_mean, float* run_var, float* scale, float* shift, float* save_mean,
                      float* save_rstd, cudaStream_t st);
int bn_scale_shift_eval(int C, const float* gamma, const float* beta, const float* run_mean, const float* run_var,
                        float eps, float* scale, float* shift, cudaStream_t st);
struct BnEvalJob { const float *gamma, *beta, *run_mean, *run_var; float *scale, *shift; int begin; };
int bn_scale_shift_eval_batched(const BnEvalJob* jobs_device, int njobs, int total_channels, float eps, cudaStream_t st);
int bn_apply(const float* y, const float* scale, const float* shift, const float* res, const float* rscale,
             const float* rshift, int relu, long long M, int C, float* out, float* out_lo, int mode,
             unsigned int* mask_out /* optional 1-bit ReLU mask, (M*C+31)/32 words */, cudaStream_t st);
int bn_backward(const float* dout, const float* act, const unsigned int* mask /* either may gate dout */, const float* y, const float* save_mean, const float* save_rstd,
                const float* gamma, long long M, int C, float* dgamma, float* dbeta, float* dy, float* dy_lo,
                float* gmask_out, int mode, float* slots /* [BN_BWD_SLOTS][2][C] zeros, left zeroed */, float* coef, cudaStream_t st);
// per-channel (sum, sum^2) partials of y[M, C] -> partial[*nblk][2][C] (bn_finalize_train reduces them)
int column_stats(const float* y, long long M, int C, float* partial, int* nblk, cudaStream_t st);
int column_sum(const float* a, long long M, int C, int Cout, float* out, float* partial, cudaStream_t st, int accumulate = 0);
int masked_add(const float* a, const float* act, const float* b, long long n, float* out, cudaStream_t st);
int stem_im2col(const float* x_nchw, int B, int H, int W, int Ho, int Wo, int K_pad, float* col, float* col_lo, int mode,
                cudaStream_t st);
int maxpool_fwd(const float* x, int B, int H, int W, int C, int Ho, int Wo, float* out, float* out_lo, int mode,
                unsigned char* argmax /* optional [B,Ho,Wo,C] window index of the first maximum */, cudaStream_t st);
int maxpool_bwd(const unsigned char* argmax, const float* dout, int B, int H, int W, int C, int Ho, int Wo, float* dx, cudaStream_t st);
int zero_insert2(const float* xs, const float* xs_lo, int B, int H, int W, int C, float* out, float* out_lo, cudaStream_t st);
int pack_weight(const float* w, int O_src, int I_src, int taps, int transpose, int O_pad, int I_pad, float* dst,
                float* dst_lo, int mode, cudaStream_t st);
struct PackJob {                              // one weight tensor of a batched packing launch
    const float* src; float* dst; float* dst_lo;
    const float* oscale;                      // optional per-output-channel factor folded into the weights (eval-mode BN scale)
    int O_src, I_src, taps, transpose, O_pad, I_pad;
    long long begin;                          // first global element index of this job
};
int pack_weights_batched(const PackJob* jobs_device, int njobs, long long total, int mode, cudaStream_t st);
int unpack_wgrad(const float* dwp, int O, int I, int taps, int I_pad, float* dw, cudaStream_t st);
int head_combine_fwd(const float* s3, const float* s4, const float* up, int B, int H3, int W3, int H4, int W4, int Cn,
                     int Cp, float* out_nchw, cudaStream_t st);
// ds3_lo / ds4_lo (both or neither): parity mode -- the gradients are stored as exact (hi, lo) TF32 splits
int head_combine_bwd(const float* dout_nchw, const float* up, int B, int H3, int W3, int H4, int W4, int Cn, int Cp,
                     float* ds3, float* ds4, cudaStream_t st, float* ds3_lo = nullptr, float* ds4_lo = nullptr);
int extract_upsample_diag(const float* w, int Cn, float* up, float* offdiag_max, cudaStream_t st);
}  // namespace tfe
