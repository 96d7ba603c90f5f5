// Per-kernel C-ABI entry points of the elementwise layer (declared in include/tinyfaces_b200.h): the BatchNorm /
// ReLU / residual, max-pool and head kernels the whole-model executor (tf_model.cu) chains between its GEMMs, exposed
// one by one so that each can be checked against its torch counterpart ON IDENTICAL INPUTS:
//   nn.BatchNorm2d (+ shortcut add + ReLU) fwd/bwd   torchvision resnet.py:143-163 (Bottleneck.forward)
//   nn.MaxPool2d(3, 2, 1) fwd/bwd                    torchvision resnet.py:197-204 (stem)
//   ConvTranspose2d(4, 2, 1) + crop + add fwd/bwd    /root/reference/tinyfaces/models/model.py:104-126
// These are the SAME kernels the executor launches (tfe:: functions), not test doubles.
#include "tf_common.cuh"
#include "tf_elementwise.h"

namespace {
struct BnWs { float* partial; float* slots; float* coef; float* scale; float* shift; };
size_t bn_ws_floats() {
    return (size_t)tfe::COLREDUCE_MAX_BLOCKS * 2 * 1024 + (size_t)tfe::BN_BWD_SLOTS * 2 * 1024 + 3 * 1024 + 2 * 1024 + 1024;
}
int carve(void* workspace, size_t bytes, BnWs* w) {
    TF_REQUIRE(workspace && bytes >= bn_ws_floats() * sizeof(float), "tf_bn_*: workspace too small (tf_bn_workspace_bytes)");
    float* p = reinterpret_cast<float*>(workspace);
    w->partial = p; p += (size_t)tfe::COLREDUCE_MAX_BLOCKS * 2 * 1024;
    w->slots = p; p += (size_t)tfe::BN_BWD_SLOTS * 2 * 1024;
    w->coef = p; p += 3 * 1024;
    w->scale = p; p += 1024;
    w->shift = p;
    return TF_OK;
}
bool pow2(int c) { return c >= 64 && c <= 1024 && (c & (c - 1)) == 0; }
}  // namespace

TF_API int tf_bn_workspace_bytes(size_t* bytes) {
    TF_REQUIRE(bytes, "tf_bn_workspace_bytes: null");
    *bytes = bn_ws_floats() * sizeof(float);
    return TF_OK;
}

// out = [relu]( BN_train(y) [+ res] ): batch statistics over the M rows of y[M, C] (biased variance), running statistics
// updated in place with `momentum` (unbiased variance), save_mean / save_rstd for the backward, optional 1-bit ReLU mask.
TF_API int tf_bn_train_fwd(const float* y, int64_t M, int C, const float* gamma, const float* beta, float eps, float momentum,
                           float* run_mean, float* run_var, const float* res, int relu, int round_tf32, float* out,
                           uint32_t* relu_mask, float* save_mean, float* save_rstd, void* workspace, size_t workspace_bytes,
                           void* stream) {
    TF_REQUIRE(y && gamma && beta && out && save_mean && save_rstd && M > 0 && pow2(C), "tf_bn_train_fwd: bad args (C must be a power of two in [64, 1024])");
    BnWs w;
    int rc = carve(workspace, workspace_bytes, &w);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    int nblk = 0;
    if ((rc = tfe::column_stats(y, M, C, w.partial, &nblk, st))) return rc;
    if ((rc = tfe::bn_finalize_train(w.partial, nblk, M, C, gamma, beta, eps, momentum, run_mean, run_var, w.scale, w.shift,
                                     save_mean, save_rstd, st))) return rc;
    return tfe::bn_apply(y, w.scale, w.shift, res, nullptr, nullptr, relu, M, C, out, nullptr, round_tf32 ? 1 : 0, relu_mask, st);
}

// eval mode: out = [relu]( (y - run_mean) / sqrt(run_var + eps) * gamma + beta [+ res] )
TF_API int tf_bn_eval_fwd(const float* y, int64_t M, int C, const float* gamma, const float* beta, const float* run_mean,
                          const float* run_var, float eps, const float* res, int relu, int round_tf32, float* out,
                          void* workspace, size_t workspace_bytes, void* stream) {
    TF_REQUIRE(y && gamma && beta && run_mean && run_var && out && M > 0 && pow2(C), "tf_bn_eval_fwd: bad args");
    BnWs w;
    int rc = carve(workspace, workspace_bytes, &w);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = tfe::bn_scale_shift_eval(C, gamma, beta, run_mean, run_var, eps, w.scale, w.shift, st))) return rc;
    return tfe::bn_apply(y, w.scale, w.shift, res, nullptr, nullptr, relu, M, C, out, nullptr, round_tf32 ? 1 : 0, nullptr, st);
}

// Backward of out = relu?(BN_train(y) ...): g = dout (* relu_mask bit), dgamma = sum g*xhat, dbeta = sum g,
// dy = gamma*rstd*(g - mean(g) - xhat*mean(g*xhat)); g_out (optional) receives g itself (the shortcut's gradient).
TF_API int tf_bn_bwd(const float* dout, const uint32_t* relu_mask, const float* y, const float* save_mean, const float* save_rstd,
                     const float* gamma, int64_t M, int C, float* dgamma, float* dbeta, float* dy, float* g_out, int round_tf32,
                     void* workspace, size_t workspace_bytes, void* stream) {
    TF_REQUIRE(dout && y && save_mean && save_rstd && gamma && dy && M > 0 && pow2(C), "tf_bn_bwd: bad args");
    BnWs w;
    int rc = carve(workspace, workspace_bytes, &w);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    TF_CHECK_CUDA(cudaMemsetAsync(w.slots, 0, (size_t)tfe::BN_BWD_SLOTS * 2 * 1024 * sizeof(float), st));
    return tfe::bn_backward(dout, nullptr, relu_mask, y, save_mean, save_rstd, gamma, M, C, dgamma, dbeta, dy, nullptr, g_out,
                            round_tf32 ? 1 : 0, w.slots, w.coef, st);
}

// nn.MaxPool2d(kernel 3, stride 2, padding 1) on NHWC: out [B, Ho, Wo, C], Ho = (H-1)/2+1; argmax (optional, uint8 window
// index kh*3+kw of the FIRST maximum in scan order -- ATen's tie rule) feeds the backward.
TF_API int tf_maxpool_fwd(const float* x, int B, int H, int W, int C, float* out, uint8_t* argmax, void* stream) {
    TF_REQUIRE(x && out && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "tf_maxpool_fwd: bad args (C % 4 == 0)");
    return tfe::maxpool_fwd(x, B, H, W, C, (H - 1) / 2 + 1, (W - 1) / 2 + 1, out, nullptr, 0, argmax, (cudaStream_t)stream);
}
TF_API int tf_maxpool_bwd(const uint8_t* argmax, const float* dout, int B, int H, int W, int C, float* dx, void* stream) {
    TF_REQUIRE(argmax && dout && dx && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "tf_maxpool_bwd: bad args");
    return tfe::maxpool_bwd(argmax, dout, B, H, W, C, (H - 1) / 2 + 1, (W - 1) / 2 + 1, dx, (cudaStream_t)stream);
}

// model.py:104-126: out[b,c,y,x] = s3[b,y,x,c] + crop(ConvTranspose2d(s4; up_w, k4 s2 p1))[b,c,y,x], NCHW result.
// s3 [B,H3,W3,Cp], s4 [B,H4,W4,Cp] NHWC with the Cn real channels first; up_w: the [Cn,Cn,4,4] ConvTranspose2d weight,
// which must be diagonal (the reference's frozen bilinear kernel); *offdiag_host (optional) receives max |off-diagonal|.
TF_API int tf_head_workspace_bytes(int Cn, size_t* bytes) {
    TF_REQUIRE(bytes && Cn > 0, "tf_head_workspace_bytes: bad args");
    *bytes = ((size_t)Cn * 16 + 64) * sizeof(float);
    return TF_OK;
}
TF_API int tf_head_upsample_add_fwd(const float* s3, const float* s4, const float* up_w, int B, int H3, int W3, int H4, int W4,
                                    int Cn, int Cp, float* out_nchw, void* workspace, size_t workspace_bytes, void* stream) {
    TF_REQUIRE(s3 && s4 && up_w && out_nchw && workspace && workspace_bytes >= ((size_t)Cn * 16 + 64) * 4, "tf_head_upsample_add_fwd: bad args");
    TF_REQUIRE(Cp >= Cn && Cp % 4 == 0 && 2 * H4 >= H3 && 2 * W4 >= W3, "tf_head_upsample_add_fwd: bad shapes");
    float* up = reinterpret_cast<float*>(workspace);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = tfe::extract_upsample_diag(up_w, Cn, up, up + (size_t)Cn * 16, st);
    if (rc) return rc;
    return tfe::head_combine_fwd(s3, s4, up, B, H3, W3, H4, W4, Cn, Cp, out_nchw, st);
}
TF_API int tf_head_upsample_add_bwd(const float* dout_nchw, const float* up_w, int B, int H3, int W3, int H4, int W4, int Cn,
                                    int Cp, float* ds3, float* ds4, void* workspace, size_t workspace_bytes, void* stream) {
    TF_REQUIRE(dout_nchw && up_w && ds3 && ds4 && workspace && workspace_bytes >= ((size_t)Cn * 16 + 64) * 4, "tf_head_upsample_add_bwd: bad args");
    float* up = reinterpret_cast<float*>(workspace);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = tfe::extract_upsample_diag(up_w, Cn, up, up + (size_t)Cn * 16, st);
    if (rc) return rc;
    return tfe::head_combine_bwd(dout_nchw, up, B, H3, W3, H4, W4, Cn, Cp, ds3, ds4, st);
}
