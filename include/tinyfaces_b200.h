/* tinyfaces_b200 -- C ABI of the Blackwell-native hot path of varunagrawal/tiny-faces-pytorch.
 *
 * One shared library (libtinyfaces_b200.so), plain pointers and sizes, no torch types.  The reference has no
 * FFI of its own (it is pure Python over torch / torchvision); each entry point below names the reference
 * interface it stands in for.  Conventions:
 *   - every pointer is a DEVICE pointer unless the comment says HOST;
 *   - the caller owns all memory: outputs and scratch are passed in; `*_workspace_bytes` reports scratch size;
 *   - variable-size outputs use a caller capacity + a device-side count;
 *   - `stream` is a cudaStream_t; all work is enqueued on it (thread-safe for distinct streams);
 *   - return 0 on success, negative on error (tf_last_error_string() has the text);
 *   - there is no CPU fallback anywhere behind this interface.
 */
#ifndef TINYFACES_B200_H
#define TINYFACES_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* tf_last_error_string(void);
int tf_version(void);
/* Experiment switches (test / A-B hooks only; 0 = the default behaviour of every key):
 *   0: wgrad plain store, no split-K   1: force accumulate in tf_conv2d_nhwc   2: force the GEMM tile width (64/128/256)
 *   3: 2 = tail-wave split-K off       4: 1 = 2-CTA GEMM everywhere it fits, 2 = never
 *   5: 1/2/3 = weight-gradient schedule (before its dgrad / after it / deferred to the next block [default])
 *   6: 1 = backward chain on the caller's stream (no priority stream)      7: 1 = force the BN-statistics epilogue
 *   8: 2 = materialise G and reduce-add (no residual epilogue in the conv1 dgrad)
 *   9: bit 0 = bn_apply iterates descending, bit 1 = BN-backward column reduction ascending
 *  10: 1 = 3x3 weight gradients with one accumulator tile per CTA (default: two)
 *  11: 1 = stride-2 dgrad by zero insertion (old path)      12: 1 = inference conv3 without the fused shortcut epilogue
 *  13: 1 = tf_nms counts its IoU pair tests (tf_nms_sweep_stats)   14: 1/2/3 = tf_nms stops after sort / sweep / resolution
 *      (stage timing; the result is then invalid)
 *  15: bit 0 / 1 / 2 = plain instead of streaming (ld.global.cs) loads in bn_apply / bn_bwd_apply / colreduce<1> (A/B) */
int tf_debug_set(int key, int value);
int tf_gemm_error_flag(int* value_host);         /* HOST out: non-zero if a tcgen05 pipeline wait timed out */

/* ---- greedy NMS: replaces torchvision.ops.nms as called at tinyfaces/evaluation.py:84 (float64 CPU tensors).
 * boxes [n,4] (x1,y1,x2,y2), scores [n]; elem_bytes 8 (float64) or 4 (float32).  keep [n] int64 receives the
 * kept ORIGINAL indices in descending-score order; num_keep is a device int64.  Bit-identical to the CPU op:
 * stable descending sort, area (x2-x1)*(y2-y1), suppress iff inter/(a_i+a_j-inter) > thr. */
int tf_nms_workspace_bytes(int64_t n, int elem_bytes, size_t* bytes_host);
int tf_nms_set_algorithm(int algo);   /* test hook: 0 auto, 1 blocked bit-matrix, 2 sort-and-sweep, 3 size-class grid */
int tf_nms(const void* boxes, const void* scores, int64_t n, int elem_bytes, double iou_threshold, int64_t* keep,
           int64_t* num_keep, void* workspace, size_t workspace_bytes, void* stream);
/* tf_nms with an explicit algorithm (0 auto, 1 blocked bit-matrix; 2 / 3: conflict edges from a 1-D sort-and-sweep along x /
 * from the size-class grid, then parallel fixed-point resolution; auto = 2 for 4096 <= n <= 3e5, 3 above, 1 below or when
 * thr < 0).  Both entry
 * points only ENQUEUE work on `stream` (no host synchronisation, graph-capturable).  The sort-and-sweep path keeps its
 * conflict-edge list in the workspace (default capacity 128 edges per box; a larger workspace is used in full): if the
 * list overflows (or, for the grid, a box is larger than 2^16 / smaller than 2^-16), *num_keep is set to -1 on the device and
 * the caller re-runs with algorithm 1, which is exact for every input.  Score order follows torch.sort (NaN first, -0.0 == +0.0, stable). */
int tf_nms_algo(const void* boxes, const void* scores, int64_t n, int elem_bytes, double iou_threshold, int algorithm,
                int64_t* keep, int64_t* num_keep, void* workspace, size_t workspace_bytes, void* stream);
/* diagnostics of the last sort-and-sweep run on this workspace (SYNCHRONISES; not on the data path): out4_host =
 * {conflict edges, IoU pair tests (counted under tf_debug_set(13, 1) only), resolution rounds, edge capacity} */
int tf_nms_sweep_stats(int64_t n, int elem_bytes, void* workspace, size_t workspace_bytes, int64_t* out4_host, void* stream);

/* ---- threshold + ordered compaction + anchor decode: replaces get_bboxes + regression_refinement
 * (tinyfaces/models/utils.py:4-100) and the sigmoid / D2H / transpose of tinyfaces/evaluation.py:61-71.
 * cls/reg/prob are float32 maps addressed by (b,y,x,channel) element strides, so both the NCHW model output and
 * the reference's NHWC numpy arrays are accepted; prob == NULL fuses the sigmoid.  templates_host: HOST [T,5]
 * float64.  invalid_x_mask reproduces the shipped utils.py:44 quirk (bit x: heat-map column x is zeroed);
 * invalid_t_mask masks templates instead.  Candidates are emitted in C order over (b,y,x,c) (utils.py:46-47). */
int tf_decode_workspace_bytes(int64_t B, int64_t H, int64_t W, size_t* bytes_host);
int tf_decode(const float* cls, const float* reg, const float* prob, const int64_t* cls_strides_host,
              const int64_t* reg_strides_host, int B, int H, int W, int T, const double* templates_host,
              float prob_thresh, uint32_t invalid_x_mask, uint32_t invalid_t_mask, const int64_t* rf_stride_host,
              const int64_t* rf_offset_host, double scale, double* boxes, double* scores, int64_t* src_index,
              int64_t capacity, int64_t* count, void* workspace, size_t workspace_bytes, void* stream);

/* ---- detection loss: replaces DetectionCriterion.forward (tinyfaces/models/loss.py:59-93) and, for the device
 * sampler, balance_sampling (tinyfaces/models/utils.py:103-163).  NCHW: output [B,5T,H,W], class_map [B,T,H,W],
 * regression_map [B,4T,H,W]; HW = H*W. */
int tf_detloss_ohem(const float* output, float* class_map /* in place */, int B, int T, int64_t HW, float thresh,
                    void* stream);
int tf_detloss_fwd_bwd(const float* output, const float* labels, const float* regression_map, int B, int T, int64_t HW,
                       float reg_weight, float* grad_output, double* sums /* [2], accumulated */, void* stream);
int tf_detloss_sample_workspace_bytes(int B, size_t* bytes_host);
int tf_detloss_sample_device(float* labels /* [B,L] in place */, int B, int64_t L, int max_pos, int max_neg,
                             uint64_t seed, void* workspace, size_t workspace_bytes, void* stream);
/* same with a DEVICE draw counter mixed into the seed and incremented by the call (a fresh sample per CUDA-graph replay) */
int tf_detloss_sample_device_ctr(float* labels, int B, int64_t L, int max_pos, int max_neg, uint64_t seed,
                                 uint64_t* counter_dev, void* workspace, size_t workspace_bytes, void* stream);

/* ---- tcgen05 implicit-GEMM convolutions (NHWC fp32 storage, TF32 tensor-core math, fp32 accumulate): the
 * building block behind every nn.Conv2d of tinyfaces/models/model.py:90-106 and its autograd backward.
 * w_packed: [Cout][k*k][Cin]; stride 1, "same" padding; x_lo/w_lo: optional low halves for the 3xTF32 mode. */
int tf_conv2d_nhwc(const float* x, const float* x_lo, int B, int H, int W, int Cin, const float* w_packed,
                   const float* w_lo, int Cout, int ksize, const float* bias, float* y, void* stream);
/* 1x1 stride-1 with the residual epilogue: y = conv(x, w) + (res_mask bit ? res : 0); res_mask NULL = unmasked.  (The
 * backward's "dx = conv1 dgrad + [out > 0] * dout" and the inference conv3 + shortcut.) */
int tf_conv2d_nhwc_res(const float* x, int B, int H, int W, int Cin, const float* w_packed, int Cout, const float* res,
                       const uint32_t* res_mask, float* y, void* stream);
int tf_conv2d_wgrad_nhwc(const float* x, const float* dy, int B, int H, int W, int Cin, int Cout, int ksize,
                         float* dw_packed /* accumulated */, void* stream);
/* stride 1 or 2 (padding ksize/2): y / dy are [B, ceil(H/s), ceil(W/s), Cout]; the stride is a TMA traversal stride */
int tf_conv2d_nhwc_strided(const float* x, int B, int H, int W, int Cin, const float* w_packed, int Cout, int ksize,
                           int stride, const float* bias, float* y, void* stream);
int tf_conv2d_wgrad_nhwc_strided(const float* x, const float* dy, int B, int H, int W, int Cin, int Cout, int ksize,
                                 int stride, float* dw_packed /* accumulated */, void* stream);
/* input gradient of a stride-2 convolution (3x3 pad 1 or 1x1) as 4 parity-class GEMMs, no zero insertion (the dgrad
 * autograd runs for the stride-2 Bottlenecks, torchvision resnet.py:134,197).  dy: [B, ceil(H/2), ceil(W/2), Cdy];
 * w_packed: [Cdx][k*k (flipped taps)][Cdy]; dx: [B,H,W,Cdx], overwritten (3x3) or -- accumulate != 0 -- added into;
 * a 1x1 touches the even pixels only, so the caller supplies the base values of dx (accumulate) or zeros. */
int tf_conv2d_dgrad_s2_nhwc(const float* dy, int B, int H, int W, int Cdy, const float* w_packed, int Cdx, int ksize,
                            int accumulate, float* dx, void* stream);

/* host-only views of the GEMM launch planning (no CUDA calls; used by the CPU tests):
 * tf_conv_plan -> out16 = {spatial, tw, th, tiles_x, tiles_y, m_tiles, bn, n_tiles, grid, two_cta, main_tiles, ksplit,
 *                          kiters, tail_pix0 & 0x7fffffff (or -1), tail_pix0 >> 31 (or -1), 0};
 * tf_dgrad_s2_taps -> number of taps (0..4) of parity class (py, px) and their (weight tap, dY x offset, dY y offset). */
int tf_conv_plan(int B, int H, int W, int Cin, int Cout, int ksize, int stride, int nseg, int plain, int has_res,
                 int num_sms, int* out16_host);
int tf_dgrad_s2_taps(int ksize, int py, int px, int* tap_w_host, int* tap_ox_host, int* tap_oy_host);

/* ---- the elementwise kernels between the GEMMs, one entry point each (NHWC fp32, tensors are [pixels, C] with C a power
 * of two in [64, 1024]); the whole-model executor below launches exactly these kernels.
 *   tf_bn_train_fwd / tf_bn_eval_fwd / tf_bn_bwd : nn.BatchNorm2d (+ shortcut add + ReLU) of torchvision resnet.py:143-163
 *     and its autograd backward.  relu_mask: 1 bit per element ((M*C+31)/32 words), bit = [out > 0] (the ReLU derivative the
 *     backward uses -- 0 at exactly 0, like ATen's threshold_backward).  round_tf32 != 0 rounds the result to TF32 (what the
 *     executor's `fast` mode stores for the next GEMM); 0 keeps fp32.
 *   tf_maxpool_fwd / _bwd : nn.MaxPool2d(3, 2, 1) (resnet.py:197-204); ties go to the first maximum in window scan order.
 *   tf_head_upsample_add_fwd / _bwd : score_res3 + crop(score4_upsample(score_res4)) of tinyfaces/models/model.py:104-126
 *     (ConvTranspose2d k4 s2 p1 with the frozen diagonal bilinear weight up_w [Cn,Cn,4,4]); s3/s4 NHWC with Cp >= Cn
 *     channels per pixel, out / dout NCHW [B,Cn,H3,W3]. */
int tf_bn_workspace_bytes(size_t* bytes_host);
int tf_bn_train_fwd(const float* y, int64_t M, int C, const float* gamma, const float* beta, float eps, float momentum,
                    float* run_mean, float* run_var, const float* res, int relu, int round_tf32, float* out,
                    uint32_t* relu_mask, float* save_mean, float* save_rstd, void* workspace, size_t workspace_bytes,
                    void* stream);
int tf_bn_eval_fwd(const float* y, int64_t M, int C, const float* gamma, const float* beta, const float* run_mean,
                   const float* run_var, float eps, const float* res, int relu, int round_tf32, float* out, void* workspace,
                   size_t workspace_bytes, void* stream);
int tf_bn_bwd(const float* dout, const uint32_t* relu_mask, const float* y, const float* save_mean, const float* save_rstd,
              const float* gamma, int64_t M, int C, float* dgamma, float* dbeta, float* dy, float* g_out, int round_tf32,
              void* workspace, size_t workspace_bytes, void* stream);
int tf_maxpool_fwd(const float* x, int B, int H, int W, int C, float* out, uint8_t* argmax, void* stream);
int tf_maxpool_bwd(const uint8_t* argmax, const float* dout, int B, int H, int W, int C, float* dx, void* stream);
int tf_head_workspace_bytes(int Cn, size_t* bytes_host);
int tf_head_upsample_add_fwd(const float* s3, const float* s4, const float* up_w, int B, int H3, int W3, int H4, int W4,
                             int Cn, int Cp, float* out_nchw, void* workspace, size_t workspace_bytes, void* stream);
int tf_head_upsample_add_bwd(const float* dout_nchw, const float* up_w, int B, int H3, int W3, int H4, int W4, int Cn, int Cp,
                             float* ds3, float* ds4, void* workspace, size_t workspace_bytes, void* stream);

/* ---- one image-pyramid level with the exact arithmetic of tinyfaces/evaluation.py:40-50 (to_pil_image, PIL bilinear
 * resize, ToTensor, Normalize).  img: float32 [3,H,W] in [0,1]; the int32 tables hold Pillow's fixed-point resampling
 * coefficients (bounds [out,2], taps [out,ksize]; NULL when that axis keeps its size); out: float32 [3,Ho,Wo]. */
int tf_pyramid_workspace_bytes(int H, int W, int Wo, size_t* bytes_host);
int tf_pyramid_level(const float* img, int H, int W, int Ho, int Wo, const int* bounds_h, const int* kk_h, int ksize_h,
                     const int* bounds_v, const int* kk_v, int ksize_v, const float* mean_host, const float* std_host,
                     float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- whole-model executor: replaces DetectionModel.forward (tinyfaces/models/model.py:89-128) and the backward
 * autograd derives from it (tinyfaces/trainer.py:86).  params / grads: HOST arrays of tf_model_num_params()
 * device pointers in tf_model_param_name() order (reference state_dict names; OIHW weights, BN vectors).
 * mode: 1 = fast (1xTF32), 2 = parity (3xTF32), 3 = mixed (the forward of mode 2, the backward GEMMs with single TF32
 * products on the hi parts of the saved operands).  x [B,3,H,W] and out [B,5T,H/8,W/8] are NCHW fp32. */
int tf_model_create(int num_templates, void** handle_host);
int tf_model_destroy(void* handle);
int tf_model_num_params(void* handle);
const char* tf_model_param_name(void* handle, int index);
int tf_model_output_shape(void* handle, int H, int W, int* H3_host, int* W3_host);
int tf_model_workspace_bytes(void* handle, int B, int H, int W, int training, int mode, size_t* bytes_host);
int tf_model_forward(void* handle, const float* x, int B, int H, int W, const void* const* params_host, int training,
                     int mode, float bn_momentum, float* out, void* workspace, size_t workspace_bytes, void* stream);
int tf_model_backward(void* handle, const float* dout, void* const* grads_host, void* stream);
/* tf_model_backward + bucket-completion events for an overlapped gradient all-reduce (the reference has no
 * multi-GPU path, trainer.py:83-87 is the single-device step this extends): events_host[k] (a caller-owned cudaEvent_t) is
 * recorded as soon as every gradient of the residual blocks with forward index >= event_first_block_host[k] has been
 * enqueued (blocks 0..29 = layer1.0 .. layer3.22; score_res4 counts as block 29, score_res3 as block 7, the stem as
 * block -1; first blocks descending).  The events live on an internal stream: wait on them, do not query the caller's. */
int tf_model_backward_ex(void* handle, const float* dout, void* const* grads_host, int num_events, void* const* events_host,
                         const int* event_first_block_host, void* stream);
int tf_model_get_tensor(void* handle, const char* name, float* dst, int64_t capacity, int* shape4_host, void* stream);
int tf_model_upsample_offdiag(void* handle, float* value_host, void* stream);

/* ---- training-target generation: replaces DataProcessor.get_heatmaps + get_regression (tinyfaces/datasets/processor.py:157-277)
 * and compute_dense_overlap (tinyfaces/datasets/dense_overlap.py:4-75): the dense template x ground-truth IoU volume (float64,
 * reference operation order, np.around(., 14)), the per-cell best object -> (tx, ty, tw, th), the per-object best cell and
 * the {-1, 0, +1} label rule with the padding-border exception.  bboxes: DEVICE [ng,4] float64, already filtered
 * (x2 > x1, y2 > y1); templates_host: HOST [nt,4]; jitter: DEVICE [vsy,vsx,nt,ng] float64 = the np.random.rand draws of
 * processor.py:203 (NULL: device noise from `seed`); pad_mask: DEVICE uint8 [vsy,vsx,nt] or NULL.  Outputs DEVICE float64:
 * class_maps [vsy,vsx,nt], regress_maps [vsy,vsx,4nt] (tx.. ty.. tw.. th..), iou_out [vsy,vsx,nt,ng] (optional, perturbed). */
int tf_targets_workspace_bytes(int vsy, int vsx, int nt, int ng, size_t* bytes_host);
int tf_heatmap_targets(const double* bboxes, int ng, const double* templates_host, int nt, int vsy, int vsx, int ofy, int ofx,
                       int sty, int stx, double pos_thresh, double neg_thresh, const double* jitter, uint64_t seed,
                       const uint8_t* pad_mask, double* class_maps, double* regress_maps, double* iou_out, void* workspace,
                       size_t workspace_bytes, void* stream);

/* ---- optimizer step: replaces optimizer.step() + scheduler.step() of tinyfaces/main.py:67-70,81-83 (torch.optim.SGD with
 * momentum and weight decay over the parameter groups of models/model.py:67-87, StepLR) on FLAT fp32 buffers, one launch per
 * gradient bucket, stream-ordered behind that bucket's all-reduce:
 *   g = grad_scale*grad + wd*p;  buf = momentum*buf + g;  p -= lr * (*lr_scale_dev) * buf      (dampening 0, no Nesterov)
 * n % 4 == 0; segment s = [seg_begin_host[s], seg_begin_host[s+1]) with its own base lr / weight decay (<= 8 segments, begins
 * multiples of 4); lr_scale_dev (optional DEVICE float) is the StepLR factor, written by tf_steplr_update:
 *   *epoch_dev += advance;  *lr_scale_dev = gamma ^ (*epoch_dev / step_size). */
int tf_sgd_step(float* params, const float* grads, float* momentum_buf, int64_t n, int num_segments,
                const int64_t* seg_begin_host, const float* seg_lr_host, const float* seg_wd_host, float momentum,
                float grad_scale, const float* lr_scale_dev, void* stream);
int tf_steplr_update(float* lr_scale_dev, int64_t* epoch_dev, int step_size, float gamma, int advance, void* stream);

#ifdef __cplusplus
}
#endif
#endif
