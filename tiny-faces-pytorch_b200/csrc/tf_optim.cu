// SGD (momentum, weight decay) + StepLR on flat fp32 buffers: the optimizer.step() / scheduler.step() of
// /root/reference/tinyfaces/main.py:67-70,81-83 (torch.optim.SGD(momentum 0.9, weight_decay 5e-4) over the four
// parameter groups of models/model.py:67-87, StepLR(step_size 20, gamma 0.1)) as ONE launch per gradient bucket, meant to
// run on the communication stream right behind that bucket's all-reduce -- the optimizer is the all-reduce's epilogue and
// overlaps the rest of the backward.
//
//   g = grad_scale * grad + wd * p;   buf = momentum * buf + g;   p -= lr * lr_scale * buf
//
// which is torch.optim.SGD's update (dampening 0, no Nesterov; a zero-initialised buf reproduces its first step, where
// buf := g).  lr_scale is read from DEVICE memory (the StepLR factor gamma^(epoch / step_size)) so that a captured CUDA
// graph follows the schedule without re-capture; tf_steplr_update writes it.
#include "tf_common.cuh"
#include <algorithm>

namespace {

constexpr int MAX_SEG = 8;
struct SgdSegs {
    long long begin[MAX_SEG + 1];        // element offsets (multiples of 4), begin[num] = n
    float lr[MAX_SEG], wd[MAX_SEG];
    int num;
};

__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                  long long n4, SgdSegs segs, float momentum, float grad_scale,
                                                  const float* __restrict__ lr_scale_dev) {
    const float lrs = lr_scale_dev ? *lr_scale_dev : 1.f;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) {
        const long long i = t * 4;
        int s = 0;
#pragma unroll
        for (int k = 1; k < MAX_SEG; ++k) if (k < segs.num && i >= segs.begin[k]) s = k;
        const float lr = segs.lr[s] * lrs, wd = segs.wd[s];
        float4 pv = *reinterpret_cast<const float4*>(p + i);
        const float4 gv = *reinterpret_cast<const float4*>(g + i);
        float4 mv = *reinterpret_cast<const float4*>(m + i);
        mv.x = momentum * mv.x + (grad_scale * gv.x + wd * pv.x); pv.x -= lr * mv.x;
        mv.y = momentum * mv.y + (grad_scale * gv.y + wd * pv.y); pv.y -= lr * mv.y;
        mv.z = momentum * mv.z + (grad_scale * gv.z + wd * pv.z); pv.z -= lr * mv.z;
        mv.w = momentum * mv.w + (grad_scale * gv.w + wd * pv.w); pv.w -= lr * mv.w;
        *reinterpret_cast<float4*>(m + i) = mv;
        *reinterpret_cast<float4*>(p + i) = pv;
    }
}
__global__ void steplr_kernel(float* lr_scale, long long* epoch, int step_size, float gamma, int advance) {
    if (advance) *epoch += advance;
    *lr_scale = powf(gamma, (float)(*epoch / step_size));
}

}  // namespace

// params / grads / momentum_buf: flat fp32 device buffers of n elements (n % 4 == 0, 16-byte aligned); segment s covers
// [seg_begin_host[s], seg_begin_host[s+1]) (the last one ends at n) with base learning rate seg_lr_host[s] and weight decay
// seg_wd_host[s]; lr_scale_dev: optional device float multiplied into every learning rate (NULL = 1).
TF_API int tf_sgd_step(float* params, const float* grads, float* momentum_buf, int64_t n, int num_segments,
                       const int64_t* seg_begin_host, const float* seg_lr_host, const float* seg_wd_host, float momentum,
                       float grad_scale, const float* lr_scale_dev, void* stream) {
    TF_REQUIRE(params && grads && momentum_buf && n >= 0 && n % 4 == 0, "tf_sgd_step: bad buffers (n must be a multiple of 4)");
    TF_REQUIRE(num_segments >= 1 && num_segments <= MAX_SEG && seg_begin_host && seg_lr_host && seg_wd_host, "tf_sgd_step: 1..8 segments");
    TF_REQUIRE((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(momentum_buf)) % 16 == 0,
               "tf_sgd_step: buffers must be 16-byte aligned");
    if (n == 0) return TF_OK;
    SgdSegs segs;
    memset(&segs, 0, sizeof(segs));
    segs.num = num_segments;
    for (int s = 0; s < num_segments; ++s) {
        TF_REQUIRE(seg_begin_host[s] % 4 == 0 && seg_begin_host[s] >= 0 && seg_begin_host[s] <= n && (s == 0 ? seg_begin_host[0] == 0 : seg_begin_host[s] >= seg_begin_host[s - 1]),
                   "tf_sgd_step: segment begins must be ascending multiples of 4 starting at 0");
        segs.begin[s] = seg_begin_host[s]; segs.lr[s] = seg_lr_host[s]; segs.wd[s] = seg_wd_host[s];
    }
    segs.begin[num_segments] = n;
    const long long n4 = n / 4;
    const int blocks = (int)std::min<long long>((n4 + 255) / 256, 148ll * 8);
    sgd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, momentum_buf, n4, segs, momentum, grad_scale, lr_scale_dev);
    TF_LAUNCH_CHECK();
    return TF_OK;
}

// StepLR on the device: *epoch_dev += advance; *lr_scale_dev = gamma ^ (epoch / step_size)   (scheduler.step(), main.py:81-83)
TF_API int tf_steplr_update(float* lr_scale_dev, int64_t* epoch_dev, int step_size, float gamma, int advance, void* stream) {
    TF_REQUIRE(lr_scale_dev && epoch_dev && step_size > 0, "tf_steplr_update: bad args");
    steplr_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(lr_scale_dev, reinterpret_cast<long long*>(epoch_dev), step_size, gamma, advance);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
