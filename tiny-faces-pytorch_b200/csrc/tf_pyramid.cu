// GPU image-pyramid level: the exact arithmetic of the reference's CPU preprocessing for one scale,
//   to_pil_image (x*255 -> uint8, truncation) -> PIL bilinear resize (antialiased two-pass resampling with
//   22-bit fixed-point coefficients and a uint8 intermediate) -> ToTensor (/255) -> Normalize ((x-mean)/std),
// i.e. /root/reference/tinyfaces/evaluation.py:40-50.  The coefficient tables are computed
// on the host exactly like Pillow's precompute_coeffs / normalize_coeffs_8bpc (tinyfaces_b200/pyramid.py); the
// kernels below reproduce ImagingResampleHorizontal_8bpc / Vertical_8bpc, so the result is bit-identical.
#include "tf_common.cuh"
#include <algorithm>

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;

__device__ __forceinline__ unsigned char clip8(int v) {
    v >>= PRECISION_BITS;
    return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// float [3,H,W] in [0,1] -> uint8 (pic.mul(255).byte(): truncation toward zero)
__global__ void quantize_kernel(const float* __restrict__ img, long long n, unsigned char* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (unsigned char)(int)(img[i] * 255.0f);
}
// horizontal pass: in [3,H,W] -> out [3,H,Wo]
__global__ void resample_h_kernel(const unsigned char* __restrict__ in, int H, int W, int Wo, const int* __restrict__ bounds,
                                  const int* __restrict__ kk, int ksize, unsigned char* __restrict__ out) {
    const long long total = 3LL * H * Wo;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int xx = (int)(t % Wo);
        const long long row = t / Wo;                      // c*H + y
        const int xmin = bounds[2 * xx], cnt = bounds[2 * xx + 1];
        const int* k = kk + (size_t)xx * ksize;
        const unsigned char* src = in + row * W + xmin;
        int ss = 1 << (PRECISION_BITS - 1);
        for (int x = 0; x < cnt; ++x) ss += (int)src[x] * k[x];
        out[t] = clip8(ss);
    }
}
// vertical pass + ToTensor + Normalize: in [3,H,Wo] -> out float [3,Ho,Wo]
__global__ void resample_v_norm_kernel(const unsigned char* __restrict__ in, int H, int Wo, int Ho, const int* __restrict__ bounds,
                                       const int* __restrict__ kk, int ksize, int resample, float m0, float m1, float m2,
                                       float s0, float s1, float s2, float* __restrict__ out) {
    const long long total = 3LL * Ho * Wo;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(t % Wo);
        const int yy = (int)((t / Wo) % Ho);
        const int c = (int)(t / ((long long)Wo * Ho));
        unsigned char v;
        if (resample) {
            const int ymin = bounds[2 * yy], cnt = bounds[2 * yy + 1];
            const int* k = kk + (size_t)yy * ksize;
            const unsigned char* src = in + ((long long)c * H + ymin) * Wo + x;
            int ss = 1 << (PRECISION_BITS - 1);
            for (int y = 0; y < cnt; ++y) ss += (int)src[(long long)y * Wo] * k[y];
            v = clip8(ss);
        } else {
            v = in[((long long)c * H + yy) * Wo + x];
        }
        const float f = __fdiv_rn((float)v, 255.0f);                                   // ToTensor
        const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
        out[t] = __fdiv_rn(__fsub_rn(f, mean), sd);                                    // Normalize
    }
}

}  // namespace

TF_API int tf_pyramid_workspace_bytes(int H, int W, int Wo, size_t* bytes) {
    TF_REQUIRE(bytes && H > 0 && W > 0 && Wo > 0, "tf_pyramid_workspace_bytes: bad args");
    *bytes = tf_align_up((size_t)3 * H * W, 256) + tf_align_up((size_t)3 * H * Wo, 256) + 512;
    return TF_OK;
}

// img: float32 [3,H,W] in [0,1].  bounds_*: int32 [out,2] (first source index, tap count); kk_*: int32 [out, ksize_*]
// fixed-point taps (device pointers; null tables = that axis keeps its size).  out: float32 [3,Ho,Wo], normalised.
TF_API int tf_pyramid_level(const float* img, int H, int W, int Ho, int Wo, const int* bounds_h, const int* kk_h, int ksize_h,
                            const int* bounds_v, const int* kk_v, int ksize_v, const float* mean_host,
                            const float* std_host, float* out, void* workspace, size_t workspace_bytes, void* stream) {
    TF_REQUIRE(img && out && mean_host && std_host && workspace, "tf_pyramid_level: null pointer");
    TF_REQUIRE(H > 0 && W > 0 && Ho > 0 && Wo > 0, "tf_pyramid_level: bad shape");
    TF_REQUIRE((bounds_h != nullptr) == (kk_h != nullptr) && (bounds_v != nullptr) == (kk_v != nullptr), "tf_pyramid_level: tables");
    TF_REQUIRE(bounds_h || Wo == W, "tf_pyramid_level: width changes but no horizontal table");
    TF_REQUIRE(bounds_v || Ho == H, "tf_pyramid_level: height changes but no vertical table");
    size_t need;
    tf_pyramid_workspace_bytes(H, W, Wo, &need);
    if (workspace_bytes < need) { tf_set_error("tf_pyramid_level: workspace %zu < %zu", workspace_bytes, need); return TF_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* q = reinterpret_cast<unsigned char*>(workspace);
    unsigned char* tmp = q + tf_align_up((size_t)3 * H * W, 256);
    const long long n = 3LL * H * W;
    quantize_kernel<<<(int)std::min<long long>((n + 255) / 256, 148 * 16), 256, 0, st>>>(img, n, q);
    const unsigned char* hsrc = q;
    if (bounds_h) {
        const long long t = 3LL * H * Wo;
        resample_h_kernel<<<(int)std::min<long long>((t + 255) / 256, 148 * 16), 256, 0, st>>>(q, H, W, Wo, bounds_h, kk_h, ksize_h, tmp);
        hsrc = tmp;
    }
    const long long t = 3LL * Ho * Wo;
    resample_v_norm_kernel<<<(int)std::min<long long>((t + 255) / 256, 148 * 16), 256, 0, st>>>(
        hsrc, H, Wo, Ho, bounds_v, kk_v, ksize_v, bounds_v ? 1 : 0, mean_host[0], mean_host[1], mean_host[2], std_host[0],
        std_host[1], std_host[2], out);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
