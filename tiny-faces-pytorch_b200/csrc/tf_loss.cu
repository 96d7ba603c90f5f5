// Detection loss: OHEM + (device) balance sampling + ONE fused forward/backward kernel.
// Replaces /root/reference/tinyfaces/models/loss.py:59-93 (DetectionCriterion.forward) and, for the
// device sampler, /root/reference/tinyfaces/models/utils.py:103-163.
//
// Layout: NCHW exactly as the caller holds it -- output [B,5T,H,W], class_map [B,T,H,W],
// regression_map [B,4T,H,W]; threads run along H*W so every access is coalesced.
// HBM-bound: 1500 B/pixel algorithmic (read 125+25+100 fp32, write 125 fp32).
#include "tf_common.cuh"
#include <algorithm>

namespace {

__device__ __forceinline__ float soft_margin(float x, float y) {
    return log1pf(expf(-x * y));                 // nn.SoftMarginLoss: log(1 + exp(-y x))
}

// loss.py:59-63 -- class_map[softmargin(cls, class_map) < thresh] = 0, IN PLACE
__global__ void ohem_kernel(const float* __restrict__ out, float* __restrict__ class_map, int T, long long HW,
                            long long total, float thresh) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long p = i % HW;
        const long long bt = i / HW;
        const long long b = bt / T, t = bt % T;
        const float y = class_map[i];
        const float x = out[(b * 5 * T + t) * HW + p];
        if (soft_margin(x, y) < thresh) class_map[i] = 0.0f;
    }
}

// loss.py:74-88 + closed-form gradient; sums[0] += sum(masked cls loss), sums[1] += sum(masked reg loss)
__global__ void __launch_bounds__(256) loss_fwd_bwd_kernel(const float* __restrict__ out,
                                                           const float* __restrict__ labels,
                                                           const float* __restrict__ regmap, int T, long long HW,
                                                           long long total, float reg_weight,
                                                           float* __restrict__ grad, double* __restrict__ sums) {
    float acc_c = 0.f, acc_r = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long p = i % HW;
        const long long bt = i / HW;
        const long long b = bt / T, t = bt % T;
        const float y = labels[i];
        const long long ob = (b * 5 * T + t) * HW + p;
        const float x = out[ob];
        float g = 0.f;
        if (y != 0.f) {
            acc_c += soft_margin(x, y);
            g = -y / (1.0f + expf(y * x));       // d/dx log(1+e^{-yx})
        }
        grad[ob] = g;
        const bool pos = y > 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long oi = ob + (long long)(k + 1) * T * HW;
            float gr = 0.f;
            if (pos) {
                const float d = out[oi] - regmap[((b * 4 + k) * T + t) * HW + p];
                const float a = fabsf(d);
                if (a < 1.0f) { acc_r += 0.5f * d * d; gr = d; }
                else          { acc_r += a - 0.5f;     gr = d > 0.f ? 1.0f : -1.0f; }
                gr *= reg_weight;
            }
            grad[oi] = gr;
        }
    }
    __shared__ double s_c[8], s_r[8];
    double wc = tf_warp_sum((double)acc_c), wr = tf_warp_sum((double)acc_r);
    if ((threadIdx.x & 31) == 0) { s_c[threadIdx.x >> 5] = wc; s_r[threadIdx.x >> 5] = wr; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0, r = 0;
        for (int i = 0; i < 8; ++i) { c += s_c[i]; r += s_r[i]; }
        atomicAdd(&sums[0], c);
        atomicAdd(&sums[1], r);
    }
}

// ---------------------------------------------------------------------------------------------
// Device balance sampler (statistically equivalent to utils.py:103-139; the bit-exact variant
// that consumes np.random lives on the host, see tinyfaces_b200/models/utils.py).
// Every label gets a 32-bit hash key; per image and per class the `limit` smallest keys survive.
// Two 16-bit radix-select passes find the threshold key; ties at the threshold are broken by ticket.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int hash_key(unsigned long long seed, unsigned long long idx) {
    unsigned long long z = seed + idx * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (unsigned int)(z >> 32);
}
struct SelState { unsigned int total, bin1, before1, bin2, before2, keep_all; };

__device__ __forceinline__ int label_class(float y) { return y > 0.f ? 0 : (y < 0.f ? 1 : -1); }

__global__ void samp_hist1_kernel(const float* __restrict__ labels, long long L, unsigned long long seed,
                                  const unsigned long long* __restrict__ ctr, unsigned int* __restrict__ hist1) {
    const int b = blockIdx.y;
    if (ctr) seed += *ctr * 0xD1342543DE82EF95ull;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < L; i += (long long)gridDim.x * blockDim.x) {
        const int c = label_class(labels[b * L + i]);
        if (c < 0) continue;
        atomicAdd(&hist1[((size_t)(b * 2 + c) << 16) + (hash_key(seed, b * L + i) >> 16)], 1u);
    }
}
// one CTA per (image, class): find the bin where the running count reaches `limit`
__global__ void __launch_bounds__(1024) samp_find_kernel(const unsigned int* __restrict__ hist, SelState* __restrict__ st,
                                                         unsigned int limit_pos, unsigned int limit_neg, int pass) {
    const int bc = blockIdx.x;
    const unsigned int limit = (bc & 1) ? limit_neg : limit_pos;
    const unsigned int* h = hist + ((size_t)bc << 16);
    __shared__ unsigned int part[1024];
    unsigned int s = 0;
    for (int k = 0; k < 64; ++k) s += h[threadIdx.x * 64 + k];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        SelState& S = st[bc];
        unsigned int need;
        if (pass == 0) {
            unsigned int tot = 0;
            for (int i = 0; i < 1024; ++i) tot += part[i];
            S.total = tot; S.keep_all = tot <= limit; need = limit;
        } else need = limit - S.before1;
        if (!S.keep_all) {
            unsigned int cum = 0; int seg = 0;
            while (seg < 1024 && cum + part[seg] < need) { cum += part[seg]; ++seg; }
            int bin = seg * 64;
            while (bin < 65536 && cum + h[bin] < need) { cum += h[bin]; ++bin; }
            if (pass == 0) { S.bin1 = bin; S.before1 = cum; } else { S.bin2 = bin; S.before2 = cum; }
        }
    }
}
__global__ void samp_hist2_kernel(const float* __restrict__ labels, long long L, unsigned long long seed,
                                  const unsigned long long* __restrict__ ctr, const SelState* __restrict__ st,
                                  unsigned int* __restrict__ hist2) {
    const int b = blockIdx.y;
    if (ctr) seed += *ctr * 0xD1342543DE82EF95ull;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < L; i += (long long)gridDim.x * blockDim.x) {
        const int c = label_class(labels[b * L + i]);
        if (c < 0) continue;
        const SelState S = st[b * 2 + c];
        if (S.keep_all) continue;
        const unsigned int key = hash_key(seed, b * L + i);
        if ((key >> 16) == S.bin1) atomicAdd(&hist2[((size_t)(b * 2 + c) << 16) + (key & 0xffffu)], 1u);
    }
}
__global__ void samp_apply_kernel(float* __restrict__ labels, long long L, unsigned long long seed,
                                  const unsigned long long* __restrict__ ctr, const SelState* __restrict__ st,
                                  unsigned int* __restrict__ ties, unsigned int limit_pos, unsigned int limit_neg) {
    const int b = blockIdx.y;
    if (ctr) seed += *ctr * 0xD1342543DE82EF95ull;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < L; i += (long long)gridDim.x * blockDim.x) {
        const int c = label_class(labels[b * L + i]);
        if (c < 0) continue;
        const SelState S = st[b * 2 + c];
        if (S.keep_all) continue;
        const unsigned int key = hash_key(seed, b * L + i);
        const unsigned int thr = (S.bin1 << 16) | S.bin2;
        bool keep = key < thr;
        if (key == thr) {
            const unsigned int need = (c ? limit_neg : limit_pos) - S.before1 - S.before2;
            keep = atomicAdd(&ties[b * 2 + c], 1u) < need;
        }
        if (!keep) labels[b * L + i] = 0.0f;
    }
}

__global__ void samp_bump_kernel(unsigned long long* ctr) { *ctr += 1; }

}  // namespace

TF_API int tf_detloss_ohem(const float* output, float* class_map, int B, int T, int64_t HW, float thresh, void* stream) {
    TF_REQUIRE(output && class_map && B > 0 && T > 0 && HW > 0, "tf_detloss_ohem: bad args");
    const long long total = (long long)B * T * HW;
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    ohem_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(output, class_map, T, HW, total, thresh);
    TF_LAUNCH_CHECK();
    return TF_OK;
}

// sums: device double[2], ACCUMULATED into (caller zeroes): {sum masked cls loss, sum masked reg loss (unweighted)}
TF_API int tf_detloss_fwd_bwd(const float* output, const float* labels, const float* regression_map, int B, int T,
                              int64_t HW, float reg_weight, float* grad_output, double* sums, void* stream) {
    TF_REQUIRE(output && labels && regression_map && grad_output && sums && B > 0 && T > 0 && HW > 0,
               "tf_detloss_fwd_bwd: bad args");
    const long long total = (long long)B * T * HW;
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 8);
    loss_fwd_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(output, labels, regression_map, T, HW, total,
                                                                  reg_weight, grad_output, sums);
    TF_LAUNCH_CHECK();
    return TF_OK;
}

TF_API int tf_detloss_sample_device_ctr(float* labels, int B, int64_t L, int max_pos, int max_neg, uint64_t seed,
                                        uint64_t* counter_dev, void* workspace, size_t workspace_bytes, void* stream);

TF_API int tf_detloss_sample_workspace_bytes(int B, size_t* bytes) {
    TF_REQUIRE(bytes && B > 0, "tf_detloss_sample_workspace_bytes: bad args");
    *bytes = (size_t)B * 2 * 65536 * 4 * 2 + (size_t)B * 2 * (sizeof(SelState) + 4) + 1024;
    return TF_OK;
}

// labels [B, L] fp32 in {-1,0,+1}, IN PLACE: at most max_pos positives and max_neg negatives survive per image.
TF_API int tf_detloss_sample_device(float* labels, int B, int64_t L, int max_pos, int max_neg, uint64_t seed,
                                    void* workspace, size_t workspace_bytes, void* stream) {
    return tf_detloss_sample_device_ctr(labels, B, L, max_pos, max_neg, seed, nullptr, workspace, workspace_bytes, stream);
}
// Same, with a DEVICE draw counter: the hash seed is seed + f(*counter_dev) and the call increments the counter, so a
// captured CUDA graph draws a fresh sample on every replay (a by-value seed would be frozen into the graph).
TF_API int tf_detloss_sample_device_ctr(float* labels, int B, int64_t L, int max_pos, int max_neg, uint64_t seed,
                                        uint64_t* counter_dev, void* workspace, size_t workspace_bytes, void* stream) {
    TF_REQUIRE(labels && workspace && B > 0 && L > 0 && max_pos >= 0 && max_neg >= 0, "tf_detloss_sample_device: bad args");
    size_t need;
    tf_detloss_sample_workspace_bytes(B, &need);
    if (workspace_bytes < need) { tf_set_error("tf_detloss_sample_device: workspace %zu < %zu", workspace_bytes, need); return TF_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    TfArena ar(workspace, workspace_bytes);
    unsigned int* hist1 = ar.take<unsigned int>((size_t)B * 2 * 65536);
    unsigned int* hist2 = ar.take<unsigned int>((size_t)B * 2 * 65536);
    SelState* sel = ar.take<SelState>(B * 2);
    unsigned int* ties = ar.take<unsigned int>(B * 2);
    const unsigned long long* ctr = reinterpret_cast<const unsigned long long*>(counter_dev);
    TF_CHECK_CUDA(cudaMemsetAsync(workspace, 0, need, st));
    dim3 grid((unsigned)std::min<long long>((L + 255) / 256, 592), B);
    samp_hist1_kernel<<<grid, 256, 0, st>>>(labels, L, seed, ctr, hist1);
    samp_find_kernel<<<B * 2, 1024, 0, st>>>(hist1, sel, max_pos, max_neg, 0);
    samp_hist2_kernel<<<grid, 256, 0, st>>>(labels, L, seed, ctr, sel, hist2);
    samp_find_kernel<<<B * 2, 1024, 0, st>>>(hist2, sel, max_pos, max_neg, 1);
    samp_apply_kernel<<<grid, 256, 0, st>>>(labels, L, seed, ctr, sel, ties, max_pos, max_neg);
    if (counter_dev) samp_bump_kernel<<<1, 1, 0, st>>>(reinterpret_cast<unsigned long long*>(counter_dev));
    TF_LAUNCH_CHECK();
    return TF_OK;
}
