// HBM-bound NHWC fp32 kernels around the convolution GEMMs: BatchNorm statistics / apply / backward,
// ReLU, residual add, max-pool, stem im2col, stride-2 subsample / zero-insert, weight (un)packing and the
// head combine (bilinear 2x upsample + crop + add + NCHW transpose).  They replace the elementwise part
// of /root/reference/tinyfaces/models/model.py:89-128 (torchvision Bottleneck.forward, resnet.py:143-163)
// and its autograd backward.  All tensors are [pixels, C] with C % 4 == 0; accesses are float4.
#include "tf_common.cuh"
#include "tf_elementwise.h"
#include "tf_conv_gemm.h"
#include <algorithm>
#include <cstdlib>

namespace {

constexpr int EW_THREADS = 256;
// Grid cap of the grid-stride kernels (CTAs of 256 threads per SM).  Measured on B200: capping lower to leave room
// for side-stream GEMM CTAs does not pay (50.1 ms/step at 32 vs 50.7 at 6).  TF_EW_BLOCKS_PER_SM overrides.
inline int ew_blocks_per_sm() {
    static int v = 0;
    if (!v) { const char* e = getenv("TF_EW_BLOCKS_PER_SM"); v = e ? atoi(e) : 32; if (v < 1) v = 1; }
    return v;
}
inline int ew_blocks(long long n, int per_thread = 1) {
    long long b = (n + (long long)EW_THREADS * per_thread - 1) / ((long long)EW_THREADS * per_thread);
    return (int)std::min<long long>(std::max<long long>(b, 1), 148LL * ew_blocks_per_sm());
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
// streaming read (evict-first): the big activation / gradient tensors are touched once per pass and must not push the
// per-channel coefficient vectors out of the (deliberately small, see prefer_shared_carveout) L1
__device__ __forceinline__ float4 ld4s(const float* p, int streaming = 1) {
    return streaming ? __ldcs(reinterpret_cast<const float4*>(p)) : *reinterpret_cast<const float4*>(p);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// hi half of the exact (hi, lo) TF32 split x = hi + lo: x ROUNDED to TF32 (not truncated), so that hi alone is an unbiased
// TF32 operand (precision "mixed" runs its backward GEMMs on the hi parts only) and |lo| <= 2^-12 |x|; x - hi is exact in fp32
__device__ __forceinline__ float hi_part(float x) { return tf_round_tf32(x); }

// ReLU masks are kept as 1 bit per element (element i -> bit i%32 of word i/32) so that the backward passes do not
// have to re-read the activation tensor: 4 consecutive elements (one float4) = one nibble.
__device__ __forceinline__ unsigned int mask_nibble(const unsigned int* __restrict__ mask, long long i) {
    return (mask[i >> 5] >> (unsigned)(i & 31)) & 0xFu;
}
__device__ __forceinline__ void apply_mask(float4& g, unsigned int nib) {
    if (!(nib & 1u)) g.x = 0.f;
    if (!(nib & 2u)) g.y = 0.f;
    if (!(nib & 4u)) g.z = 0.f;
    if (!(nib & 8u)) g.w = 0.f;
}

// write v either rounded to tf32 (fast mode), or split into (hi, lo) (parity mode), or unchanged
__device__ __forceinline__ void store_act(float* out, float* out_lo, long long i, float4 v, int mode) {
    if (mode == 1) {
        v.x = tf_round_tf32(v.x); v.y = tf_round_tf32(v.y); v.z = tf_round_tf32(v.z); v.w = tf_round_tf32(v.w);
        st4(out + i, v);
    } else if (mode == 2) {
        float4 h = make_float4(hi_part(v.x), hi_part(v.y), hi_part(v.z), hi_part(v.w));
        st4(out + i, h);
        st4(out_lo + i, make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w));
    } else {
        st4(out + i, v);
    }
}

// ------------------------------------------------------------------------------------------------
// per-channel column reductions over [M, C]: partial[blk][k][C], k = 0..1
//   mode 0: (sum y, sum y^2)                                  -- BN forward statistics
//   mode 1: (sum g, sum g*xhat), g = dout (* [act > 0])       -- BN backward
//   mode 2: (sum y, 0)                                        -- bias gradient
// ------------------------------------------------------------------------------------------------
// No shared memory on purpose: these kernels run next to side-stream weight-gradient GEMM CTAs that own all but ~1 KB of
// an SM's shared memory -- a block with even a few KB of static smem would wait for those CTAs to retire instead of
// overlapping them.  Thread t works on float4 channel group t / lanes of row lane t % lanes (lanes = 1024 / C rows per
// block, a power of two <= 16, so the row lanes of one channel group sit in adjacent lanes of one warp).
template <int MODE>
__global__ void __launch_bounds__(EW_THREADS) colreduce_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                               const float* __restrict__ act,
                                                               const unsigned int* __restrict__ mask,
                                                               const float* __restrict__ mean,
                                                               const float* __restrict__ rstd, long long M, int C,
                                                               float* __restrict__ partial, int slots, int descending,
                                                               int streaming) {
    const int groups = C / 4;                       // float4 channel groups (16 .. 256, a power of two)
    const int lanes = EW_THREADS / groups;          // row lanes per block (1 .. 16)
    const int g = threadIdx.x / lanes, rl = threadIdx.x % lanes;
    float4 s0 = make_float4(0, 0, 0, 0), s1 = s0;
    float4 mu = s0, rs = s0;
    if (MODE == 1) { mu = ld4(mean + g * 4); rs = ld4(rstd + g * 4); }
    {
        const long long step = (long long)gridDim.x * lanes;
        long long r = (long long)blockIdx.x * lanes + rl;
        if (MODE == 1 && mask) {
            // 4 rows per iteration with every load issued up front: the bytes in flight have to come from each thread
            for (; r + 3 * step < M; r += 4 * step) {
                float4 v[4], y[4]; unsigned int nib[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const long long row = descending ? M - 1 - (r + u * step) : r + u * step;
                    const long long i = row * C + g * 4;
                    v[u] = ld4s(a + i, streaming); y[u] = ld4s(b + i, streaming); nib[u] = mask_nibble(mask, i);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    apply_mask(v[u], nib[u]);
                    s0.x += v[u].x; s0.y += v[u].y; s0.z += v[u].z; s0.w += v[u].w;
                    s1.x += v[u].x * (y[u].x - mu.x) * rs.x; s1.y += v[u].y * (y[u].y - mu.y) * rs.y;
                    s1.z += v[u].z * (y[u].z - mu.z) * rs.z; s1.w += v[u].w * (y[u].w - mu.w) * rs.w;
                }
            }
        }
        for (; r < M; r += step) {
            const long long i = (descending ? M - 1 - r : r) * C + g * 4;
            float4 v = ld4(a + i);
            if (MODE == 0) {
                s0.x += v.x; s0.y += v.y; s0.z += v.z; s0.w += v.w;
                s1.x += v.x * v.x; s1.y += v.y * v.y; s1.z += v.z * v.z; s1.w += v.w * v.w;
            } else if (MODE == 1) {
                if (mask) apply_mask(v, mask_nibble(mask, i));
                else if (act) { float4 o = ld4(act + i); if (!(o.x > 0.f)) v.x = 0.f; if (!(o.y > 0.f)) v.y = 0.f; if (!(o.z > 0.f)) v.z = 0.f; if (!(o.w > 0.f)) v.w = 0.f; }
                float4 y = ld4(b + i);
                s0.x += v.x; s0.y += v.y; s0.z += v.z; s0.w += v.w;
                s1.x += v.x * (y.x - mu.x) * rs.x; s1.y += v.y * (y.y - mu.y) * rs.y;
                s1.z += v.z * (y.z - mu.z) * rs.z; s1.w += v.w * (y.w - mu.w) * rs.w;
            } else {
                s0.x += v.x; s0.y += v.y; s0.z += v.z; s0.w += v.w;
            }
        }
    }
    for (int o = lanes >> 1; o > 0; o >>= 1) {       // row lanes of a group are adjacent lanes of the warp
        s0.x += __shfl_xor_sync(0xffffffffu, s0.x, o); s0.y += __shfl_xor_sync(0xffffffffu, s0.y, o);
        s0.z += __shfl_xor_sync(0xffffffffu, s0.z, o); s0.w += __shfl_xor_sync(0xffffffffu, s0.w, o);
        s1.x += __shfl_xor_sync(0xffffffffu, s1.x, o); s1.y += __shfl_xor_sync(0xffffffffu, s1.y, o);
        s1.z += __shfl_xor_sync(0xffffffffu, s1.z, o); s1.w += __shfl_xor_sync(0xffffffffu, s1.w, o);
    }
    if (rl == 0) {
        if (slots > 0) {
            // BN backward: blocks accumulate into `slots` pre-zeroed rows (fire-and-forget fp32 reductions at L2), so
            // the finalize kernel behind this one reads a handful of rows -- one L2 round trip -- instead of one row
            // per block: with a weight-gradient GEMM streaming through HBM next to it that latency chain was ~35 us
            float* p = partial + (size_t)(blockIdx.x % slots) * 2 * C + g * 4;
            atomicAdd(p, s0.x); atomicAdd(p + 1, s0.y); atomicAdd(p + 2, s0.z); atomicAdd(p + 3, s0.w);
            atomicAdd(p + C, s1.x); atomicAdd(p + C + 1, s1.y); atomicAdd(p + C + 2, s1.z); atomicAdd(p + C + 3, s1.w);
        } else {
            float* p = partial + (size_t)blockIdx.x * 2 * C;
            st4(p + g * 4, s0);
            st4(p + C + g * 4, s1);
        }
    }
}

// Sum of the per-block partials of 32 channels with FIN_LANES row lanes per channel (fp64), valid for threadIdx.y == 0.
// 16 KB of static shared memory: small enough to sit next to a side-stream weight-gradient CTA (which leaves 32 KB).
constexpr int FIN_LANES = 32;
template <bool CLEAN>
__device__ __forceinline__ void reduce_partials(float* __restrict__ partial, int nblk, int C, int c, double& s, double& q) {
    __shared__ double sh[2][FIN_LANES][32];
    s = 0; q = 0;
    if (c < C) {
        // four independent rows per iteration: this is a latency chain through L2 (up to 592 rows / 32 lanes deep)
        double sa[4] = {0, 0, 0, 0}, qa[4] = {0, 0, 0, 0};
        int b = threadIdx.y;
        for (; b + 3 * FIN_LANES < nblk; b += 4 * FIN_LANES) {
            float v[4], w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { float* r = partial + (size_t)(b + u * FIN_LANES) * 2 * C + c; v[u] = r[0]; w[u] = r[C]; }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                sa[u] += v[u]; qa[u] += w[u];
                if (CLEAN) { float* r = partial + (size_t)(b + u * FIN_LANES) * 2 * C + c; r[0] = 0.f; r[C] = 0.f; }   // slot rows are left zeroed
            }
        }
        for (; b < nblk; b += FIN_LANES) {
            float* r = partial + (size_t)b * 2 * C + c;
            sa[0] += r[0]; qa[0] += r[C];
            if (CLEAN) { r[0] = 0.f; r[C] = 0.f; }
        }
        s = (sa[0] + sa[1]) + (sa[2] + sa[3]); q = (qa[0] + qa[1]) + (qa[2] + qa[3]);
    }
    sh[0][threadIdx.y][threadIdx.x] = s; sh[1][threadIdx.y][threadIdx.x] = q;
    __syncthreads();
    if (threadIdx.y == 0)
        for (int l = 1; l < FIN_LANES; ++l) { s += sh[0][l][threadIdx.x]; q += sh[1][l][threadIdx.x]; }
}
// BN forward finalize (training): batch mean / biased var -> scale, shift, saved mean / rstd, running stats
__global__ void __launch_bounds__(32 * FIN_LANES) bn_finalize_train_kernel(float* __restrict__ partial, int nblk, int C, long long M,
                                         const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                         float momentum, float* __restrict__ run_mean, float* __restrict__ run_var,
                                         float* __restrict__ scale, float* __restrict__ shift,
                                         float* __restrict__ save_mean, float* __restrict__ save_rstd) {
    const int c = blockIdx.x * 32 + threadIdx.x;
    double s, q;
    reduce_partials<false>(partial, nblk, C, c, s, q);
    if (threadIdx.y != 0 || c >= C) return;
    const double mean = s / (double)M;
    double var = q / (double)M - mean * mean;
    if (var < 0) var = 0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * rstd;
    scale[c] = sc;
    shift[c] = beta[c] - (float)mean * sc;
    save_mean[c] = (float)mean;
    save_rstd[c] = rstd;
    if (run_mean) {
        const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
        run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * (float)mean;
        run_var[c] = (1.f - momentum) * run_var[c] + momentum * (float)unbiased;
    }
}
// eval: scale / shift from the running statistics
__global__ void bn_finalize_eval_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                        const float* __restrict__ run_mean, const float* __restrict__ run_var, float eps,
                                        float* __restrict__ scale, float* __restrict__ shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float sc = gamma[c] / sqrtf(run_var[c] + eps);
    scale[c] = sc;
    shift[c] = beta[c] - run_mean[c] * sc;
}
// every eval-mode BN of the network in one launch
__global__ void bn_eval_batched_kernel(const tfe::BnEvalJob* __restrict__ jobs, int njobs, int total, float eps) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    int lo = 0, hi = njobs - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (jobs[mid].begin <= t) lo = mid; else hi = mid - 1; }
    const tfe::BnEvalJob j = jobs[lo];
    const int c = t - j.begin;
    const float sc = j.gamma[c] / sqrtf(j.run_var[c] + eps);
    j.scale[c] = sc;
    j.shift[c] = j.beta[c] - j.run_mean[c] * sc;
}
// BN backward finalize: dgamma, dbeta and the per-channel coefficients of the apply pass
__global__ void __launch_bounds__(32 * FIN_LANES) bn_bwd_finalize_kernel(float* __restrict__ partial, int nblk, int C, long long M,
                                       const float* __restrict__ gamma, const float* __restrict__ rstd,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta,
                                       float* __restrict__ coef /* [3][C]: gamma*rstd, mean(g), mean(g*xhat) */) {
    const int c = blockIdx.x * 32 + threadIdx.x;
    double s, q;
    reduce_partials<true>(partial, nblk, C, c, s, q);
    if (threadIdx.y != 0 || c >= C) return;
    if (dgamma) dgamma[c] = (float)q;
    if (dbeta) dbeta[c] = (float)s;
    coef[c] = gamma[c] * rstd[c];
    coef[C + c] = (float)(s / (double)M);
    coef[2 * C + c] = (float)(q / (double)M);
}
__global__ void __launch_bounds__(32 * FIN_LANES) colsum_finalize_kernel(float* __restrict__ partial, int nblk, int C, int Cout,
                                       float* __restrict__ out, int accumulate) {
    const int c = blockIdx.x * 32 + threadIdx.x;
    double s, q;
    reduce_partials<false>(partial, nblk, C, c, s, q);
    if (threadIdx.y != 0 || c >= Cout) return;
    out[c] = accumulate ? (float)((double)out[c] + s) : (float)s;
}

// out = act( y*scale + shift (+ res | + res*rscale + rshift) ); optional 1-bit ReLU mask of the result
// HOIST: per-channel vectors in registers for the whole kernel (loop-invariant channel group) instead of re-read per float4
template <bool HOIST>
__global__ void __launch_bounds__(EW_THREADS) bn_apply_kernel(const float* __restrict__ y, const float* __restrict__ scale,
                                                              const float* __restrict__ shift,
                                                              const float* __restrict__ res,
                                                              const float* __restrict__ rscale,
                                                              const float* __restrict__ rshift, int relu, long long n4,
                                                              int C, float* __restrict__ out, float* __restrict__ out_lo,
                                                              int mode, unsigned int* __restrict__ mask_out, int descending,
                                                              int streaming) {
    // descending: walk the tensor from its end -- the GEMM that produced y wrote it front to back, so its tail is what
    // is still in L2, and the consumer GEMM then starts at the front, which this pass wrote last
    const long long n4_up = (n4 + 31) & ~31ll;          // whole warps iterate together (the mask needs shuffles)
    const long long nchunks = (n4_up + blockDim.x - 1) / blockDim.x;
    // a chunk is blockDim * 4 = 1024 consecutive floats, a multiple of C (a power of two <= 1024): the channel group of a
    // thread is the same in every chunk, so its scale / shift live in registers for the whole kernel
    const int c = (int)((threadIdx.x * 4) & (C - 1));
    float4 sc = make_float4(0, 0, 0, 0), sh = sc;
    float4 ra = make_float4(1.f, 1.f, 1.f, 1.f), rb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (HOIST) {
        sc = ld4(scale + c); sh = ld4(shift + c);
        if (res && rscale) { ra = ld4(rscale + c); rb = ld4(rshift + c); }
    }
    for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const long long t = (descending ? nchunks - 1 - ch : ch) * blockDim.x + threadIdx.x;
        if (t >= n4_up) continue;                       // (only whole warps drop out: n4_up and blockDim are multiples of 32)
        const long long i = t * 4;
        const bool live = t < n4;
        float4 v = make_float4(0, 0, 0, 0);
        if (live) {
            v = ld4s(y + i, streaming);
            if (!HOIST) { sc = ld4(scale + c); sh = ld4(shift + c); }
            v.x = v.x * sc.x + sh.x; v.y = v.y * sc.y + sh.y; v.z = v.z * sc.z + sh.z; v.w = v.w * sc.w + sh.w;
            if (res) {
                float4 r = ld4s(res + i, streaming);
                if (!HOIST && rscale) { ra = ld4(rscale + c); rb = ld4(rshift + c); }
                if (rscale) { r.x = r.x * ra.x + rb.x; r.y = r.y * ra.y + rb.y; r.z = r.z * ra.z + rb.z; r.w = r.w * ra.w + rb.w; }
                v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
            }
            if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            store_act(out, out_lo, i, v, mode);
        }
        if (mask_out) {
            unsigned int nib = (v.x > 0.f ? 1u : 0u) | (v.y > 0.f ? 2u : 0u) | (v.z > 0.f ? 4u : 0u) | (v.w > 0.f ? 8u : 0u);
            unsigned int w = nib << ((threadIdx.x & 7) * 4);
            w |= __shfl_xor_sync(0xffffffffu, w, 1);
            w |= __shfl_xor_sync(0xffffffffu, w, 2);
            w |= __shfl_xor_sync(0xffffffffu, w, 4);
            if ((threadIdx.x & 7) == 0 && live) mask_out[i >> 5] = w;
        }
    }
}

// dy = coef0 * (g - coef1 - xhat*coef2),  g = dout (* [act > 0]);  optionally gmask_out = g (identity branch)
template <bool HOIST>
__global__ void __launch_bounds__(EW_THREADS) bn_bwd_apply_kernel(const float* __restrict__ dout,
                                                                  const float* __restrict__ act,
                                                                  const unsigned int* __restrict__ mask,
                                                                  const float* __restrict__ y,
                                                                  const float* __restrict__ mean,
                                                                  const float* __restrict__ rstd,
                                                                  const float* __restrict__ coef, long long n4, int C,
                                                                  float* __restrict__ dy, float* __restrict__ dy_lo,
                                                                  float* __restrict__ gmask_out, int mode, int streaming) {
    // blockDim * 4 = 1024 floats per block row is a multiple of C: a thread's channel group is loop-invariant, so the five
    // per-channel vectors are read ONCE into registers (they used to be re-read from L1/L2 for every float4: 5 of the 8 loads
    // per iteration -- with the max-shared carveout the L1 is too small to keep them next to the streamed tensors)
    const int c = (int)((threadIdx.x * 4) & (C - 1));
    float4 mu, rs, c0, c1, c2;
    if (HOIST) { mu = ld4(mean + c); rs = ld4(rstd + c); c0 = ld4(coef + c); c1 = ld4(coef + C + c); c2 = ld4(coef + 2 * C + c); }
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // two independent float4 groups per iteration: all six tensor loads are issued before the first use
    for (; t + stride < n4; t += 2 * stride) {
        const long long i0 = t * 4, i1 = (t + stride) * 4;
        float4 g0 = ld4s(dout + i0, streaming), g1 = ld4s(dout + i1, streaming);
        const float4 v0 = ld4s(y + i0, streaming), v1 = ld4s(y + i1, streaming);
        if (mask) { const unsigned int n0 = mask_nibble(mask, i0), n1 = mask_nibble(mask, i1); apply_mask(g0, n0); apply_mask(g1, n1); }
        else if (act) {
            const float4 o0 = ld4(act + i0), o1 = ld4(act + i1);
            if (!(o0.x > 0.f)) g0.x = 0.f; if (!(o0.y > 0.f)) g0.y = 0.f; if (!(o0.z > 0.f)) g0.z = 0.f; if (!(o0.w > 0.f)) g0.w = 0.f;
            if (!(o1.x > 0.f)) g1.x = 0.f; if (!(o1.y > 0.f)) g1.y = 0.f; if (!(o1.z > 0.f)) g1.z = 0.f; if (!(o1.w > 0.f)) g1.w = 0.f;
        }
        if (gmask_out) { st4(gmask_out + i0, g0); st4(gmask_out + i1, g1); }
        if (!HOIST) { mu = ld4(mean + c); rs = ld4(rstd + c); c0 = ld4(coef + c); c1 = ld4(coef + C + c); c2 = ld4(coef + 2 * C + c); }
        float4 r0, r1;
        r0.x = c0.x * (g0.x - c1.x - (v0.x - mu.x) * rs.x * c2.x); r0.y = c0.y * (g0.y - c1.y - (v0.y - mu.y) * rs.y * c2.y);
        r0.z = c0.z * (g0.z - c1.z - (v0.z - mu.z) * rs.z * c2.z); r0.w = c0.w * (g0.w - c1.w - (v0.w - mu.w) * rs.w * c2.w);
        r1.x = c0.x * (g1.x - c1.x - (v1.x - mu.x) * rs.x * c2.x); r1.y = c0.y * (g1.y - c1.y - (v1.y - mu.y) * rs.y * c2.y);
        r1.z = c0.z * (g1.z - c1.z - (v1.z - mu.z) * rs.z * c2.z); r1.w = c0.w * (g1.w - c1.w - (v1.w - mu.w) * rs.w * c2.w);
        store_act(dy, dy_lo, i0, r0, mode);
        store_act(dy, dy_lo, i1, r1, mode);
    }
    for (; t < n4; t += stride) {
        const long long i = t * 4;
        float4 g = ld4s(dout + i, streaming);
        if (mask) apply_mask(g, mask_nibble(mask, i));
        else if (act) { float4 o = ld4(act + i); if (!(o.x > 0.f)) g.x = 0.f; if (!(o.y > 0.f)) g.y = 0.f; if (!(o.z > 0.f)) g.z = 0.f; if (!(o.w > 0.f)) g.w = 0.f; }
        if (gmask_out) st4(gmask_out + i, g);
        const float4 v = ld4s(y + i, streaming);
        if (!HOIST) { mu = ld4(mean + c); rs = ld4(rstd + c); c0 = ld4(coef + c); c1 = ld4(coef + C + c); c2 = ld4(coef + 2 * C + c); }
        float4 r;
        r.x = c0.x * (g.x - c1.x - (v.x - mu.x) * rs.x * c2.x);
        r.y = c0.y * (g.y - c1.y - (v.y - mu.y) * rs.y * c2.y);
        r.z = c0.z * (g.z - c1.z - (v.z - mu.z) * rs.z * c2.z);
        r.w = c0.w * (g.w - c1.w - (v.w - mu.w) * rs.w * c2.w);
        store_act(dy, dy_lo, i, r, mode);
    }
}

// out = a (* [act > 0]) (+ b)
__global__ void __launch_bounds__(EW_THREADS) masked_add_kernel(const float* __restrict__ a, const float* __restrict__ act,
                                                                const float* __restrict__ b, long long n4,
                                                                float* __restrict__ out) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) {
        const long long i = t * 4;
        float4 g = ld4(a + i);
        if (act) { float4 o = ld4(act + i); if (!(o.x > 0.f)) g.x = 0.f; if (!(o.y > 0.f)) g.y = 0.f; if (!(o.z > 0.f)) g.z = 0.f; if (!(o.w > 0.f)) g.w = 0.f; }
        if (b) { float4 v = ld4(b + i); g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w; }
        st4(out + i, g);
    }
}

// ------------------------------------------------------------------------------------------------ stem
// im2col for conv1 (7x7, stride 2, pad 3) straight from the caller's NCHW image:
//   col[(b,oy,ox)][k], k = c*49 + kh*7 + kw (== OIHW flattening of conv1.weight), zero-padded to K_pad
// One block = 64 consecutive output pixels of one output row: the 3 x 7 input rows they touch (133 columns) are staged
// in shared memory with coalesced row reads, then the 64 x K_pad patch matrix goes out as whole 640-byte rows.
constexpr int IM2COL_PIX = 64, IM2COL_TW = 136;
__global__ void __launch_bounds__(EW_THREADS) stem_im2col_kernel(const float* __restrict__ x, int B, int H, int W, int Ho,
                                                                 int Wo, int K_pad, float* __restrict__ col,
                                                                 float* __restrict__ col_lo, int mode) {
    __shared__ float t[21 * IM2COL_TW];
    const int ox0 = blockIdx.x * IM2COL_PIX, oy = blockIdx.y, b = blockIdx.z;
    const int ix0 = 2 * ox0 - 3;
    for (int e = threadIdx.x; e < 21 * IM2COL_TW; e += EW_THREADS) {
        const int rowi = e / IM2COL_TW, cx = e - rowi * IM2COL_TW;       // rowi = c * 7 + kh
        const int c = rowi / 7, kh = rowi - c * 7;
        const int iy = 2 * oy + kh - 3, ix = ix0 + cx;
        float v = 0.f;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(x + ((size_t)(b * 3 + c) * H + iy) * W + ix);
        t[e] = v;
    }
    __syncthreads();
    const int kq = K_pad / 4;
    const int npix = min(IM2COL_PIX, Wo - ox0);
    const size_t row0 = ((size_t)b * Ho + oy) * Wo + ox0;
    for (int e = threadIdx.x; e < npix * kq; e += EW_THREADS) {
        const int pl = e / kq, k0 = (e - pl * kq) * 4;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + j;
            float val = 0.f;
            if (k < 147) {
                const int c = k / 49, r = k - c * 49, kh = r / 7, kw = r - kh * 7;
                val = t[(c * 7 + kh) * IM2COL_TW + 2 * pl + kw];
            }
            v[j] = val;
        }
        store_act(col, col_lo, (long long)((row0 + pl) * K_pad + k0), make_float4(v[0], v[1], v[2], v[3]), mode);
    }
}

// max-pool 3x3 stride 2 pad 1, NHWC.  Blocks walk output rows (b, oy); threads cover (ox, 4-channel group): 32-bit index
// arithmetic only (the first version spent most of its time in 64-bit divisions).
__global__ void __launch_bounds__(EW_THREADS) maxpool_fwd_kernel(const float* __restrict__ x, int B, int H, int W, int C,
                                                                 int Ho, int Wo, float* __restrict__ out,
                                                                 float* __restrict__ out_lo, int mode,
                                                                 unsigned char* __restrict__ argmax) {
    const int cgs = C / 4, per_row = Wo * cgs;
    for (int row = blockIdx.x; row < B * Ho; row += gridDim.x) {
        const int b = row / Ho, oy = row - b * Ho;
        const float* xb = x + (size_t)b * H * W * C;
        for (int t = threadIdx.x; t < per_row; t += EW_THREADS) {
            const int ox = t / cgs, cg = t - ox * cgs;
            float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            int am[4] = {-1, -1, -1, -1};
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int iy = oy * 2 + kh - 1;
                if (iy < 0 || iy >= H) continue;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int ix = ox * 2 + kw - 1;
                    if (ix < 0 || ix >= W) continue;
                    const float4 v4 = ld4(xb + ((size_t)iy * W + ix) * C + cg * 4);
                    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j)     // first maximum in (kh, kw) scan order wins (ATen's CPU max_pool2d rule)
                        if (v[j] > m[j] || am[j] < 0) { m[j] = v[j]; am[j] = kh * 3 + kw; }
                }
            }
            const long long o = ((long long)row * Wo + ox) * C + cg * 4;
            store_act(out, out_lo, o, make_float4(m[0], m[1], m[2], m[3]), mode);
            if (argmax) *reinterpret_cast<uchar4*>(argmax + o) = make_uchar4(am[0], am[1], am[2], am[3]);
        }
    }
}
// backward as a gather over the saved window argmax: input pixel (iy, ix) collects dout of the (at most 4)
// windows whose arg-max it is.  Blocks walk input rows (b, iy); float4 over channels.
__global__ void __launch_bounds__(EW_THREADS) maxpool_bwd_kernel(const unsigned char* __restrict__ argmax,
                                                                 const float* __restrict__ dout, int B, int H, int W, int C,
                                                                 int Ho, int Wo, float* __restrict__ dx) {
    const int cgs = C / 4, per_row = W * cgs;
    for (int row = blockIdx.x; row < B * H; row += gridDim.x) {
        const int b = row / H, iy = row - b * H;
        const int oy_lo = iy / 2, oy_hi = min(Ho - 1, (iy + 1) / 2);
        for (int t = threadIdx.x; t < per_row; t += EW_THREADS) {
            const int ix = t / cgs, cg = t - ix * cgs;
            const int ox_lo = ix / 2, ox_hi = min(Wo - 1, (ix + 1) / 2);
            float4 acc = make_float4(0, 0, 0, 0);
            for (int oy = oy_lo; oy <= oy_hi; ++oy)
                for (int ox = ox_lo; ox <= ox_hi; ++ox) {
                    const int want = (iy - (oy * 2 - 1)) * 3 + (ix - (ox * 2 - 1));
                    const size_t o = (((size_t)b * Ho + oy) * Wo + ox) * C + cg * 4;
                    const uchar4 am = *reinterpret_cast<const uchar4*>(argmax + o);
                    const float4 g = ld4(dout + o);
                    if (am.x == want) acc.x += g.x;
                    if (am.y == want) acc.y += g.y;
                    if (am.z == want) acc.z += g.z;
                    if (am.w == want) acc.w += g.w;
                }
            st4(dx + ((size_t)row * W + ix) * C + cg * 4, acc);
        }
    }
}

// out[b,y,x,:] = (y,x both even) ? xs[b,y/2,x/2,:] : 0     (adjoint of a stride-2 subsampling)
__global__ void __launch_bounds__(EW_THREADS) zero_insert2_kernel(const float* __restrict__ xs, const float* __restrict__ xs_lo,
                                                                  int B, int H, int W, int C, int Ho, int Wo,
                                                                  float* __restrict__ out, float* __restrict__ out_lo) {
    const long long total = (long long)B * H * W * (C / 4);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(t % (C / 4));
        const long long pix = t / (C / 4);
        const int x = (int)(pix % W), y = (int)((pix / W) % H), b = (int)(pix / ((long long)W * H));
        float4 v = make_float4(0, 0, 0, 0), l = v;
        if (!(x & 1) && !(y & 1)) {
            const long long src = (((long long)b * Ho + y / 2) * Wo + x / 2) * C + cg * 4;
            v = ld4(xs + src);
            if (xs_lo) l = ld4(xs_lo + src);
        }
        st4(out + pix * C + cg * 4, v);
        if (out_lo) st4(out_lo + pix * C + cg * 4, l);
    }
}

// ------------------------------------------------------------------------------------------------ weights
// dst[o][t][i] (o < O_pad, i < I_pad) from an OIHW source:
//   transpose == 0:  src[o][i][t]                 (fprop:  [Cout][taps][Cin])
//   transpose == 1:  src[i][o][taps-1-t]          (dgrad:  [Cin][flipped taps][Cout])
__global__ void pack_weight_kernel(const float* __restrict__ w, int O_src, int I_src, int taps, int transpose, int O_pad,
                                   int I_pad, float* __restrict__ dst, float* __restrict__ dst_lo, int mode) {
    const long long total = (long long)O_pad * taps * I_pad;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % I_pad), tp = (int)((t / I_pad) % taps), o = (int)(t / ((long long)I_pad * taps));
        float v = 0.f;
        if (!transpose) { if (o < O_src && i < I_src) v = w[((long long)o * I_src + i) * taps + tp]; }
        else            { if (i < O_src && o < I_src) v = w[((long long)i * I_src + o) * taps + (taps - 1 - tp)]; }
        if (mode == 1) dst[t] = tf_round_tf32(v);
        else if (mode == 2) { const float h = hi_part(v); dst[t] = h; dst_lo[t] = v - h; }
        else dst[t] = v;
    }
}
// all weight tensors of one pass in a single launch: element t belongs to the job whose [begin, next begin) holds it
__global__ void pack_weights_batched_kernel(const tfe::PackJob* __restrict__ jobs, int njobs, long long total, int mode) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = njobs - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (jobs[mid].begin <= t) lo = mid; else hi = mid - 1; }
        const tfe::PackJob j = jobs[lo];
        const long long u = t - j.begin;
        const int i = (int)(u % j.I_pad), tp = (int)((u / j.I_pad) % j.taps), o = (int)(u / ((long long)j.I_pad * j.taps));
        float v = 0.f;
        if (!j.transpose) { if (o < j.O_src && i < j.I_src) { v = j.src[((long long)o * j.I_src + i) * j.taps + tp]; if (j.oscale) v *= j.oscale[o]; } }
        else              { if (i < j.O_src && o < j.I_src) v = j.src[((long long)i * j.I_src + o) * j.taps + (j.taps - 1 - tp)]; }
        if (mode == 1) j.dst[u] = tf_round_tf32(v);
        else if (mode == 2) { const float h = hi_part(v); j.dst[u] = h; j.dst_lo[u] = v - h; }
        else j.dst[u] = v;
    }
}
// OIHW gradient from the packed [O_pad][taps][I_pad] wgrad output
__global__ void unpack_wgrad_kernel(const float* __restrict__ dwp, int O, int I, int taps, int I_pad,
                                    float* __restrict__ dw) {
    const long long total = (long long)O * I * taps;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int tp = (int)(t % taps), i = (int)((t / taps) % I), o = (int)(t / ((long long)taps * I));
        dw[t] = dwp[((long long)o * taps + tp) * I_pad + i];
    }
}

// ------------------------------------------------------------------------------------------------ heads
// model.py:104-126: out[b,c,y,x] = s3[b,y,x,c] + sum_{i,j} s4[b,i,j,c] * up[c][y+1-2i][x+1-2j]   (NCHW result)
// One block = 32 consecutive x of one (b, y): the values are computed channel-fastest (coalesced NHWC reads), transposed
// through shared memory and written as 128-byte rows of the NCHW result.
constexpr int HEAD_TX = 32, HEAD_CMAX = 256;
__global__ void __launch_bounds__(EW_THREADS) head_combine_fwd_kernel(const float* __restrict__ s3, const float* __restrict__ s4,
                                                                      const float* __restrict__ up /*[Cn][16]*/, int B,
                                                                      int H3, int W3, int H4, int W4, int Cn, int Cp,
                                                                      float* __restrict__ out) {
    extern __shared__ float tile[];                     // [Cp][HEAD_TX + 1]
    const int x0 = blockIdx.x * HEAD_TX, y = blockIdx.y, b = blockIdx.z;
    const int nx = min(HEAD_TX, W3 - x0);
    const int i0 = (y + 1) >> 1;
    for (int e = threadIdx.x; e < nx * Cp; e += EW_THREADS) {
        const int xl = e / Cp, c = e - xl * Cp, x = x0 + xl;
        float v = 0.f;
        if (c < Cn) {
            v = s3[(((size_t)b * H3 + y) * W3 + x) * Cp + c];
            const int j0 = (x + 1) >> 1;
#pragma unroll
            for (int di = 0; di < 2; ++di) {
                const int i = i0 - di, ky = y + 1 - 2 * i;
                if (i < 0 || i >= H4) continue;
#pragma unroll
                for (int dj = 0; dj < 2; ++dj) {
                    const int j = j0 - dj, kx = x + 1 - 2 * j;
                    if (j < 0 || j >= W4) continue;
                    v += s4[(((size_t)b * H4 + i) * W4 + j) * Cp + c] * up[c * 16 + ky * 4 + kx];
                }
            }
        }
        tile[c * (HEAD_TX + 1) + xl] = v;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < Cn * HEAD_TX; e += EW_THREADS) {
        const int c = e / HEAD_TX, xl = e - c * HEAD_TX;
        if (xl < nx) out[(((size_t)b * Cn + c) * H3 + y) * W3 + x0 + xl] = tile[c * (HEAD_TX + 1) + xl];
    }
}
// adjoint, step 1: ds3[b,y,x,c] = dout[b,c,y,x] (padded channels = 0) -- an NCHW -> NHWC transpose through shared memory
__global__ void __launch_bounds__(EW_THREADS) head_transpose_bwd_kernel(const float* __restrict__ dout, int B, int H3, int W3,
                                                                        int Cn, int Cp, float* __restrict__ ds3,
                                                                        float* __restrict__ ds3_lo) {
    extern __shared__ float tile[];                     // [Cp][HEAD_TX + 1]
    const int x0 = blockIdx.x * HEAD_TX, y = blockIdx.y, b = blockIdx.z;
    const int nx = min(HEAD_TX, W3 - x0);
    for (int e = threadIdx.x; e < Cp * HEAD_TX; e += EW_THREADS) {
        const int c = e / HEAD_TX, xl = e - c * HEAD_TX;
        tile[c * (HEAD_TX + 1) + xl] = (c < Cn && xl < nx) ? dout[(((size_t)b * Cn + c) * H3 + y) * W3 + x0 + xl] : 0.f;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nx * Cp; e += EW_THREADS) {
        const int xl = e / Cp, c = e - xl * Cp;
        const float v = tile[c * (HEAD_TX + 1) + xl];
        const size_t o = (((size_t)b * H3 + y) * W3 + x0 + xl) * Cp + c;
        if (ds3_lo) { const float h = hi_part(v); ds3[o] = h; ds3_lo[o] = v - h; }      // parity mode: exact (hi, lo) TF32 split
        else ds3[o] = v;
    }
}
// adjoint, step 2: ds4[b,i,j,c] = sum_{y,x} ds3[b,y,x,c] * up[c][y+1-2i][x+1-2j]   (channel-fastest on both sides)
__global__ void __launch_bounds__(EW_THREADS) head_upsample_bwd_kernel(const float* __restrict__ ds3, const float* __restrict__ ds3_lo,
                                                                       const float* __restrict__ up,
                                                                       int B, int H3, int W3, int H4, int W4, int Cn, int Cp,
                                                                       float* __restrict__ ds4, float* __restrict__ ds4_lo) {
    const int total = B * H4 * W4 * Cp;
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < total; u += gridDim.x * blockDim.x) {
        const int c = u % Cp, pix = u / Cp;
        const int j = pix % W4, i = (pix / W4) % H4, b = pix / (W4 * H4);
        float v = 0.f;
        if (c < Cn) {
#pragma unroll
            for (int ky = 0; ky < 4; ++ky) {
                const int y = 2 * i - 1 + ky;
                if (y < 0 || y >= H3) continue;
#pragma unroll
                for (int kx = 0; kx < 4; ++kx) {
                    const int x = 2 * j - 1 + kx;
                    if (x < 0 || x >= W3) continue;
                    const size_t o = (((size_t)b * H3 + y) * W3 + x) * Cp + c;
                    v += (ds3_lo ? ds3[o] + ds3_lo[o] : ds3[o]) * up[c * 16 + ky * 4 + kx];
                }
            }
        }
        if (ds4_lo) { const float h = hi_part(v); ds4[u] = h; ds4_lo[u] = v - h; }
        else ds4[u] = v;
    }
}
// up[c][k] = w[c][c][k]  (the ConvTranspose2d weight must be diagonal, model.py:45-65); offdiag = max |off-diagonal|
__global__ void extract_upsample_diag_kernel(const float* __restrict__ w, int Cn, float* __restrict__ up,
                                             float* __restrict__ offdiag_max) {
    const long long total = (long long)Cn * Cn * 16;
    float m = 0.f;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(t % 16), co = (int)((t / 16) % Cn), ci = (int)(t / (16LL * Cn));
        if (ci == co) up[ci * 16 + k] = w[t];
        else m = fmaxf(m, fabsf(w[t]));
    }
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16)); m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(offdiag_max), __float_as_int(m));
}

}  // namespace

// ================================================================================================ host API
// (internal C++ interface used by tf_model.cu)

namespace tfe {

// The chain's elementwise kernels ask for the MAX-shared-memory carveout although they use (almost) none: an SM only
// changes its L1 / shared split when it is idle, so a side-stream weight-gradient CTA (194 KB of shared memory) could
// not join an SM that was running L1-preferring blocks until all of them had retired (seen as a ~55 us late start).
template <typename K>
static int prefer_shared_carveout(K kernel) {
    TF_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    return TF_OK;
}
#define RC_CARVEOUT(kernel)                                                         \
    do {                                                                            \
        static bool done_ = false;                                                  \
        if (!done_) { int rc_ = prefer_shared_carveout(kernel); if (rc_) return rc_; done_ = true; }  \
    } while (0)

static int reduce_blocks(long long M, int C) {
    const int lanes = EW_THREADS / (C / 4);
    long long b = (M + (long long)lanes * 8 - 1) / ((long long)lanes * 8);
    return (int)std::min<long long>(std::max<long long>(b, 1), COLREDUCE_MAX_BLOCKS);
}

int bn_finalize_train(float* partial, int nblk, long long M, int C, const float* gamma, const float* beta, float eps,
                      float momentum, float* run_mean, float* run_var, float* scale, float* shift, float* save_mean,
                      float* save_rstd, cudaStream_t st) {
    bn_finalize_train_kernel<<<(C + 31) / 32, dim3(32, FIN_LANES), 0, st>>>(partial, nblk, C, M, gamma, beta, eps, momentum, run_mean,
                                                              run_var, scale, shift, save_mean, save_rstd);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
int bn_scale_shift_eval(int C, const float* gamma, const float* beta, const float* run_mean, const float* run_var,
                        float eps, float* scale, float* shift, cudaStream_t st) {
    bn_finalize_eval_kernel<<<(C + 127) / 128, 128, 0, st>>>(C, gamma, beta, run_mean, run_var, eps, scale, shift);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
int bn_scale_shift_eval_batched(const BnEvalJob* jobs_device, int njobs, int total_channels, float eps, cudaStream_t st) {
    bn_eval_batched_kernel<<<(total_channels + 255) / 256, 256, 0, st>>>(jobs_device, njobs, total_channels, eps);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
int bn_apply(const float* y, const float* scale, const float* shift, const float* res, const float* rscale,
             const float* rshift, int relu, long long M, int C, float* out, float* out_lo, int mode, unsigned int* mask_out,
             cudaStream_t st) {
    TF_REQUIRE(C >= 4 && C <= 1024 && (C & (C - 1)) == 0, "bn_apply: C=%d must be a power of two in [4, 1024]", C);
    const long long n4 = M * C / 4;
    if (tfg::debug_flag(15) & 8)
        bn_apply_kernel<true><<<ew_blocks(n4, 2), EW_THREADS, 0, st>>>(y, scale, shift, res, rscale, rshift, relu, n4, C, out, out_lo, mode, mask_out,
                                                                       tfg::debug_flag(9) & 1, !(tfg::debug_flag(15) & 1));
    else
        bn_apply_kernel<false><<<ew_blocks(n4, 2), EW_THREADS, 0, st>>>(y, scale, shift, res, rscale, rshift, relu, n4, C, out, out_lo, mode, mask_out,
                                                                        tfg::debug_flag(9) & 1, !(tfg::debug_flag(15) & 1));
    TF_LAUNCH_CHECK();
    return TF_OK;
}
int bn_backward(const float* dout, const float* act, const unsigned int* mask, const float* y, const float* save_mean, const float* save_rstd,
                const float* gamma, long long M, int C, float* dgamma, float* dbeta, float* dy, float* dy_lo,
                float* gmask_out, int mode, float* slots /* [BN_BWD_SLOTS][2][C], zero on entry, left zeroed */, float* coef, cudaStream_t st) {
    TF_REQUIRE(C >= 64 && C <= 1024 && (C & (C - 1)) == 0, "bn_backward: C=%d unsupported (power of two in [64, 1024])", C);
    RC_CARVEOUT(colreduce_kernel<1>); RC_CARVEOUT(bn_bwd_finalize_kernel); RC_CARVEOUT(bn_bwd_apply_kernel<true>); RC_CARVEOUT(bn_bwd_apply_kernel<false>);
    int nb = reduce_blocks(M, C);
    if (tfg::debug_flag(15) & 32) {          // A/B: twice the blocks (the slot accumulation does not depend on the block count)
        const int lanes = EW_THREADS / (C / 4);
        nb = (int)std::min<long long>(std::max<long long>((M + (long long)lanes * 4 - 1) / ((long long)lanes * 4), 1), 2 * COLREDUCE_MAX_BLOCKS);
    }
    colreduce_kernel<1><<<nb, EW_THREADS, 0, st>>>(dout, y, act, mask, save_mean, save_rstd, M, C, slots, BN_BWD_SLOTS, !((tfg::debug_flag(9) >> 1) & 1), !(tfg::debug_flag(15) & 4));     // descending by default (measured -0.2 ms/step); tf_debug_set(9, 2): ascending
    bn_bwd_finalize_kernel<<<(C + 31) / 32, dim3(32, FIN_LANES), 0, st>>>(slots, nb < BN_BWD_SLOTS ? nb : BN_BWD_SLOTS, C, M, gamma, save_rstd, dgamma, dbeta, coef);
    const long long n4 = M * C / 4;
    if (tfg::debug_flag(15) & 16)          // A/B: per-channel vectors re-read per float4 (round-1 form)
        bn_bwd_apply_kernel<false><<<ew_blocks(n4, 2), EW_THREADS, 0, st>>>(dout, act, mask, y, save_mean, save_rstd, coef, n4, C, dy, dy_lo,
                                                                           gmask_out, mode, !(tfg::debug_flag(15) & 2));
    else
        bn_bwd_apply_kernel<true><<<ew_blocks(n4, 2), EW_THREADS, 0, st>>>(dout, act, mask, y, save_mean, save_rstd, coef, n4, C, dy, dy_lo,
                                                                          gmask_out, mode, !(tfg::debug_flag(15) & 2));
    TF_LAUNCH_CHECK();
    return TF_OK;
}
int column_stats(const float* y, long long M, int C, float* partial, int* nblk, cudaStream_t st) {
    TF_REQUIRE(C >= 64 && C <= 1024 && (C & (C - 1)) == 0 && M > 0, "column_stats: C=%d unsupported (power of two in [64, 1024])", C);
    const int nb = reduce_blocks(M, C);
    colreduce_kernel<0><<<nb, EW_THREADS, 0, st>>>(y, nullptr, nullptr, nullptr, nullptr, nullptr, M, C, partial, 0, 0, 0);
    TF_LAUNCH_CHECK();
    *nblk = nb;
    return TF_OK;
}
int column_sum(const float* a, long long M, int C, int Cout, float* out, float* partial, cudaStream_t st, int accumulate) {
    RC_CARVEOUT(colreduce_kernel<2>); RC_CARVEOUT(colsum_finalize_kernel);
    TF_REQUIRE(C >= 64 && C <= 1024 && (C & (C - 1)) == 0, "column_sum: C=%d unsupported (power of two in [64, 1024])", C);
    const int nb = reduce_blocks(M, C);
    colreduce_kernel<2><<<nb, EW_THREADS, 0, st>>>(a, nullptr, nullptr, nullptr, nullptr, nullptr, M, C, partial, 0, 0, 0);
    colsum_finalize_kernel<<<(C + 31) / 32, dim3(32, FIN_LANES), 0, st>>>(partial, nb, C, Cout, out, accumulate);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
int masked_add(const float* a, const float* act, const float* b, long long n, float* out, cudaStream_t st) {
    RC_CARVEOUT(masked_add_kernel);
    masked_add_kernel<<<ew_blocks(n / 4, 2), EW_THREADS, 0, st>>>(a, act, b, n / 4, out);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
int stem_im2col(const float* x_nchw, int B, int H, int W, int Ho, int Wo, int K_pad, float* col, float* col_lo, int mode,
                cudaStream_t st) {
    TF_REQUIRE(K_pad % 4 == 0 && K_pad >= 148 && Ho <= 65535 && B <= 65535, "stem_im2col: unsupported shape");
    stem_im2col_kernel<<<dim3((Wo + IM2COL_PIX - 1) / IM2COL_PIX, Ho, B), EW_THREADS, 0, st>>>(x_nchw, B, H, W, Ho, Wo, K_pad, col, col_lo, mode);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
int maxpool_fwd(const float* x, int B, int H, int W, int C, int Ho, int Wo, float* out, float* out_lo, int mode,
                unsigned char* argmax, cudaStream_t st) {
    maxpool_fwd_kernel<<<std::min(B * Ho, 148 * 16), EW_THREADS, 0, st>>>(x, B, H, W, C, Ho, Wo, out, out_lo, mode, argmax);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
int maxpool_bwd(const unsigned char* argmax, const float* dout, int B, int H, int W, int C, int Ho, int Wo, float* dx, cudaStream_t st) {
    RC_CARVEOUT(maxpool_bwd_kernel);
    maxpool_bwd_kernel<<<std::min(B * H, 148 * 16), EW_THREADS, 0, st>>>(argmax, dout, B, H, W, C, Ho, Wo, dx);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
int zero_insert2(const float* xs, const float* xs_lo, int B, int H, int W, int C, float* out, float* out_lo, cudaStream_t st) {
    RC_CARVEOUT(zero_insert2_kernel);
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    zero_insert2_kernel<<<ew_blocks((long long)B * H * W * C / 4), EW_THREADS, 0, st>>>(xs, xs_lo, B, H, W, C, Ho, Wo, out, out_lo);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
int pack_weight(const float* w, int O_src, int I_src, int taps, int transpose, int O_pad, int I_pad, float* dst,
                float* dst_lo, int mode, cudaStream_t st) {
    pack_weight_kernel<<<ew_blocks((long long)O_pad * taps * I_pad), EW_THREADS, 0, st>>>(w, O_src, I_src, taps, transpose, O_pad, I_pad, dst, dst_lo, mode);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
int pack_weights_batched(const PackJob* jobs_device, int njobs, long long total, int mode, cudaStream_t st) {
    pack_weights_batched_kernel<<<ew_blocks(total), EW_THREADS, 0, st>>>(jobs_device, njobs, total, mode);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
int unpack_wgrad(const float* dwp, int O, int I, int taps, int I_pad, float* dw, cudaStream_t st) {
    RC_CARVEOUT(unpack_wgrad_kernel);
    unpack_wgrad_kernel<<<ew_blocks((long long)O * I * taps), EW_THREADS, 0, st>>>(dwp, O, I, taps, I_pad, dw);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
int head_combine_fwd(const float* s3, const float* s4, const float* up, int B, int H3, int W3, int H4, int W4, int Cn,
                     int Cp, float* out_nchw, cudaStream_t st) {
    TF_REQUIRE(Cp <= HEAD_CMAX && H3 <= 65535 && B <= 65535, "head_combine_fwd: unsupported shape");
    const size_t smem = (size_t)Cp * (HEAD_TX + 1) * sizeof(float);
    head_combine_fwd_kernel<<<dim3((W3 + HEAD_TX - 1) / HEAD_TX, H3, B), EW_THREADS, smem, st>>>(s3, s4, up, B, H3, W3, H4, W4, Cn, Cp, out_nchw);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
int head_combine_bwd(const float* dout_nchw, const float* up, int B, int H3, int W3, int H4, int W4, int Cn, int Cp,
                     float* ds3, float* ds4, cudaStream_t st, float* ds3_lo, float* ds4_lo) {
    TF_REQUIRE(Cp <= HEAD_CMAX && H3 <= 65535 && B <= 65535, "head_combine_bwd: unsupported shape");
    TF_REQUIRE((long long)B * H4 * W4 * Cp < (1ll << 31), "head_combine_bwd: tensor too large");
    const size_t smem = (size_t)Cp * (HEAD_TX + 1) * sizeof(float);
    head_transpose_bwd_kernel<<<dim3((W3 + HEAD_TX - 1) / HEAD_TX, H3, B), EW_THREADS, smem, st>>>(dout_nchw, B, H3, W3, Cn, Cp, ds3, ds3_lo);
    head_upsample_bwd_kernel<<<ew_blocks((long long)B * H4 * W4 * Cp), EW_THREADS, 0, st>>>(ds3, ds3_lo, up, B, H3, W3, H4, W4, Cn, Cp, ds4, ds4_lo);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
int extract_upsample_diag(const float* w, int Cn, float* up, float* offdiag_max, cudaStream_t st) {
    TF_CHECK_CUDA(cudaMemsetAsync(offdiag_max, 0, sizeof(float), st));
    extract_upsample_diag_kernel<<<ew_blocks((long long)Cn * Cn * 16), EW_THREADS, 0, st>>>(w, Cn, up, offdiag_max);
    TF_LAUNCH_CHECK();
    return TF_OK;
}

}  // namespace tfe
