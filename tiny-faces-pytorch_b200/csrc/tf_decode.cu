// Threshold + order-preserving compaction + anchor decode of the detection maps.
// Replaces /root/reference/tinyfaces/models/utils.py:4-100 (get_bboxes + regression_refinement) and the
// sigmoid / D2H / transpose steps of /root/reference/tinyfaces/evaluation.py:61-71.
//
// Candidates are emitted in the reference's np.where order, i.e. C order over (b, y, x, c) (utils.py:46-47),
// so keep indices of the NMS that follows index the same concatenation.  HBM-bound: the 125-channel fp32 map
// is read once (500 B/pixel, sigmoid fused) and 48 B are written per candidate.
//
// Three launches: per-pixel pass counts -> single-CTA scan of the per-block totals -> emit.
// Box arithmetic follows the reference's operation order in float64 with explicit _rn intrinsics
// (numpy does not contract a*b+c into an FMA); exp is evaluated to float32 precision as numpy does.
#include "tf_common.cuh"

namespace {

constexpr int DEC_THREADS = 256;
constexpr int MAX_T = 32;

struct DecodeParams {
    const float* cls; const float* reg; const float* prob;   // prob may be null -> sigmoid(cls)
    long long cb, cy, cx, cc;        // element strides of cls/prob for (b, y, x, c)
    long long rb, ry, rx, rc;        // element strides of reg for (b, y, x, channel)
    int B, H, W, T;
    float thresh;
    unsigned int invalid_x;          // bit x set -> column x of the map is zeroed (shipped quirk, utils.py:44)
    unsigned int invalid_t;          // bit c set -> template c is zeroed (the intended behaviour)
    long long stride_y, offset_y, stride_x, offset_x;   // rf['stride'], rf['offset'] (integers)
    double factor;                   // 1 / scale
    double cw[MAX_T], ch[MAX_T];     // template extents + 1 (utils.py:54-55)
};

__device__ __forceinline__ bool passes(const DecodeParams& p, long long base, int x, int c, float& logit) {
    logit = p.cls[base + c * p.cc];
    float pr = p.prob ? p.prob[base + c * p.cc] : 1.0f / (1.0f + __expf(-logit));
    if (x < 32 && ((p.invalid_x >> x) & 1u)) pr = 0.0f;
    if ((p.invalid_t >> c) & 1u) pr = 0.0f;
    return pr > p.thresh;
}

__global__ void __launch_bounds__(DEC_THREADS) count_kernel(DecodeParams p, long long npix,
                                                            unsigned char* __restrict__ pix_count,
                                                            int* __restrict__ block_count) {
    const long long pix = (long long)blockIdx.x * DEC_THREADS + threadIdx.x;
    int cnt = 0;
    if (pix < npix) {
        const int x = (int)(pix % p.W);
        const int y = (int)((pix / p.W) % p.H);
        const int b = (int)(pix / ((long long)p.W * p.H));
        const long long base = b * p.cb + y * p.cy + x * p.cx;
        for (int c = 0; c < p.T; ++c) { float l; cnt += passes(p, base, x, c, l) ? 1 : 0; }
        pix_count[pix] = (unsigned char)cnt;
    }
    __shared__ int s[DEC_THREADS / 32];
    int w = tf_warp_sum(cnt);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < DEC_THREADS / 32; ++i) t += s[i];
        block_count[blockIdx.x] = t;
    }
}

// exclusive scan of the per-block totals by one CTA (chunks of 1024 with a running carry)
__global__ void __launch_bounds__(1024) scan_blocks_kernel(const int* __restrict__ block_count, int nblocks,
                                                           long long* __restrict__ block_offset,
                                                           long long* __restrict__ total) {
    __shared__ long long warp_tot[32];
    __shared__ long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + threadIdx.x;
        long long v = i < nblocks ? block_count[i] : 0, incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
        __syncthreads();
        long long wbase = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wbase += warp_tot[w];
        const long long carry = carry_s;
        if (i < nblocks) block_offset[i] = carry + wbase + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + wbase + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

__global__ void __launch_bounds__(DEC_THREADS) emit_kernel(DecodeParams p, long long npix,
                                                           const unsigned char* __restrict__ pix_count,
                                                           const long long* __restrict__ block_offset,
                                                           double* __restrict__ boxes, double* __restrict__ scores,
                                                           long long* __restrict__ src, long long capacity) {
    const long long pix = (long long)blockIdx.x * DEC_THREADS + threadIdx.x;
    const int cnt = pix < npix ? pix_count[pix] : 0;
    // block-exclusive scan of cnt
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    __shared__ int warp_tot[DEC_THREADS / 32];
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wbase += warp_tot[w];
    long long pos = block_offset[blockIdx.x] + wbase + incl - cnt;
    if (cnt == 0) return;
    const int x = (int)(pix % p.W);
    const int y = (int)((pix / p.W) % p.H);
    const int b = (int)(pix / ((long long)p.W * p.H));
    const long long base = b * p.cb + y * p.cy + x * p.cx;
    const long long rbase = b * p.rb + y * p.ry + x * p.rx;
    const double cyc = (double)(y * p.stride_y + p.offset_y);      // utils.py:53 (integer arithmetic)
    const double cxc = (double)(x * p.stride_x + p.offset_x);
    for (int c = 0; c < p.T; ++c) {
        float logit;
        if (!passes(p, base, x, c, logit)) continue;
        if (pos < capacity) {
            const float tx = p.reg[rbase + (long long)c * p.rc];
            const float ty = p.reg[rbase + (long long)(p.T + c) * p.rc];
            const float tw = p.reg[rbase + (long long)(2 * p.T + c) * p.rc];
            const float th = p.reg[rbase + (long long)(3 * p.T + c) * p.rc];
            const double cw = p.cw[c], ch = p.ch[c];
            const double rcx = __dadd_rn(cxc, __dmul_rn(cw, (double)tx));          // utils.py:81-85
            const double rcy = __dadd_rn(cyc, __dmul_rn(ch, (double)ty));
            const double rcw = __dmul_rn(cw, (double)(float)exp((double)tw));      // utils.py:87-88 (fp32 exp)
            const double rch = __dmul_rn(ch, (double)(float)exp((double)th));
            const double hw = __ddiv_rn(rcw, 2.0), hh = __ddiv_rn(rch, 2.0);
            double* o = boxes + pos * 4;                                             // utils.py:97-98, :73-74
            o[0] = __dmul_rn(__dsub_rn(rcx, hw), p.factor);
            o[1] = __dmul_rn(__dsub_rn(rcy, hh), p.factor);
            o[2] = __dmul_rn(__dadd_rn(rcx, hw), p.factor);
            o[3] = __dmul_rn(__dadd_rn(rcy, hh), p.factor);
            scores[pos] = (double)logit;
            if (src) src[pos] = (pix * p.T) + c;
        }
        ++pos;
    }
}

}  // namespace

TF_API int tf_decode_workspace_bytes(int64_t B, int64_t H, int64_t W, size_t* bytes) {
    TF_REQUIRE(bytes && B > 0 && H > 0 && W > 0, "tf_decode_workspace_bytes: bad args");
    const long long npix = B * H * W;
    const long long nblk = (npix + DEC_THREADS - 1) / DEC_THREADS;
    *bytes = tf_align_up(npix, 256) + tf_align_up(nblk * 4, 256) + tf_align_up(nblk * 8, 256) + 1024;
    return TF_OK;
}

// templates_host: HOST pointer to T x 5 float64 rows (x1, y1, x2, y2, scale), T <= 32.
TF_API int tf_decode(const float* cls, const float* reg, const float* prob, const int64_t* cls_strides,
                     const int64_t* reg_strides, int B, int H, int W, int T, const double* templates_host,
                     float prob_thresh, uint32_t invalid_x_mask, uint32_t invalid_t_mask, const int64_t* rf_stride,
                     const int64_t* rf_offset, double scale, double* boxes, double* scores, int64_t* src_index,
                     int64_t capacity, int64_t* count, void* workspace, size_t workspace_bytes, void* stream) {
    TF_REQUIRE(cls && reg && cls_strides && reg_strides && templates_host && rf_stride && rf_offset && count,
               "tf_decode: null pointer");
    TF_REQUIRE(B > 0 && H > 0 && W > 0 && T > 0 && T <= MAX_T, "tf_decode: bad shape B=%d H=%d W=%d T=%d", B, H, W, T);
    TF_REQUIRE(capacity == 0 || (boxes && scores), "tf_decode: null output");
    size_t need;
    tf_decode_workspace_bytes(B, H, W, &need);
    if (workspace_bytes < need) { tf_set_error("tf_decode: workspace %zu < %zu", workspace_bytes, need); return TF_ERR_WORKSPACE; }
    DecodeParams p;
    p.cls = cls; p.reg = reg; p.prob = prob;
    p.cb = cls_strides[0]; p.cy = cls_strides[1]; p.cx = cls_strides[2]; p.cc = cls_strides[3];
    p.rb = reg_strides[0]; p.ry = reg_strides[1]; p.rx = reg_strides[2]; p.rc = reg_strides[3];
    p.B = B; p.H = H; p.W = W; p.T = T; p.thresh = prob_thresh;
    p.invalid_x = invalid_x_mask; p.invalid_t = invalid_t_mask;
    p.stride_y = rf_stride[0]; p.stride_x = rf_stride[1]; p.offset_y = rf_offset[0]; p.offset_x = rf_offset[1];
    p.factor = 1.0 / scale;
    for (int c = 0; c < T; ++c) {
        const double* t = templates_host + 5 * c;
        p.cw[c] = t[2] - t[0] + 1;
        p.ch[c] = t[3] - t[1] + 1;
    }
    const long long npix = (long long)B * H * W;
    const int nblk = (int)((npix + DEC_THREADS - 1) / DEC_THREADS);
    TfArena ar(workspace, workspace_bytes);
    unsigned char* pix_count = ar.take<unsigned char>(npix);
    int* block_count = ar.take<int>(nblk);
    long long* block_offset = ar.take<long long>(nblk);
    cudaStream_t st = (cudaStream_t)stream;
    count_kernel<<<nblk, DEC_THREADS, 0, st>>>(p, npix, pix_count, block_count);
    scan_blocks_kernel<<<1, 1024, 0, st>>>(block_count, nblk, block_offset, (long long*)count);
    emit_kernel<<<nblk, DEC_THREADS, 0, st>>>(p, npix, pix_count, block_offset, boxes, scores, (long long*)src_index, capacity);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
