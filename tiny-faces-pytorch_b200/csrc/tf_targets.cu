// Training-target generation on the GPU: the dense template-vs-ground-truth IoU volume and the class / regression heat
// maps derived from it.  Replaces DataProcessor.get_heatmaps + get_regression
// (/root/reference/tinyfaces/datasets/processor.py:157-277) and compute_dense_overlap
// (/root/reference/tinyfaces/datasets/dense_overlap.py:4-75, a 4-deep pure-Python loop over gt x template x x x y that
// dominates the reference's DataLoader time).  SURVEY.md section 8f.3.
//
// All arithmetic is float64 in the reference's operation order (explicit _rn intrinsics, no FMA contraction), so the IoU
// volume -- including its np.around(., 14) -- the arg-max choices and therefore the class map are bit-identical; only
// tw / th go through log(), which CUDA evaluates to <= 1 ulp (checked at 1e-14 relative).
//
//   K1  one thread per (y, x, template): IoU against every ground-truth box (rounded to 14 decimals), + 1e-6 * jitter
//       (the reference's tie-breaking noise, processor.py:203), best object (first maximum) -> tx, ty, tw, th, best IoU
//   K2  one block per ground-truth box: its best (y, x, template) cell over the whole volume (first maximum)
//   K3  one thread per cell: label = +1 (a box's best cell with IoU > neg, or IoU >= pos), 0 (gray zone), -1; border rule
#include "tf_common.cuh"
#include <algorithm>

namespace {

constexpr int MAX_TPL = 64;
struct TargetParams {
    int vsy, vsx, nt, ng;
    double ofy, ofx, sty, stx;
    double pos, neg;
    double dx1[MAX_TPL], dy1[MAX_TPL], dx2[MAX_TPL], dy2[MAX_TPL];
};

// np.around(v, 14): rint(v * 1e14) / 1e14
__device__ __forceinline__ double around14(double v) { return __ddiv_rn(rint(__dmul_rn(v, 1e14)), 1e14); }

__global__ void __launch_bounds__(256) iou_kernel(const TargetParams p, const double* __restrict__ boxes,
                                                  const double* __restrict__ jitter, unsigned long long seed,
                                                  double* __restrict__ iou_out, double* __restrict__ iou_p_out,
                                                  double* __restrict__ best_iou, double* __restrict__ regress) {
    const long long cells = (long long)p.vsy * p.vsx * p.nt;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    const int t = (int)(c % p.nt);
    const int x = (int)((c / p.nt) % p.vsx);
    const int y = (int)(c / ((long long)p.nt * p.vsx));
    // dense_overlap.py:46-52: cx = ofx + x * (stx / zmx), zm = 1
    const double cx = __dadd_rn(p.ofx, __dmul_rn((double)x, p.stx)), cy = __dadd_rn(p.ofy, __dmul_rn((double)y, p.sty));
    const double x1 = __dadd_rn(p.dx1[t], cx), y1 = __dadd_rn(p.dy1[t], cy), x2 = __dadd_rn(p.dx2[t], cx), y2 = __dadd_rn(p.dy2[t], cy);
    const double fh = __dadd_rn(__dsub_rn(p.dy2[t], p.dy1[t]), 1.0), fw = __dadd_rn(__dsub_rn(p.dx2[t], p.dx1[t]), 1.0);
    const double farea = __dmul_rn(fw, fh);
    double best = 0.0; int bestg = 0;
    for (int g = 0; g < p.ng; ++g) {
        const double gx1 = boxes[4 * g], gy1 = boxes[4 * g + 1], gx2 = boxes[4 * g + 2], gy2 = boxes[4 * g + 3];
        const double bw = __dadd_rn(__dsub_rn(gx2, gx1), 1.0), bh = __dadd_rn(__dsub_rn(gy2, gy1), 1.0);
        const double barea = __dmul_rn(bw, bh);
        const double xx1 = x1 >= gx1 ? x1 : gx1, yy1 = y1 >= gy1 ? y1 : gy1;         // python max(a, b)
        const double xx2 = gx2 < x2 ? gx2 : x2, yy2 = gy2 < y2 ? gy2 : y2;           // python min(a, b)
        const double iw = __dadd_rn(__dsub_rn(xx2, xx1), 1.0), ih = __dadd_rn(__dsub_rn(yy2, yy1), 1.0);
        double o = 0.0;
        if (ih > 0.0 && iw > 0.0) {
            const double ia = __dmul_rn(iw, ih);
            o = __ddiv_rn(ia, __dsub_rn(__dadd_rn(farea, barea), ia));
        }
        o = around14(o);
        const long long idx = c * p.ng + g;
        if (iou_out) iou_out[idx] = o;
        double j;
        if (jitter) j = jitter[idx];
        else {                                       // device noise: splitmix64 of (seed, idx) -> [0, 1)
            unsigned long long z = seed + (unsigned long long)idx * 0x9E3779B97F4A7C15ull;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
            j = (double)(z >> 11) * (1.0 / 9007199254740992.0);
        }
        const double op = __dadd_rn(o, __dmul_rn(1e-6, j));                           // processor.py:203
        iou_p_out[idx] = op;
        if (g == 0 || op > best) { best = op; bestg = g; }                             // argmax: first maximum
    }
    best_iou[c] = best;
    // get_regression (processor.py:166-221) for the best object of this cell
    const double gx1 = boxes[4 * bestg], gy1 = boxes[4 * bestg + 1], gx2 = boxes[4 * bestg + 2], gy2 = boxes[4 * bestg + 3];
    const double fcx = __ddiv_rn(__dadd_rn(gx1, gx2), 2.0), fcy = __ddiv_rn(__dadd_rn(gy1, gy2), 2.0);
    const double tx = __ddiv_rn(__dsub_rn(fcx, cx), fw), ty = __ddiv_rn(__dsub_rn(fcy, cy), fh);
    const double tw = log(__ddiv_rn(__dadd_rn(__dsub_rn(gx2, gx1), 1.0), fw));
    const double th = log(__ddiv_rn(__dadd_rn(__dsub_rn(gy2, gy1), 1.0), fh));
    double* r = regress + ((long long)y * p.vsx + x) * 4 * p.nt;                        // channels: tx[nt] ty[nt] tw[nt] th[nt]
    r[t] = tx; r[p.nt + t] = ty; r[2 * p.nt + t] = tw; r[3 * p.nt + t] = th;
}

// per ground-truth box: the first cell holding its maximum perturbed IoU (np.argmax over the flattened volume)
__global__ void __launch_bounds__(1024) object_best_kernel(const double* __restrict__ iou_p, long long cells, int ng, double neg,
                                                           unsigned char* __restrict__ flag) {
    const int g = blockIdx.x;
    __shared__ double s_val[32];
    __shared__ long long s_idx[32];
    double v = -1.0; long long vi = 0x7fffffffffffffffll;
    for (long long c = threadIdx.x; c < cells; c += blockDim.x) {
        const double o = iou_p[c * ng + g];
        if (o > v) { v = o; vi = c; }                 // increasing c per thread: keeps the first maximum
    }
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, off);
        const long long oi = __shfl_xor_sync(0xffffffffu, vi, off);
        if (ov > v || (ov == v && oi < vi)) { v = ov; vi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = v; s_idx[threadIdx.x >> 5] = vi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        v = threadIdx.x < (blockDim.x >> 5) ? s_val[threadIdx.x] : -1.0;
        vi = threadIdx.x < (blockDim.x >> 5) ? s_idx[threadIdx.x] : 0x7fffffffffffffffll;
        for (int off = 16; off > 0; off >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, v, off);
            const long long oi = __shfl_xor_sync(0xffffffffu, vi, off);
            if (ov > v || (ov == v && oi < vi)) { v = ov; vi = oi; }
        }
        if (threadIdx.x == 0 && v > neg) flag[vi] = 1;                  // processor.py:252-256
    }
}

__global__ void __launch_bounds__(256) label_kernel(const TargetParams p, const double* __restrict__ best_iou,
                                                    const unsigned char* __restrict__ flag, const unsigned char* __restrict__ pad_mask,
                                                    double* __restrict__ class_maps, double* __restrict__ regress) {
    const long long cells = (long long)p.vsy * p.vsx * p.nt;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    double cls = -1.0;
    if (p.ng > 0) {
        const double b = best_iou[c];
        if (flag[c] || b >= p.pos) cls = 1.0;                           // processor.py:252-260
        else if (p.neg <= b && b < p.pos) cls = 0.0;                    // gray zone, :262-268
    }
    if (pad_mask && pad_mask[c] && cls != -1.0) {                       // :271-273 (only the tx channel is cleared)
        cls = 0.0;
        const int t = (int)(c % p.nt);
        regress[(c / p.nt) * 4 * p.nt + t] = 0.0;
    }
    class_maps[c] = cls;
}

}  // namespace

TF_API int tf_targets_workspace_bytes(int vsy, int vsx, int nt, int ng, size_t* bytes) {
    TF_REQUIRE(bytes && vsy > 0 && vsx > 0 && nt > 0 && nt <= MAX_TPL && ng >= 0, "tf_targets_workspace_bytes: bad args");
    const size_t cells = (size_t)vsy * vsx * nt;
    *bytes = tf_align_up(cells * (size_t)std::max(ng, 1) * 8, 256) + tf_align_up(cells * 8, 256) + tf_align_up(cells, 256) + 1024;
    return TF_OK;
}

// bboxes: DEVICE [ng,4] float64 (x1,y1,x2,y2; already filtered: x2 > x1, y2 > y1, processor.py:236-240); templates_host:
// HOST [nt,4]; rf offset / stride as in the reference (y first); jitter: DEVICE [vsy,vsx,nt,ng] float64 in [0,1) -- the
// np.random.rand draws of processor.py:203 -- or NULL for device noise from `seed`; pad_mask: DEVICE uint8 [vsy,vsx,nt] or NULL.
// Outputs (DEVICE float64): class_maps [vsy,vsx,nt], regress_maps [vsy,vsx,4*nt], iou_out [vsy,vsx,nt,ng] (optional; the
// PERTURBED volume, which is what get_heatmaps returns).
TF_API int tf_heatmap_targets(const double* bboxes, int ng, const double* templates_host, int nt, int vsy, int vsx, int ofy, int ofx,
                              int sty, int stx, double pos_thresh, double neg_thresh, const double* jitter, uint64_t seed,
                              const uint8_t* pad_mask, double* class_maps, double* regress_maps, double* iou_out, void* workspace,
                              size_t workspace_bytes, void* stream) {
    TF_REQUIRE(templates_host && class_maps && regress_maps && workspace, "tf_heatmap_targets: null pointer");
    TF_REQUIRE(vsy > 0 && vsx > 0 && nt > 0 && nt <= MAX_TPL && ng >= 0 && (ng == 0 || bboxes), "tf_heatmap_targets: bad shape");
    size_t need;
    tf_targets_workspace_bytes(vsy, vsx, nt, ng, &need);
    if (workspace_bytes < need) { tf_set_error("tf_heatmap_targets: workspace %zu < %zu", workspace_bytes, need); return TF_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    TargetParams p;
    p.vsy = vsy; p.vsx = vsx; p.nt = nt; p.ng = ng; p.ofy = ofy; p.ofx = ofx; p.sty = sty; p.stx = stx; p.pos = pos_thresh; p.neg = neg_thresh;
    for (int t = 0; t < nt; ++t) { p.dx1[t] = templates_host[4 * t]; p.dy1[t] = templates_host[4 * t + 1]; p.dx2[t] = templates_host[4 * t + 2]; p.dy2[t] = templates_host[4 * t + 3]; }
    const long long cells = (long long)vsy * vsx * nt;
    TfArena ar(workspace, workspace_bytes);
    double* iou_p = ar.take<double>((size_t)cells * std::max(ng, 1));
    double* best = ar.take<double>(cells);
    unsigned char* flag = ar.take<unsigned char>(cells);
    const int blocks = (int)((cells + 255) / 256);
    TF_CHECK_CUDA(cudaMemsetAsync(regress_maps, 0, (size_t)vsy * vsx * 4 * nt * sizeof(double), st));
    if (ng > 0) {
        TF_CHECK_CUDA(cudaMemsetAsync(flag, 0, (size_t)cells, st));
        iou_kernel<<<blocks, 256, 0, st>>>(p, bboxes, jitter, seed, nullptr, iou_p, best, regress_maps);
        object_best_kernel<<<ng, 1024, 0, st>>>(iou_p, cells, ng, neg_thresh, flag);
        if (iou_out) TF_CHECK_CUDA(cudaMemcpyAsync(iou_out, iou_p, (size_t)cells * ng * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    label_kernel<<<blocks, 256, 0, st>>>(p, best, flag, pad_mask, class_maps, regress_maps);
    TF_LAUNCH_CHECK();
    return TF_OK;
}
