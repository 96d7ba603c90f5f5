/* Plain-C restatement of greedy NMS as torchvision.ops.nms computes it on CPU
 * (TEST INFRASTRUCTURE ONLY -- never linked into the product library).
 *
 * The reference calls torchvision.ops.nms at /root/reference/tinyfaces/
 * evaluation.py:84 on float64 CPU tensors.  torchvision's C++ source is not
 * vendored under /root/reference (pinned torchvision ^0.18.0, pyproject.toml:17;
 * installed 0.26.0); its published algorithm is restated here and pinned against
 * the installed op by tests/golden/nms_*.npz (oracle/make_golden.py) and
 * tests/test_oracle_golden.py:
 *   - stable sort by score, descending (ties: lower index first), in torch.sort's order: every NaN is the
 *     largest value and NaNs tie with each other; -0.0 == +0.0
 *   - area = (x2-x1)*(y2-y1)   (no +1)
 *   - suppress j iff inter / (area_i + area_j - inter) > thr   (strict; NaN never)
 *   - keep indices returned in descending-score order, int64
 * Compile: see oracle/Makefile (gcc -O2 -ffp-contract=off -shared -fPIC).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int score_gt(double x, double y) {            /* torch.sort(descending=True): NaN > everything, NaN == NaN */
    if (x != x) return y == y;
    return x > y;
}

static void merge_sort_desc(const double *s, int64_t *idx, int64_t *tmp, int64_t n) {
    /* bottom-up stable merge sort on indices, descending by s[] */
    for (int64_t w = 1; w < n; w *= 2) {
        for (int64_t lo = 0; lo < n; lo += 2 * w) {
            int64_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            int64_t a = lo, b = mid, o = lo;
            while (a < mid && b < hi) tmp[o++] = score_gt(s[idx[b]], s[idx[a]]) ? idx[b++] : idx[a++];
            while (a < mid) tmp[o++] = idx[a++];
            while (b < hi) tmp[o++] = idx[b++];
        }
        memcpy(idx, tmp, (size_t)n * sizeof(int64_t));
    }
}

/* returns number kept; keep[] must hold n entries */
int64_t tf_oracle_nms_f64(const double *boxes, const double *scores, int64_t n, double thr,
                          int64_t *keep) {
    if (n <= 0) return 0;
    int64_t *order = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    int64_t *tmp = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    double *area = (double *)malloc((size_t)n * sizeof(double));
    unsigned char *dead = (unsigned char *)calloc((size_t)n, 1);
    for (int64_t i = 0; i < n; ++i) {
        order[i] = i;
        area[i] = (boxes[4 * i + 2] - boxes[4 * i]) * (boxes[4 * i + 3] - boxes[4 * i + 1]);
    }
    merge_sort_desc(scores, order, tmp, n);
    int64_t k = 0;
    for (int64_t a = 0; a < n; ++a) {
        int64_t i = order[a];
        if (dead[i]) continue;
        keep[k++] = i;
        double ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
        double ia = area[i];
        for (int64_t b = a + 1; b < n; ++b) {
            int64_t j = order[b];
            if (dead[j]) continue;
            /* std::max(a, b) = (a < b) ? b : a; std::min(a, b) = (b < a) ? b : a -- NaN coordinates propagate like this */
            double xx1 = ix1 < boxes[4 * j] ? boxes[4 * j] : ix1;
            double yy1 = iy1 < boxes[4 * j + 1] ? boxes[4 * j + 1] : iy1;
            double xx2 = boxes[4 * j + 2] < ix2 ? boxes[4 * j + 2] : ix2;
            double yy2 = boxes[4 * j + 3] < iy2 ? boxes[4 * j + 3] : iy2;
            double w = xx2 - xx1, h = yy2 - yy1;
            w = 0 < w ? w : 0;
            h = 0 < h ? h : 0;
            double inter = w * h;
            double ovr = inter / (ia + area[j] - inter);
            if (ovr > thr) dead[j] = 1;
        }
    }
    free(order); free(tmp); free(area); free(dead);
    return k;
}
