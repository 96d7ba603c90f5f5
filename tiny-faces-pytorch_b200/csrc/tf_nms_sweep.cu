// Near-linear exact greedy NMS for large N (same result as the sequential algorithm of torchvision.ops.nms,
// the op the reference calls at /root/reference/tinyfaces/evaluation.py:84; bit-identical keep indices).
//
//   1. stable descending radix sort by score -> rank order (ties: lower original index first)
//   2. sort-and-sweep along x: boxes are sorted by x1; box a only meets boxes whose x1 lies in [x1_a, x2_a], so
//      the number of exact IoU tests is sum_a #{b : x1_a <= x1_b <= x2_a} instead of N^2/2
//   3. every conflicting pair (IoU > thr, evaluated with the reference's operation order) becomes an edge
//      (earlier rank, later rank) appended to ONE unordered list by warp-aggregated atomics -- a single sweep; the
//      first version ran the sweep twice (count pass, scan, CSR fill pass) and the IoU tests are the cost
//   4. greedy resolution as a monotone fixed point: a box is KEPT once all its earlier conflicting boxes are
//      REMOVED, REMOVED once any of them is KEPT.  One round = an edge pass (an edge whose earlier box is KEPT removes
//      the later one; an edge whose earlier box is still undecided blocks it) + a node pass (undecided and not
//      blocked -> KEPT).  Every round decides at least the first undecided box and in practice the dependency
//      chains are short, so a few rounds settle all N boxes in parallel.
//   5. order-preserving compaction of the kept ranks -> original indices, descending score.
//
// Greedy NMS is defined by exactly this recurrence (keep(j) <=> no kept i < j with IoU(i, j) > thr), so the
// fixed point equals the sequential answer.  Used for thr >= 0 (a negative threshold makes every pair conflict)
// when the edge list fits the workspace; otherwise tf_nms falls back to the blocked bit-matrix path.
#include "tf_common.cuh"
#include "tf_nms_common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <algorithm>

using namespace tfnms;

namespace {

enum : unsigned char { UNDECIDED = 0, KEPT = 1, REMOVED = 2 };

__global__ void iota2_kernel(int* v, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}
template <typename T>
__global__ void gather_rank_kernel(const T* __restrict__ boxes, const int* __restrict__ order, int n,
                                   Box<T>* __restrict__ sb, T* __restrict__ area, T* __restrict__ xkey) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Box<T> b = reinterpret_cast<const Box<T>*>(boxes)[order[i]];
    sb[i] = b;
    area[i] = Arith<T>::mul(Arith<T>::sub(b.x2, b.x1), Arith<T>::sub(b.y2, b.y1));
    xkey[i] = b.x1;
}
// x-ordered copies of the boxes so that the sweep streams them instead of gathering through xorder
template <typename T>
__global__ void gather_x_kernel(const Box<T>* __restrict__ sb, const T* __restrict__ area, const int* __restrict__ xorder, int n,
                                Box<T>* __restrict__ xb, T* __restrict__ xarea) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = xorder[i];
    xb[i] = sb[r];
    xarea[i] = area[r];
}
// One WARP per box a (position p in x order): the 32 lanes test 32 consecutive x-successors per step, so dense
// inputs (real detections overlap thousands of x-neighbours) stay parallel.  Conflicts go to the edge list as
// (earlier rank, later rank); entries past the capacity are counted but not stored (the host then falls back).
template <typename T>
__global__ void __launch_bounds__(256) sweep_kernel(const Box<T>* __restrict__ xb, const T* __restrict__ xarea,
                                                    const int* __restrict__ xorder, int n, double thr,
                                                    unsigned long long* __restrict__ nedges, int2* __restrict__ edges,
                                                    unsigned long long cap) {
    const int p = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (p >= n) return;
    const int a = xorder[p];                       // rank (score order) of this box
    const Box<T> A = xb[p];
    const T aa = xarea[p];
    for (int base = p + 1; base < n; base += 32) {
        const int q = base + lane;
        bool live = false, hit = false;
        int b = 0;
        if (q < n) {
            const Box<T> Bx = xb[q];
            live = Bx.x1 <= A.x2;                  // x order: once this fails, it fails for every later q
            if (live && Bx.y1 < A.y2 && A.y1 < Bx.y2) {
                b = xorder[q];
                hit = a < b ? suppresses<T>(A, aa, Bx, xarea[q], thr, true) : suppresses<T>(Bx, xarea[q], A, aa, thr, true);
            }
        }
        const unsigned int m = __ballot_sync(0xffffffffu, hit);
        if (m) {
            unsigned long long start = 0;
            if (lane == 0) start = atomicAdd(nedges, (unsigned long long)__popc(m));
            start = __shfl_sync(0xffffffffu, start, 0);
            const unsigned long long e = start + __popc(m & ((1u << lane) - 1u));
            if (hit && e < cap) edges[e] = a < b ? make_int2(a, b) : make_int2(b, a);
        }
        if (!__any_sync(0xffffffffu, live)) break;
        if (!__shfl_sync(0xffffffffu, (int)live, 31) ) break;      // lane 31 past the x range: so is everything after
    }
}
// one relaxation round = edge pass + node pass; states only move UNDECIDED -> KEPT / REMOVED (both final), so racing /
// stale reads only ever delay a decision to the next round
__global__ void edge_pass_kernel(const int2* __restrict__ edges, unsigned long long nedges, volatile unsigned char* state,
                                 unsigned char* __restrict__ blocked) {
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < nedges;
         e += (unsigned long long)gridDim.x * blockDim.x) {
        const int2 ed = edges[e];                  // x = earlier rank, y = later rank
        if (state[ed.y] != UNDECIDED) continue;
        const unsigned char si = state[ed.x];
        if (si == KEPT) state[ed.y] = REMOVED;
        else if (si == UNDECIDED) blocked[ed.y] = 1;
    }
}
__global__ void node_pass_kernel(volatile unsigned char* state, unsigned char* __restrict__ blocked, int n,
                                 int* __restrict__ undecided) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    bool still = false;
    if (j < n && state[j] == UNDECIDED) {
        if (!blocked[j]) state[j] = KEPT;          // every earlier conflicting box was REMOVED (or there is none)
        else { blocked[j] = 0; still = true; }
    }
    if (undecided) {
        const unsigned int m = __ballot_sync(0xffffffffu, still);
        if (m && (threadIdx.x & 31) == 0) atomicAdd(undecided, __popc(m));
    }
}
__global__ void flags_kernel(const unsigned char* __restrict__ state, int n, unsigned char* __restrict__ flags) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) flags[j] = state[j] == KEPT ? 1 : 0;
}
__global__ void widen_kernel(const int* __restrict__ sel, const int* __restrict__ nsel, long long* __restrict__ keep,
                             long long* __restrict__ num_keep, int n) {
    const int k = *nsel;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) keep[i] = sel[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) *num_keep = k;
}

template <typename T>
struct SweepPlan {
    size_t sort_bytes = 0, select_bytes = 0, total = 0;
    size_t edge_cap = 0;
    explicit SweepPlan(int64_t n) {
        cub::DeviceRadixSort::SortPairsDescending(nullptr, sort_bytes, (const T*)nullptr, (T*)nullptr, (const int*)nullptr, (int*)nullptr, (int)n);
        size_t s2 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, s2, (const T*)nullptr, (T*)nullptr, (const int*)nullptr, (int*)nullptr, (int)n);
        sort_bytes = sort_bytes > s2 ? sort_bytes : s2;
        cub::DeviceSelect::Flagged(nullptr, select_bytes, (const int*)nullptr, (const unsigned char*)nullptr, (int*)nullptr, (int*)nullptr, (int)n);
        edge_cap = (size_t)n * 128 + (1u << 19);                         // (earlier, later) pairs
        if (edge_cap > 0x70000000ull) edge_cap = 0x70000000ull;
        size_t a = 0;
        auto add = [&](size_t b) { a = tf_align_up(a, 256) + b; };
        add(4 * n); add(4 * n); add(sizeof(T) * n);                      // iota, order, sorted keys
        add(sort_bytes); add(select_bytes);
        add(sizeof(Box<T>) * n); add(sizeof(T) * n);                      // boxes / areas by rank
        add(sizeof(T) * n); add(sizeof(T) * n); add(4 * n);               // x keys, sorted x keys, x order
        add(sizeof(Box<T>) * n); add(sizeof(T) * n);                      // boxes / areas in x order
        add(8 * edge_cap);                                                // edge list
        add(n); add(n); add(n); add(4 * n);                               // state, blocked, flags, selected
        add(256);
        total = a + 256;
    }
};

}  // namespace

namespace tfnms {

size_t sweep_workspace_bytes(int64_t n, int elem_bytes) {
    return elem_bytes == 8 ? SweepPlan<double>(n).total : SweepPlan<float>(n).total;
}

// returns TF_OK, an error, or +1 when the edge list does not fit (caller falls back to the bit-matrix path)
template <typename T>
int run_nms_sweep(const void* boxes, const void* scores, int64_t n64, double thr, long long* keep, long long* num_keep,
                  void* ws, size_t ws_bytes, cudaStream_t st) {
    const int n = (int)n64;
    SweepPlan<T> plan(n);
    if (ws_bytes < plan.total) { tf_set_error("tf_nms(sweep): workspace %zu < required %zu", ws_bytes, plan.total); return TF_ERR_WORKSPACE; }
    TfArena ar(ws, ws_bytes);
    int* iota = ar.take<int>(n);
    int* order = ar.take<int>(n);
    T* keys = ar.take<T>(n);
    void* sort_tmp = ar.take<char>(plan.sort_bytes);
    void* select_tmp = ar.take<char>(plan.select_bytes);
    Box<T>* sb = ar.take<Box<T>>(n);
    T* area = ar.take<T>(n);
    T* xkey = ar.take<T>(n);
    T* xsorted = ar.take<T>(n);
    int* xorder = ar.take<int>(n);
    Box<T>* xb = ar.take<Box<T>>(n);
    T* xarea = ar.take<T>(n);
    int2* edges = ar.take<int2>(plan.edge_cap);
    unsigned char* state = ar.take<unsigned char>(n);
    unsigned char* blocked = ar.take<unsigned char>(n);
    unsigned char* flags = ar.take<unsigned char>(n);
    int* selected = ar.take<int>(n);
    int* scalars = ar.take<int>(16);                       // [0] undecided, [1] selected count, [2..3] edge counter (u64)
    unsigned long long* nedges_dev = reinterpret_cast<unsigned long long*>(scalars + 2);
    const int nb = (n + 255) / 256;

    iota2_kernel<<<nb, 256, 0, st>>>(iota, n);
    size_t sb_bytes = plan.sort_bytes;
    TF_CHECK_CUDA(cub::DeviceRadixSort::SortPairsDescending(sort_tmp, sb_bytes, (const T*)scores, keys, (const int*)iota, order, n,
                                                            0, (int)sizeof(T) * 8, st));
    gather_rank_kernel<T><<<nb, 256, 0, st>>>((const T*)boxes, order, n, sb, area, xkey);
    sb_bytes = plan.sort_bytes;
    TF_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(sort_tmp, sb_bytes, (const T*)xkey, xsorted, (const int*)iota, xorder, n, 0,
                                                  (int)sizeof(T) * 8, st));
    gather_x_kernel<T><<<nb, 256, 0, st>>>(sb, area, xorder, n, xb, xarea);
    TF_CHECK_CUDA(cudaMemsetAsync(scalars, 0, 16 * sizeof(int), st));
    TF_CHECK_CUDA(cudaMemsetAsync(state, 0, (size_t)n, st));
    TF_CHECK_CUDA(cudaMemsetAsync(blocked, 0, (size_t)n, st));
    const int sweep_blocks = (int)(((long long)n * 32 + 255) / 256);
    sweep_kernel<T><<<sweep_blocks, 256, 0, st>>>(xb, xarea, xorder, n, thr, nedges_dev, edges, (unsigned long long)plan.edge_cap);
    unsigned long long total_edges = 0;
    TF_CHECK_CUDA(cudaMemcpyAsync(&total_edges, nedges_dev, 8, cudaMemcpyDeviceToHost, st));
    TF_CHECK_CUDA(cudaStreamSynchronize(st));
    if (total_edges > (unsigned long long)plan.edge_cap) return 1;
    const int eb = (int)std::min<unsigned long long>((total_edges + 255) / 256, 148ull * 16);
    for (int round = 0; round < n + 8;) {
        TF_CHECK_CUDA(cudaMemsetAsync(scalars, 0, 4, st));
        for (int k = 0; k < 4; ++k, ++round) {
            if (total_edges) edge_pass_kernel<<<eb, 256, 0, st>>>(edges, total_edges, state, blocked);
            node_pass_kernel<<<nb, 256, 0, st>>>(state, blocked, n, k == 3 ? scalars : nullptr);
        }
        int undecided = 0;
        TF_CHECK_CUDA(cudaMemcpyAsync(&undecided, scalars, 4, cudaMemcpyDeviceToHost, st));
        TF_CHECK_CUDA(cudaStreamSynchronize(st));
        if (undecided == 0) break;
    }
    flags_kernel<<<nb, 256, 0, st>>>(state, n, flags);
    size_t sel_bytes = plan.select_bytes;
    TF_CHECK_CUDA(cub::DeviceSelect::Flagged(select_tmp, sel_bytes, (const int*)order, flags, selected, scalars + 1, n, st));
    widen_kernel<<<nb < 1024 ? nb : 1024, 256, 0, st>>>(selected, scalars + 1, keep, num_keep, n);
    TF_LAUNCH_CHECK();
    return TF_OK;
}

template int run_nms_sweep<double>(const void*, const void*, int64_t, double, long long*, long long*, void*, size_t, cudaStream_t);
template int run_nms_sweep<float>(const void*, const void*, int64_t, double, long long*, long long*, void*, size_t, cudaStream_t);

}  // namespace tfnms
