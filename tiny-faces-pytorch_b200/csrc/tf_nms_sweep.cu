// Near-linear exact greedy NMS for large N (same result as the sequential algorithm of torchvision.ops.nms,
// the op the reference calls at /root/reference/tinyfaces/evaluation.py:84; bit-identical keep indices).
//
//   1. stable descending radix sort by score (canonical keys: torch.sort's NaN / -0.0 order) -> rank order
//   2. sort-and-sweep along x: boxes are sorted by a monotone float32 image of x1 (4 radix passes instead of 8 -- the x
//      order only prunes, it never decides); box a meets the boxes q behind it while f(x1_q) <= f(x2_a), a superset of
//      x1_q <= x2_a, so every pair that can overlap is visited exactly once
//   3. every conflicting pair (IoU > thr, evaluated with the reference's operation order) becomes an edge
//      (earlier rank, later rank) appended to ONE unordered list by warp-aggregated atomics
//   4. greedy resolution as a monotone fixed point: a box is KEPT once all its earlier conflicting boxes are
//      REMOVED, REMOVED once any of them is KEPT.  ONE cooperative kernel iterates rounds of (edge pass, grid sync, node
//      pass, grid sync) until no box is undecided; the edge pass compacts the list as it goes (an edge whose earlier box
//      is decided can never matter again), so late rounds touch only the few long dependency chains.
//   5. order-preserving compaction of the kept ranks -> original indices, descending score.
//
// Everything is enqueued on the caller's stream: no host synchronisation, no host read-back (round 1 polled the
// undecided count from the host every 4 rounds and read the edge count back).  If the edge list
// does not fit the workspace the result is flagged on the device (num_keep = -1) and the caller re-runs the blocked
// bit-matrix algorithm -- the host learns it when it reads the count, which it must do anyway to use `keep`.
//
// Greedy NMS is defined by exactly this recurrence (keep(j) <=> no kept i < j with IoU(i, j) > thr), so the
// fixed point equals the sequential answer.  Used for thr >= 0 (a negative threshold makes every pair conflict).
#include "tf_common.cuh"
#include "tf_nms_common.cuh"
#include "tf_conv_gemm.h"
#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <algorithm>

using namespace tfnms;
namespace cg = cooperative_groups;

namespace {

enum : unsigned char { UNDECIDED = 0, KEPT = 1, REMOVED = 2 };

// scalars block (64-bit words, zeroed together with state / blocked by ONE memset):
//   [0] edges emitted by the sweep   [1], [2] edge counts of the ping-pong lists   [3], [4] "some box is still undecided"
//   [5] IoU pair tests (only with tf_debug_set(13, 1))   [6] rounds   [7] selected count (int)
//   [8], [9] ~enc(min x1), ~enc(min y1) of the valid boxes (grid variant)   [10] grid flag: a box size outside the level range
//   [11..44] level_start[0..33] of the grid order
constexpr int SC_EDGES = 0, SC_ECNT = 1, SC_UNDEC = 3, SC_PAIRS = 5, SC_ROUNDS = 6, SC_SEL = 7, SC_XMIN = 8, SC_YMIN = 9, SC_GFLAG = 10,
              SC_LEVEL = 11, SC_WORDS = 48;
constexpr int GRID_LEVELS = 32, GRID_LEVEL_BIAS = 16;          // level = ilogb(max(w, h)) + 16: sizes 2^-16 .. 2^16 px
constexpr unsigned long long GRID_INVALID = 1ull << 37;

__device__ __forceinline__ float xkey_of(double v) { return __double2float_rn(v); }      // monotone non-decreasing
__device__ __forceinline__ float xkey_of(float v) { return v; }
__device__ __forceinline__ float lo_of(double v) { return __double2float_rd(v); }        // <= v
__device__ __forceinline__ float lo_of(float v) { return v; }
__device__ __forceinline__ float hi_of(double v) { return __double2float_ru(v); }        // >= v
__device__ __forceinline__ float hi_of(float v) { return v; }

// 16-byte sweep record of a box in x order: everything the candidate filter needs.  The sweep visits ~N * (boxes per
// x-extent) candidates but only a few percent of them overlap in y: the filter reads this record (one 16 B load) and
// touches the exact 32-byte box only for the survivors.  The float y-range is widened outwards (lo rounded down, hi
// rounded up), so the filter is a superset of the exact test -- exactness is decided on the full-precision box.
struct alignas(16) SweepRec { float x1, ylo, yhi; int rank; };

template <typename T>
__global__ void gather_rank_kernel(const T* __restrict__ boxes, const int* __restrict__ order, int n,
                                   Box<T>* __restrict__ sb, T* __restrict__ area, float* __restrict__ xkey) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Box<T> b = reinterpret_cast<const Box<T>*>(boxes)[order[i]];
    sb[i] = b;
    area[i] = Arith<T>::mul(Arith<T>::sub(b.x2, b.x1), Arith<T>::sub(b.y2, b.y1));
    xkey[i] = xkey_of(b.x1);
}
// x-ordered copies of the boxes so that the sweep streams them instead of gathering through xorder
template <typename T>
__global__ void gather_x_kernel(const Box<T>* __restrict__ sb, const T* __restrict__ area, const int* __restrict__ xorder,
                                const float* __restrict__ xsorted, int n, Box<T>* __restrict__ xb, T* __restrict__ xarea,
                                SweepRec* __restrict__ rec) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = xorder[i];
    const Box<T> b = sb[r];
    xb[i] = b;
    xarea[i] = area[r];
    SweepRec q;
    q.x1 = xsorted[i]; q.ylo = lo_of(b.y1); q.yhi = hi_of(b.y2); q.rank = r;
    rec[i] = q;
}
// One WARP per box a (position p in x order): the 32 lanes test 32 consecutive x-successors per step, so dense
// inputs (real detections overlap thousands of x-neighbours) stay parallel.  Conflicts go to the edge list as
// (earlier rank, later rank); entries past the capacity are counted but not stored (the result is then flagged).
// (A variant that staged the successor records of 8 consecutive boxes through shared memory -- one L2 read per block instead
// of one per warp -- was measured and removed: 0.97 vs 0.85 ms at 10^5 random boxes, 7.4 vs 4.6 ms on dense pyramid candidates:
// the block waits for its slowest warp and pays two barriers per tile, and the L2 is not what limits the per-warp version.)
// One WARP per box a (position p in x order).  Phase 1 (cheap, every lane busy): the 32 lanes filter 32 consecutive
// x-successors per step on the 16-byte records and push the survivors into a per-warp queue in shared memory.  Phase 2
// (expensive, fp64 with a division): whenever 32 survivors are queued, every lane runs one exact IoU test.  Without the
// queue ~2/3 of the steps had one or two lanes in the exact test while the other thirty waited for its latency.
// Conflicts go to the edge list as (earlier rank, later rank); entries past the capacity are counted but not stored.
template <typename T>
__device__ __forceinline__ void sweep_exact(const Box<T>* __restrict__ xb, const T* __restrict__ xarea, const SweepRec* __restrict__ rec,
                                            const Box<T>& A, T aa, int a, int q, bool active, double thr, int lane,
                                            unsigned long long* __restrict__ scalars, int2* __restrict__ edges, unsigned long long cap,
                                            unsigned int& tested) {
    bool hit = false;
    int b = 0;
    if (active) {
        const Box<T> Bx = xb[q];
        if (Bx.y1 < A.y2 && A.y1 < Bx.y2) {
            b = rec[q].rank;
            ++tested;
            hit = a < b ? suppresses<T>(A, aa, Bx, xarea[q], thr, true) : suppresses<T>(Bx, xarea[q], A, aa, thr, true);
        }
    }
    const unsigned int m = __ballot_sync(0xffffffffu, hit);
    if (m) {
        unsigned long long start = 0;
        if (lane == 0) start = atomicAdd(scalars + SC_EDGES, (unsigned long long)__popc(m));
        start = __shfl_sync(0xffffffffu, start, 0);
        const unsigned long long e = start + __popc(m & ((1u << lane) - 1u));
        if (hit && e < cap) edges[e] = a < b ? make_int2(a, b) : make_int2(b, a);
    }
}
template <typename T>
__global__ void __launch_bounds__(256) sweep_kernel(const Box<T>* __restrict__ xb, const T* __restrict__ xarea,
                                                    const SweepRec* __restrict__ rec, int n,
                                                    double thr, unsigned long long* __restrict__ scalars,
                                                    int2* __restrict__ edges, unsigned long long cap, int count_pairs) {
    __shared__ int queue[8][64];
    const int p = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (p >= n) return;                            // (whole warps leave together: p is warp-uniform)
    int* wq = queue[threadIdx.x >> 5];
    const SweepRec ra = rec[p];
    const int a = ra.rank;                         // rank (score order) of this box
    const Box<T> A = xb[p];
    const T aa = xarea[p];
    const float ax2 = xkey_of(A.x2);               // NaN: every comparison fails -> no candidates (such a box never conflicts)
    unsigned int tested = 0;
    int qn = 0;                                    // queued survivors (warp-uniform, < 32 at the top of every step)
    // SWEEP_W * 32 successors per step (SWEEP_W records per lane, all loads issued before any is used): the loop is a chain of
    // L2 round trips -- the next records are only requested once the warp knows that the x range goes on -- so more
    // records per trip is fewer trips (measured at 10^5 boxes: 32 per step 0.45 ms, 64 per step 0.28 ms)
    constexpr int SWEEP_W = 4;
    for (int base = p + 1; base < n; base += 32 * SWEEP_W) {
        SweepRec r[SWEEP_W];
        bool pass[SWEEP_W], live_last = false;
#pragma unroll
        for (int h = 0; h < SWEEP_W; ++h) {
            const int q = base + 32 * h + lane;
            r[h] = SweepRec{0.f, 0.f, 0.f, 0};
            if (q < n) r[h] = rec[q];
        }
#pragma unroll
        for (int h = 0; h < SWEEP_W; ++h) {
            const int q = base + 32 * h + lane;
            const bool live = q < n && r[h].x1 <= ax2;                           // x order: once this fails, it fails for every later q
            pass[h] = live && r[h].ylo < ra.yhi && ra.ylo < r[h].yhi;            // widened float y ranges: a superset of the exact test
            if (h == SWEEP_W - 1) live_last = live;
        }
#pragma unroll
        for (int h = 0; h < SWEEP_W; ++h) {
            const unsigned int m = __ballot_sync(0xffffffffu, pass[h]);
            if (m) {
                if (pass[h]) wq[qn + __popc(m & ((1u << lane) - 1u))] = base + 32 * h + lane;
                qn += __popc(m);
                __syncwarp();
                if (qn >= 32) {
                    const int cand = wq[lane];
                    const int spill = lane + 32 < qn ? wq[lane + 32] : 0;
                    __syncwarp();
                    if (lane + 32 < qn) wq[lane] = spill;                  // move the overflow to the front
                    qn -= 32;
                    __syncwarp();
                    sweep_exact<T>(xb, xarea, rec, A, aa, a, cand, true, thr, lane, scalars, edges, cap, tested);
                }
            }
        }
        if (!__shfl_sync(0xffffffffu, (int)live_last, 31)) break;  // lane 31 of the last group past the x range: so is everything after
    }
    if (qn > 0) {
        const int cand = lane < qn ? wq[lane] : 0;
        sweep_exact<T>(xb, xarea, rec, A, aa, a, cand, lane < qn, thr, lane, scalars, edges, cap, tested);
    }
    if (count_pairs) {
        tested = (unsigned int)tf_warp_sum((int)tested);
        if (lane == 0 && tested) atomicAdd(scalars + SC_PAIRS, (unsigned long long)tested);
    }
}



template <typename T>
__device__ __forceinline__ void grid_exact(const Box<T>* __restrict__ xb, const T* __restrict__ xarea, const int* __restrict__ gorder,
                                           const Box<T>& A, T aa, int a, int q, bool active, double thr, int lane,
                                           unsigned long long* __restrict__ scalars, int2* __restrict__ edges, unsigned long long cap,
                                           unsigned int& tested) {
    bool hit = false;
    int b = 0;
    if (active) {
        const Box<T> Bx = xb[q];
        if (Bx.y1 < A.y2 && A.y1 < Bx.y2 && Bx.x1 < A.x2 && A.x1 < Bx.x2) {
            b = gorder[q];
            ++tested;
            hit = a < b ? suppresses<T>(A, aa, Bx, xarea[q], thr, true) : suppresses<T>(Bx, xarea[q], A, aa, thr, true);
        }
    }
    const unsigned int m = __ballot_sync(0xffffffffu, hit);
    if (m) {
        unsigned long long start = 0;
        if (lane == 0) start = atomicAdd(scalars + SC_EDGES, (unsigned long long)__popc(m));
        start = __shfl_sync(0xffffffffu, start, 0);
        const unsigned long long e = start + __popc(m & ((1u << lane) - 1u));
        if (hit && e < cap) edges[e] = a < b ? make_int2(a, b) : make_int2(b, a);
    }
}

// ------------------------------------------------------------------------------------------------ size-class grid
// Candidate generation for boxes of very different sizes (dense pyramid detections: the 1-D sweep above visits every box
// whose x1 falls into a's x-extent, whatever its y -- for boxes as wide as the image that is a constant fraction of N).
// Boxes are binned by SIZE CLASS l = floor(log2(max(w, h))) and, inside a class, by the grid cell of their (x1, y1) corner
// with cell size g_l = 2^(l+1) > max(w, h).  A box b of class l that overlaps a box a with max(w_a, h_a) <= g_l has
// x1_b in (x1_a - g_l, x2_a), so cell(x1_b) in [cell(x1_a) - 1, cell(x2_a)]: at most 3 x 3 cells of class l.  Box a looks
// into every class >= its own (a pair of different classes is found from the smaller box; a pair of the same class
// from the box that comes first in the sorted order), i.e. every overlapping pair exactly once.  Cell coordinates are
// clamped to 16 bits (monotone, so ranges stay supersets); the exact fp64 IoU test decides, as in the sweep.
struct alignas(16) GridRec { float x1, y1, x2, y2; };          // widened outwards: lo rounded down, hi rounded up

__device__ __forceinline__ unsigned long long enc_order(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dec_order(unsigned long long e) {
    const unsigned long long b = (e >> 63) ? (e & 0x7fffffffffffffffull) : ~e;
    return __longlong_as_double((long long)b);
}
template <typename T>
__device__ __forceinline__ bool grid_valid(const Box<T>& b) {
    const double x1 = (double)b.x1, y1 = (double)b.y1, x2 = (double)b.x2, y2 = (double)b.y2;
    return isfinite(x1) && isfinite(y1) && isfinite(x2) && isfinite(y2) && x2 > x1 && y2 > y1;      // others never overlap positively
}
template <typename T>
__global__ void grid_prep_kernel(const T* __restrict__ boxes, const int* __restrict__ order, int n, Box<T>* __restrict__ sb,
                                 T* __restrict__ area, unsigned long long* __restrict__ scalars) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long ix = 0, iy = 0;                      // ~enc: 0 = "no value"
    if (i < n) {
        const Box<T> b = reinterpret_cast<const Box<T>*>(boxes)[order[i]];
        sb[i] = b;
        area[i] = Arith<T>::mul(Arith<T>::sub(b.x2, b.x1), Arith<T>::sub(b.y2, b.y1));
        if (grid_valid<T>(b)) { ix = ~enc_order((double)b.x1); iy = ~enc_order((double)b.y1); }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ox = __shfl_xor_sync(0xffffffffu, ix, o), oy = __shfl_xor_sync(0xffffffffu, iy, o);
        ix = ox > ix ? ox : ix; iy = oy > iy ? oy : iy;
    }
    if ((threadIdx.x & 31) == 0) {
        if (ix) atomicMax(scalars + SC_XMIN, ix);
        if (iy) atomicMax(scalars + SC_YMIN, iy);
    }
}
__device__ __forceinline__ unsigned long long grid_key(int level, long long cy, long long cx) {
    cy = cy < 0 ? 0 : (cy > 65535 ? 65535 : cy);
    cx = cx < 0 ? 0 : (cx > 65535 ? 65535 : cx);
    return ((unsigned long long)level << 32) | ((unsigned long long)cy << 16) | (unsigned long long)cx;
}
__device__ __forceinline__ long long grid_cell(double v, double vmin, int level) {
    // cell size 2^(level - BIAS + 1): scaling by a power of two is exact, floor and the subtraction are monotone
    const double c = floor(ldexp(v - vmin, -(level - GRID_LEVEL_BIAS + 1)));
    return c < -1.0 ? -1 : (c > 70000.0 ? 70000 : (long long)c);
}
template <typename T>
__global__ void grid_key_kernel(const Box<T>* __restrict__ sb, int n, unsigned long long* __restrict__ scalars,
                                unsigned long long* __restrict__ keys, int* __restrict__ iota) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    iota[i] = i;
    const Box<T> b = sb[i];
    unsigned long long key = GRID_INVALID;
    if (grid_valid<T>(b)) {
        const double w = (double)b.x2 - (double)b.x1, h = (double)b.y2 - (double)b.y1;
        const int level = ilogb(w > h ? w : h) + GRID_LEVEL_BIAS;
        if (level < 0 || level >= GRID_LEVELS) scalars[SC_GFLAG] = 1;          // size outside 2^-16 .. 2^16: result flagged
        else {
            const double xmin = dec_order(~scalars[SC_XMIN]), ymin = dec_order(~scalars[SC_YMIN]);
            key = grid_key(level, grid_cell((double)b.y1, ymin, level), grid_cell((double)b.x1, xmin, level));
        }
    }
    keys[i] = key;
}
// first position whose key is >= target in the sorted keys[lo, hi) -- 32-ary search, all lanes of the warp take part
__device__ __forceinline__ int warp_lower_bound(const unsigned long long* __restrict__ keys, int lo, int hi, unsigned long long target, int lane) {
    while (hi - lo > 32) {
        const int step = (hi - lo) / 32;
        const int probe = lo + (lane + 1) * step - 1;
        const unsigned int m = __ballot_sync(0xffffffffu, keys[probe] < target);
        const int c = __popc(m);                       // sorted: the lanes that answer "less" form a prefix
        const int nlo = lo + c * step;
        if (c < 32) hi = lo + (c + 1) * step;
        lo = nlo;
    }
    const unsigned int m = __ballot_sync(0xffffffffu, lo + lane < hi && keys[lo + lane] < target);
    return lo + __popc(m);
}
__global__ void grid_levels_kernel(const unsigned long long* __restrict__ keys, int n, unsigned long long* __restrict__ scalars) {
    const int lane = threadIdx.x & 31;
    for (int l = 0; l <= GRID_LEVELS + 1; ++l) {
        const unsigned long long target = l <= GRID_LEVELS ? ((unsigned long long)l << 32) : (GRID_INVALID + 1);
        const int pos = warp_lower_bound(keys, 0, n, l == GRID_LEVELS ? GRID_INVALID : target, lane);
        if (lane == 0) scalars[SC_LEVEL + l] = (unsigned long long)pos;      // [GRID_LEVELS] = first invalid box
    }
}
template <typename T>
__global__ void grid_gather_kernel(const Box<T>* __restrict__ sb, const T* __restrict__ area, const int* __restrict__ gorder, int n,
                                   Box<T>* __restrict__ xb, T* __restrict__ xarea, GridRec* __restrict__ rec) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = gorder[i];
    const Box<T> b = sb[r];
    xb[i] = b;
    xarea[i] = area[r];
    GridRec q;
    q.x1 = lo_of(b.x1); q.y1 = lo_of(b.y1); q.x2 = hi_of(b.x2); q.y2 = hi_of(b.y2);
    rec[i] = q;
}
template <typename T>
__global__ void __launch_bounds__(256) grid_sweep_kernel(const Box<T>* __restrict__ xb, const T* __restrict__ xarea,
                                                         const GridRec* __restrict__ rec, const int* __restrict__ gorder,
                                                         const unsigned long long* __restrict__ keys, int n, double thr,
                                                         unsigned long long* __restrict__ scalars, int2* __restrict__ edges,
                                                         unsigned long long cap, int count_pairs) {
    __shared__ int queue[8][64];
    const int p = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (p >= n) return;
    const unsigned long long mykey = keys[p];
    if (mykey >= GRID_INVALID) return;                 // (warp-uniform) boxes that can never overlap positively
    int* wq = queue[threadIdx.x >> 5];
    const int a = gorder[p];
    const Box<T> A = xb[p];
    const T aa = xarea[p];
    const GridRec ra = rec[p];
    const double xmin = dec_order(~scalars[SC_XMIN]), ymin = dec_order(~scalars[SC_YMIN]);
    const int la = (int)(mykey >> 32);
    unsigned int tested = 0;
    int qn = 0;
    for (int l = la; l < GRID_LEVELS; ++l) {
        const int lv0 = (int)scalars[SC_LEVEL + l], lv1 = (int)scalars[SC_LEVEL + l + 1];
        if (lv1 <= lv0) continue;                      // empty size class
        const long long cx0 = grid_cell((double)A.x1, xmin, l) - 1, cx1 = grid_cell((double)A.x2, xmin, l);
        const long long cy0 = grid_cell((double)A.y1, ymin, l) - 1, cy1 = grid_cell((double)A.y2, ymin, l);
        const long long r0 = cy0 < 0 ? 0 : (cy0 > 65535 ? 65535 : cy0), r1 = cy1 > 65535 ? 65535 : cy1;      // (clamped like the keys)
        for (long long r = r0; r <= r1; ++r) {
            const unsigned long long klo = grid_key(l, r, cx0), khi = grid_key(l, r, cx1);
            int q0 = warp_lower_bound(keys, lv0, lv1, klo, lane);
            if (l == la && q0 <= p) q0 = p + 1;        // same class: only boxes behind a in the sorted order
            for (int base = q0; base < lv1; base += 32) {
                const int q = base + lane;
                bool live = false, pass = false;
                if (q < lv1) {
                    live = keys[q] <= khi;             // sorted: once this fails it fails for every later q of the class
                    if (live) {
                        const GridRec rq = rec[q];
                        pass = rq.x1 < ra.x2 && ra.x1 < rq.x2 && rq.y1 < ra.y2 && ra.y1 < rq.y2;
                    }
                }
                const unsigned int m = __ballot_sync(0xffffffffu, pass);
                if (m) {
                    if (pass) wq[qn + __popc(m & ((1u << lane) - 1u))] = q;
                    qn += __popc(m);
                    __syncwarp();
                    if (qn >= 32) {
                        const int cand = wq[lane];
                        const int spill = lane + 32 < qn ? wq[lane + 32] : 0;
                        __syncwarp();
                        if (lane + 32 < qn) wq[lane] = spill;
                        qn -= 32;
                        __syncwarp();
                        grid_exact<T>(xb, xarea, gorder, A, aa, a, cand, true, thr, lane, scalars, edges, cap, tested);
                    }
                }
                if (!__shfl_sync(0xffffffffu, (int)live, 31)) break;
            }
        }
    }
    if (qn > 0) {
        const int cand = lane < qn ? wq[lane] : 0;
        grid_exact<T>(xb, xarea, gorder, A, aa, a, cand, lane < qn, thr, lane, scalars, edges, cap, tested);
    }
    if (count_pairs) {
        tested = (unsigned int)tf_warp_sum((int)tested);
        if (lane == 0 && tested) atomicAdd(scalars + SC_PAIRS, (unsigned long long)tested);
    }
}

// Greedy resolution, all rounds in one cooperative kernel.  States only move UNDECIDED -> KEPT / REMOVED (both final), so
// racing / stale reads inside a pass only ever delay a decision to the next round.
__global__ void __launch_bounds__(256) resolve_kernel(int2* e0, int2* e1, unsigned long long cap, volatile unsigned char* state,
                                                      volatile unsigned char* blocked, int n,
                                                      volatile unsigned long long* scalars) {
    cg::grid_group grid = cg::this_grid();
    unsigned long long ne = scalars[SC_EDGES];
    if (ne > cap || scalars[SC_GFLAG] != 0) return;              // overflow: flagged by finish_kernel (grid-uniform exit)
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    int2* src = e0; int2* dst = e1;
    int round = 0;
    for (;; ++round) {
        const int par = round & 1;
        // ---- edge pass (+ compaction into the other list)
        unsigned long long* out_cnt = const_cast<unsigned long long*>(scalars) + SC_ECNT + par;
        const unsigned long long ne_up = (ne + 31) & ~31ull;                 // whole warps iterate together
        for (unsigned long long e = tid; e < ne_up; e += stride) {
            bool keep_edge = false;
            int2 ed = make_int2(0, 0);
            if (e < ne) {
                ed = src[e];                                                  // x = earlier rank, y = later rank
                if (state[ed.y] == UNDECIDED) {
                    const unsigned char sx = state[ed.x];
                    if (sx == KEPT) state[ed.y] = REMOVED;
                    else if (sx == UNDECIDED) { blocked[ed.y] = 1; keep_edge = true; }
                }
            }
            const unsigned int m = __ballot_sync(0xffffffffu, keep_edge);
            if (m) {
                unsigned long long start = 0;
                if (lane == 0) start = atomicAdd(out_cnt, (unsigned long long)__popc(m));
                start = __shfl_sync(0xffffffffu, start, 0);
                if (keep_edge) dst[start + __popc(m & ((1u << lane) - 1u))] = ed;
            }
        }
        grid.sync();
        // ---- node pass
        bool still = false;
        for (unsigned long long j = tid; j < (unsigned long long)n; j += stride) {
            if (state[j] == UNDECIDED) {
                if (!blocked[j]) state[j] = KEPT;          // every earlier conflicting box was REMOVED (or there is none)
                else { blocked[j] = 0; still = true; }
            }
        }
        if (still) scalars[SC_UNDEC + par] = 1;
        if (tid == 0) { scalars[SC_ECNT + (par ^ 1)] = 0; scalars[SC_UNDEC + (par ^ 1)] = 0; }     // next round's outputs
        grid.sync();
        if (!scalars[SC_UNDEC + par]) break;
        ne = scalars[SC_ECNT + par];
        int2* t = src; src = dst; dst = t;
    }
    if (tid == 0) scalars[SC_ROUNDS] = (unsigned long long)(round + 1);
}
__global__ void flags_kernel(const unsigned char* __restrict__ state, int n, unsigned char* __restrict__ flags) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) flags[j] = state[j] == KEPT ? 1 : 0;
}
__global__ void finish_kernel(const int* __restrict__ sel, const unsigned long long* __restrict__ scalars, unsigned long long cap,
                              long long* __restrict__ keep, long long* __restrict__ num_keep) {
    const bool overflow = scalars[SC_EDGES] > cap || scalars[SC_GFLAG] != 0;   // edge list too small, or a box size outside the grid's classes
    const int k = overflow ? 0 : *reinterpret_cast<const int*>(scalars + SC_SEL);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) keep[i] = sel[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) *num_keep = overflow ? -1 : k;
}

template <typename T>
struct SweepPlan {
    size_t sort_bytes = 0, select_bytes = 0, total = 0;
    size_t edge_cap = 0, zero_bytes = 0;
    SweepPlan(int64_t n, size_t ws_bytes) {
        cub::DeviceRadixSort::SortPairsDescending(nullptr, sort_bytes, (const T*)nullptr, (T*)nullptr, (const int*)nullptr, (int*)nullptr, (int)n);
        size_t s2 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, s2, (const float*)nullptr, (float*)nullptr, (const int*)nullptr, (int*)nullptr, (int)n);
        sort_bytes = sort_bytes > s2 ? sort_bytes : s2;
        cub::DeviceRadixSort::SortPairs(nullptr, s2, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const int*)nullptr, (int*)nullptr, (int)n);
        sort_bytes = sort_bytes > s2 ? sort_bytes : s2;
        cub::DeviceSelect::Flagged(nullptr, select_bytes, (const int*)nullptr, (const unsigned char*)nullptr, (int*)nullptr, (int*)nullptr, (int)n);
        zero_bytes = tf_align_up((size_t)n, 256) * 2 + SC_WORDS * 8;     // state | blocked | scalars
        size_t a = 0;
        auto add = [&](size_t b) { a = tf_align_up(a, 256) + b; };
        add(4 * n); add(4 * n); add(sizeof(T) * n); add(sizeof(T) * n);   // iota, order, canonical keys, sorted keys
        add(sort_bytes); add(select_bytes);
        add(sizeof(Box<T>) * n); add(sizeof(T) * n);                      // boxes / areas by rank
        add(4 * n); add(4 * n); add(4 * n);                               // x keys (float), sorted x keys, x order
        add(sizeof(Box<T>) * n); add(sizeof(T) * n); add(16 * n);         // boxes / areas / sweep records in x (or grid) order
        add(8 * n); add(8 * n);                                           // grid keys, sorted grid keys
        add(zero_bytes); add(n); add(4 * n);                              // state | blocked | scalars, flags, selected
        const size_t fixed = a + 1024;
        // two ping-pong edge lists: the default is 128 conflicts per box; a larger caller workspace buys a larger list
        edge_cap = (size_t)n * 128 + (1u << 19);
        if (ws_bytes > fixed + 2 * 8 * edge_cap) edge_cap = (ws_bytes - fixed) / 16;
        if (edge_cap > 0x70000000ull) edge_cap = 0x70000000ull;
        total = fixed + 2 * 8 * edge_cap;
    }
};

int g_resolve_grid = 0;
int resolve_grid() {
    if (g_resolve_grid) return g_resolve_grid;
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, resolve_kernel, 256, 0) != cudaSuccess || per_sm < 1)
        return 0;
    g_resolve_grid = sms * (per_sm < 4 ? per_sm : 4);
    return g_resolve_grid;
}

}  // namespace

namespace tfnms {

size_t sweep_workspace_bytes(int64_t n, int elem_bytes) {
    return elem_bytes == 8 ? SweepPlan<double>(n, 0).total : SweepPlan<float>(n, 0).total;
}

// Enqueues the whole algorithm on `st`; an edge-list overflow is reported through *num_keep = -1 on the device.
// stop_after (test / bench hook, tf_debug_set(14, k)): 1 = sorts + gathers only, 2 = + sweep, 3 = + resolution.
template <typename T>
int run_nms_sweep(const void* boxes, const void* scores, int64_t n64, double thr, long long* keep, long long* num_keep,
                  void* ws, size_t ws_bytes, cudaStream_t st, int use_grid) {
    const int n = (int)n64;
    SweepPlan<T> plan(n, ws_bytes);
    if (ws_bytes < plan.total) { tf_set_error("tf_nms(sweep): workspace %zu < required %zu", ws_bytes, plan.total); return TF_ERR_WORKSPACE; }
    TfArena ar(ws, ws_bytes);
    int* iota = ar.take<int>(n);
    int* order = ar.take<int>(n);
    T* keys_in = ar.take<T>(n);
    T* keys = ar.take<T>(n);
    void* sort_tmp = ar.take<char>(plan.sort_bytes);
    void* select_tmp = ar.take<char>(plan.select_bytes);
    Box<T>* sb = ar.take<Box<T>>(n);
    T* area = ar.take<T>(n);
    float* xkey = ar.take<float>(n);
    float* xsorted = ar.take<float>(n);
    int* xorder = ar.take<int>(n);
    Box<T>* xb = ar.take<Box<T>>(n);
    T* xarea = ar.take<T>(n);
    SweepRec* rec = ar.take<SweepRec>(n);
    unsigned long long* gkey_in = ar.take<unsigned long long>(n);
    unsigned long long* gkey = ar.take<unsigned long long>(n);
    unsigned char* zero = ar.take<unsigned char>(plan.zero_bytes);
    unsigned char* state = zero;
    unsigned char* blocked = zero + tf_align_up((size_t)n, 256);
    unsigned long long* scalars = reinterpret_cast<unsigned long long*>(zero + 2 * tf_align_up((size_t)n, 256));
    unsigned char* flags = ar.take<unsigned char>(n);
    int* selected = ar.take<int>(n);
    int2* edges0 = ar.take<int2>(plan.edge_cap);
    int2* edges1 = ar.take<int2>(plan.edge_cap);
    const int nb = (n + 255) / 256;
    const int stop_after = tfg::debug_flag(14);

    TF_CHECK_CUDA(cudaMemsetAsync(zero, 0, plan.zero_bytes, st));                       // the only memset of the algorithm
    prep_keys_kernel<T><<<nb, 256, 0, st>>>((const T*)scores, n, keys_in, iota);
    size_t sb_bytes = plan.sort_bytes;
    TF_CHECK_CUDA(cub::DeviceRadixSort::SortPairsDescending(sort_tmp, sb_bytes, (const T*)keys_in, keys, (const int*)iota, order, n,
                                                            0, (int)sizeof(T) * 8, st));
    if (use_grid > 0) {
        // size-class grid: boxes by rank + min corner -> (class, cell) keys -> sort -> grid-ordered copies -> neighbourhood scan
        grid_prep_kernel<T><<<nb, 256, 0, st>>>((const T*)boxes, order, n, sb, area, scalars);
        grid_key_kernel<T><<<nb, 256, 0, st>>>(sb, n, scalars, gkey_in, iota);
        sb_bytes = plan.sort_bytes;
        TF_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(sort_tmp, sb_bytes, (const unsigned long long*)gkey_in, gkey, (const int*)iota, xorder, n, 0, 38, st));
        grid_levels_kernel<<<1, 32, 0, st>>>(gkey, n, scalars);
        grid_gather_kernel<T><<<nb, 256, 0, st>>>(sb, area, xorder, n, xb, xarea, reinterpret_cast<GridRec*>(rec));
        if (stop_after == 1) { TF_LAUNCH_CHECK(); return TF_OK; }
        const int sweep_blocks = (int)(((long long)n * 32 + 255) / 256);
        grid_sweep_kernel<T><<<sweep_blocks, 256, 0, st>>>(xb, xarea, reinterpret_cast<const GridRec*>(rec), xorder, gkey, n, thr, scalars, edges0,
                                                           (unsigned long long)plan.edge_cap, tfg::debug_flag(13));
    } else {
    gather_rank_kernel<T><<<nb, 256, 0, st>>>((const T*)boxes, order, n, sb, area, xkey);
    sb_bytes = plan.sort_bytes;
    TF_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(sort_tmp, sb_bytes, (const float*)xkey, xsorted, (const int*)iota, xorder, n, 0, 32, st));
    gather_x_kernel<T><<<nb, 256, 0, st>>>(sb, area, xorder, xsorted, n, xb, xarea, rec);
    if (stop_after == 1) { TF_LAUNCH_CHECK(); return TF_OK; }
    const int sweep_blocks = (int)(((long long)n * 32 + 255) / 256);
    sweep_kernel<T><<<sweep_blocks, 256, 0, st>>>(xb, xarea, rec, n, thr, scalars, edges0, (unsigned long long)plan.edge_cap, tfg::debug_flag(13));
    }
    if (stop_after == 2) { TF_LAUNCH_CHECK(); return TF_OK; }
    {
        const int grid = resolve_grid();
        TF_REQUIRE(grid > 0, "tf_nms(sweep): cannot size the cooperative resolve kernel");
        unsigned long long cap = (unsigned long long)plan.edge_cap;
        volatile unsigned char* vstate = state;
        volatile unsigned char* vblocked = blocked;
        volatile unsigned long long* vsc = scalars;
        int nn = n;
        void* args[] = {&edges0, &edges1, &cap, &vstate, &vblocked, &nn, &vsc};
        TF_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)resolve_kernel, dim3(grid), dim3(256), args, 0, st));
    }
    if (stop_after == 3) { TF_LAUNCH_CHECK(); return TF_OK; }
    flags_kernel<<<nb, 256, 0, st>>>(state, n, flags);
    size_t sel_bytes = plan.select_bytes;
    TF_CHECK_CUDA(cub::DeviceSelect::Flagged(select_tmp, sel_bytes, (const int*)order, flags, selected,
                                             reinterpret_cast<int*>(scalars + SC_SEL), n, st));
    finish_kernel<<<nb < 1024 ? nb : 1024, 256, 0, st>>>(selected, scalars, (unsigned long long)plan.edge_cap, keep, num_keep);
    TF_LAUNCH_CHECK();
    return TF_OK;
}

// diagnostic read-back (synchronises): {edges, IoU pair tests (needs tf_debug_set(13, 1)), rounds, edge capacity}
template <typename T>
int sweep_stats(int64_t n, void* ws, size_t ws_bytes, long long* out4, cudaStream_t st) {
    SweepPlan<T> plan(n, ws_bytes);
    TfArena ar(ws, ws_bytes);
    ar.take<int>(n); ar.take<int>(n); ar.take<T>(n); ar.take<T>(n); ar.take<char>(plan.sort_bytes); ar.take<char>(plan.select_bytes);
    ar.take<Box<T>>(n); ar.take<T>(n); ar.take<float>(n); ar.take<float>(n); ar.take<int>(n); ar.take<Box<T>>(n); ar.take<T>(n);
    ar.take<SweepRec>(n); ar.take<unsigned long long>(n); ar.take<unsigned long long>(n);
    unsigned char* zero = ar.take<unsigned char>(plan.zero_bytes);
    unsigned long long h[SC_WORDS];
    TF_CHECK_CUDA(cudaMemcpyAsync(h, zero + 2 * tf_align_up((size_t)n, 256), sizeof(h), cudaMemcpyDeviceToHost, st));
    TF_CHECK_CUDA(cudaStreamSynchronize(st));
    out4[0] = (long long)h[SC_EDGES]; out4[1] = (long long)h[SC_PAIRS]; out4[2] = (long long)h[SC_ROUNDS]; out4[3] = (long long)plan.edge_cap;
    return TF_OK;
}

template int run_nms_sweep<double>(const void*, const void*, int64_t, double, long long*, long long*, void*, size_t, cudaStream_t, int);
template int run_nms_sweep<float>(const void*, const void*, int64_t, double, long long*, long long*, void*, size_t, cudaStream_t, int);
template int sweep_stats<double>(int64_t, void*, size_t, long long*, cudaStream_t);
template int sweep_stats<float>(int64_t, void*, size_t, long long*, cudaStream_t);

}  // namespace tfnms
