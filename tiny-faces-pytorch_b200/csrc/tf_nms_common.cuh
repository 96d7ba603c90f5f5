// Shared pieces of the two NMS implementations: exact-arithmetic helpers and the suppression predicate.
#pragma once
#include "tf_common.cuh"

namespace tfnms {

template <typename T> struct Arith;
template <> struct Arith<double> {
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};
template <> struct Arith<float> {
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};

template <typename T> struct alignas(4 * sizeof(T)) Box { T x1, y1, x2, y2; };

// torchvision nms_kernel_impl: ovr = inter / (iarea + jarea - inter); suppress iff ovr > thr
template <typename T>
__device__ __forceinline__ bool suppresses(const Box<T>& a, T aa, const Box<T>& b, T ba, double thr,
                                           bool prefilter) {
    using A = Arith<T>;
    // std::max(a, b) = (a < b) ? b : a and std::min(a, b) = (b < a) ? b : a, spelled out so that NaN coordinates and
    // signed zeros propagate exactly as in the C++ reference
    T xx1 = a.x1 < b.x1 ? b.x1 : a.x1;       // std::max(ix1, x1[j])
    T yy1 = a.y1 < b.y1 ? b.y1 : a.y1;
    T xx2 = b.x2 < a.x2 ? b.x2 : a.x2;       // std::min(ix2, x2[j])
    T yy2 = b.y2 < a.y2 ? b.y2 : a.y2;
    T w = A::sub(xx2, xx1), h = A::sub(yy2, yy1);
    // thr >= 0: a pair without positive overlap has ovr == 0 or NaN and can never suppress
    if (prefilter && !(w > (T)0 && h > (T)0)) return false;
    w = (T)0 < w ? w : (T)0;                 // std::max(0, w)
    h = (T)0 < h ? h : (T)0;
    T inter = A::mul(w, h);
    T ovr = A::div(inter, A::sub(A::add(aa, ba), inter));
    return (double)ovr > thr;
}

// Sort keys with torch.sort's ordering (the stable descending sort torchvision's nms applies to the scores): every NaN is
// the largest value and all NaNs compare equal; -0.0 == +0.0.  A radix sort orders floats by bit pattern (-0.0 < +0.0,
// NaNs by sign and payload), so the keys are canonicalised first: -0.0 -> +0.0, any NaN -> the positive quiet NaN (which the
// radix order places above +inf).  The payload (original index) keeps ties stable.
__device__ __forceinline__ double canonical_key(double s) {
    if (s != s) return __longlong_as_double(0x7FF8000000000000ll);
    return s == 0.0 ? 0.0 : s;
}
__device__ __forceinline__ float canonical_key(float s) {
    if (s != s) return __int_as_float(0x7FC00000);
    return s == 0.f ? 0.f : s;
}
template <typename T>
__global__ void prep_keys_kernel(const T* __restrict__ scores, int n, T* __restrict__ keys, int* __restrict__ iota) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = canonical_key(scores[i]); iota[i] = i; }
}

}  // namespace tfnms
