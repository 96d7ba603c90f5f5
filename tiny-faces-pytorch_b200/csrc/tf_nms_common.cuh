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
    T xx1 = a.x1 > b.x1 ? a.x1 : b.x1;       // std::max(ix1, x1[j])
    T yy1 = a.y1 > b.y1 ? a.y1 : b.y1;
    T xx2 = b.x2 < a.x2 ? b.x2 : a.x2;       // std::min(ix2, x2[j])
    T yy2 = b.y2 < a.y2 ? b.y2 : a.y2;
    T w = A::sub(xx2, xx1), h = A::sub(yy2, yy1);
    // thr >= 0: a pair without positive overlap has ovr == 0 or NaN and can never suppress
    if (prefilter && !(w > (T)0 && h > (T)0)) return false;
    w = w > (T)0 ? w : (T)0;                 // std::max(0, w)
    h = h > (T)0 ? h : (T)0;
    T inter = A::mul(w, h);
    T ovr = A::div(inter, A::sub(A::add(aa, ba), inter));
    return (double)ovr > thr;
}


}  // namespace tfnms
