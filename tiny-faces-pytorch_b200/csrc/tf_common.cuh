// Shared helpers for the tinyfaces_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>

#define TF_OK 0
#define TF_ERR_INVALID -1     // bad argument / unsupported shape
#define TF_ERR_CUDA -2        // a CUDA runtime/driver call failed
#define TF_ERR_WORKSPACE -3   // caller workspace too small
#define TF_ERR_CAPACITY -4    // caller output capacity too small
#define TF_ERR_INTERNAL -5

#define TF_API extern "C" __attribute__((visibility("default")))

void tf_set_error(const char* fmt, ...);

#define TF_CHECK_CUDA(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            tf_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,           \
                         cudaGetErrorString(_e));                                        \
            return TF_ERR_CUDA;                                                          \
        }                                                                                \
    } while (0)

#define TF_REQUIRE(cond, ...)                                                            \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            tf_set_error(__VA_ARGS__);                                                   \
            return TF_ERR_INVALID;                                                       \
        }                                                                                \
    } while (0)

#define TF_LAUNCH_CHECK() TF_CHECK_CUDA(cudaGetLastError())

static inline size_t tf_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// bump allocator over a caller-provided workspace
struct TfArena {
    char* base;
    size_t size, off;
    TfArena(void* p, size_t n) : base((char*)p), size(n), off(0) {}
    template <typename T>
    T* take(size_t count) {
        off = tf_align_up(off, 256);
        T* r = (T*)(base + off);
        off += count * sizeof(T);
        return r;
    }
    bool ok() const { return off <= size; }
};

#ifdef __CUDACC__
__device__ __forceinline__ float tf_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double tf_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int tf_warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// round-to-nearest (ties away) fp32 -> tf32 kept in an fp32 container
__device__ __forceinline__ float tf_round_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
#endif  // __CUDACC__
