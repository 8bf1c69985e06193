// Shared host/device helpers for libocr_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ocr_b200.h"

namespace ocr {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define OCR_CHECK_ARG(cond, ...)            \
    do {                                    \
        if (!(cond)) {                      \
            ocr::set_error(__VA_ARGS__);    \
            return OCR_EINVAL;              \
        }                                   \
    } while (0)

#define OCR_CHECK_CUDA(expr)                                                                  \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ocr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return OCR_ECUDA;                                                                 \
        }                                                                                     \
    } while (0)

#define OCR_CHECK_LAUNCH()                                                                    \
    do {                                                                                      \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) {                                                              \
            ocr::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return OCR_ECUDA;                                                                 \
        }                                                                                     \
        ocr::count_launch();                                                                  \
    } while (0)

constexpr int kMaxDynSmem = 226 * 1024;  // dynamic smem budget per CTA (227 KB usable on sm_100, minus 1 KB for static)

#ifdef __CUDACC__
constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFullMask, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}
// streaming (read-once / write-once) global accesses: keep them out of L1
__device__ __forceinline__ float ld_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(float* p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v));
}
#endif

}  // namespace ocr
