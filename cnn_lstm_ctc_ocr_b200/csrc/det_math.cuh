// Deterministic binary32 exp / log / log1p / log-sum-exp for the beam-search decoder.
//
// The CTC beam search ranks beams by float32 log-probabilities; a last-ulp difference between two
// math libraries can flip a near-tie at the beam cut and change the decoded labels.  These
// functions are therefore written as an explicit sequence of correctly-rounded IEEE-754 binary32
// operations (add, mul, div, fma, round-to-nearest-even integer conversion, exponent-field
// arithmetic) with contraction ruled out by intrinsics, so any IEEE machine that follows the same
// recipe gets the same bits (the CPU test oracle does, independently).
//
// Recipe (Cephes single-precision polynomials):
//   exp(x), x in [-86, 0] (0 below -86):  n = rint(x*log2e); r = x - n*ln2 (two-step, fma);
//        p = Horner degree 5 in r;  e = fma(p, r*r, r) + 1;  result = e * 2^n via the exponent field.
//   log(x), x positive normal: x = m*2^e, m in [0.5,1); m < sqrt(1/2) ? (e--, m = 2m-1) : (m = m-1);
//        z = m*m; y = Horner degree 8 in m, times m*z; y += e*ln2_lo; y -= z/2; r = m + y; r += e*ln2_hi.
//   log1p(y), y in [0,1]: u = 1+y; u == 1 ? y : log(u) * (y / (u-1)).
//   lse2(a,b): both -inf -> -inf; hi + log1p(exp(lo - hi)).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace ocr {

__device__ __forceinline__ float det_expf(float x) {
    if (!(x >= -86.0f)) return 0.0f;
    const float n = rintf(__fmul_rn(x, 1.44269504088896341f));
    float r = __fmaf_rn(n, -0.693359375f, x);
    r = __fmaf_rn(n, 2.12194440e-4f, r);
    float p = 1.9875691500E-4f;
    p = __fmaf_rn(p, r, 1.3981999507E-3f);
    p = __fmaf_rn(p, r, 8.3334519073E-3f);
    p = __fmaf_rn(p, r, 4.1665795894E-2f);
    p = __fmaf_rn(p, r, 1.6666665459E-1f);
    p = __fmaf_rn(p, r, 5.0000001201E-1f);
    const float r2 = __fmul_rn(r, r);
    const float e = __fadd_rn(__fmaf_rn(p, r2, r), 1.0f);
    const int ni = (int)n;
    return __uint_as_float(__float_as_uint(e) + ((unsigned)ni << 23));
}

__device__ __forceinline__ float det_logf(float x) {
    const unsigned b = __float_as_uint(x);
    int e = (int)((b >> 23) & 0xffu) - 126;
    float m = __uint_as_float((b & 0x007fffffu) | 0x3f000000u);
    if (m < 0.707106781186547524f) { e -= 1; m = __fadd_rn(__fadd_rn(m, m), -1.0f); }
    else { m = __fadd_rn(m, -1.0f); }
    const float z = __fmul_rn(m, m);
    float y = 7.0376836292E-2f;
    y = __fmaf_rn(y, m, -1.1514610310E-1f);
    y = __fmaf_rn(y, m, 1.1676998740E-1f);
    y = __fmaf_rn(y, m, -1.2420140846E-1f);
    y = __fmaf_rn(y, m, 1.4249322787E-1f);
    y = __fmaf_rn(y, m, -1.6668057665E-1f);
    y = __fmaf_rn(y, m, 2.0000714765E-1f);
    y = __fmaf_rn(y, m, -2.4999993993E-1f);
    y = __fmaf_rn(y, m, 3.3333331174E-1f);
    y = __fmul_rn(__fmul_rn(y, m), z);
    const float fe = (float)e;
    if (e != 0) y = __fmaf_rn(-2.12194440e-4f, fe, y);
    y = __fmaf_rn(-0.5f, z, y);
    float r = __fadd_rn(m, y);
    if (e != 0) r = __fmaf_rn(0.693359375f, fe, r);
    return r;
}

__device__ __forceinline__ float det_log1pf(float y) {
    const float u = __fadd_rn(1.0f, y);
    if (u == 1.0f) return y;
    return __fmul_rn(det_logf(u), __fdiv_rn(y, __fadd_rn(u, -1.0f)));
}

__device__ __forceinline__ float det_lse2(float a, float b) {
    if (a == -CUDART_INF_F && b == -CUDART_INF_F) return -CUDART_INF_F;
    const float hi = a > b ? a : b;
    const float lo = a > b ? b : a;
    return __fadd_rn(hi, det_log1pf(det_expf(__fadd_rn(lo, -hi))));
}

}  // namespace ocr
