/* TEST INFRASTRUCTURE ONLY -- CPU oracle, never shipped, never on the product path.
 *
 * Deterministic single-precision exp / log / log1p / log-sum-exp.
 *
 * Why: the reference's CTC beam-search decoder (tf.nn.ctc_beam_search_decoder, called at
 * /root/reference/src/weinman/test.py:84-88 and src/weinman/client.py:227-231) ranks beams by
 * float32 log-probabilities built from exp/log1p.  libm's and the GPU's expf/log1pf differ in
 * the last ulp, which can flip a near-tie at the beam cut.  To make "bit-exact decode" a
 * property that can actually be tested, both the oracle (this file) and the CUDA path
 * (cnn_lstm_ctc_ocr_b200/csrc/det_math.cuh, written independently from this spec) evaluate
 * the SAME sequence of IEEE-754 binary32 operations (add, mul, div, fma, rint, integer bit
 * manipulation -- each correctly rounded on both CPU and GPU), so they agree bit for bit.
 *
 * Spec (all ops round-to-nearest-even binary32, NO contraction other than the fmaf shown):
 *   det_expf(x): x < -86 -> 0; n = rintf(x*LOG2E); r = fmaf(n,-LN2_HI,x); r = fmaf(n,-LN2_LO,r);
 *                p = Horner(EXP_C0..C5, r) with fmaf; e = fmaf(p, r*r, r) + 1; scale by 2^n through
 *                the exponent field.  (Cephes expf polynomial; <= 2 ulp on [-86, 0].)
 *   det_logf(x): x normal positive. frexp via bits -> m in [0.5,1), e; if m < SQRT_HALF
 *                {e--; m = m+m-1} else {m = m-1}; z = m*m; y = Horner(LOG_C0..C8, m) * m * z;
 *                if e: y = fmaf(LOG_LN2_LO, e, y); y = fmaf(-0.5, z, y); r = m + y;
 *                if e: r = fmaf(LOG_LN2_HI, e, r).  (Cephes logf polynomial.)
 *   det_log1pf(y): y in [0,1]: u = 1 + y; u == 1 ? y : det_logf(u) * (y / (u - 1)).
 *   det_lse2(a,b): both -inf -> -inf; hi = max, lo = min; hi + det_log1pf(det_expf(lo - hi)).
 *
 * Compile with -ffp-contract=off so the C compiler neither fuses nor reorders.
 */
#ifndef ORACLE_DET_MATH_H
#define ORACLE_DET_MATH_H
#include <math.h>
#include <stdint.h>
#include <string.h>

#define DET_LOG2E   1.44269504088896341f
#define DET_LN2_HI  0.693359375f
#define DET_LN2_LO  -2.12194440e-4f

static inline float det_bits2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t det_f2bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

static inline float det_expf(float x) {
    if (!(x >= -86.0f)) return 0.0f; /* also catches -inf and NaN */
    float n = rintf(x * DET_LOG2E);
    float r = fmaf(n, -DET_LN2_HI, x);
    r = fmaf(n, -DET_LN2_LO, r);
    float p = 1.9875691500E-4f;
    p = fmaf(p, r, 1.3981999507E-3f);
    p = fmaf(p, r, 8.3334519073E-3f);
    p = fmaf(p, r, 4.1665795894E-2f);
    p = fmaf(p, r, 1.6666665459E-1f);
    p = fmaf(p, r, 5.0000001201E-1f);
    float r2 = r * r;
    float e = fmaf(p, r2, r) + 1.0f;
    int32_t ni = (int32_t)n;
    return det_bits2f(det_f2bits(e) + ((uint32_t)ni << 23));
}

static inline float det_logf(float x) {
    uint32_t b = det_f2bits(x);
    int32_t e = (int32_t)((b >> 23) & 0xffu) - 126;             /* x = m * 2^e, m in [0.5,1) */
    float m = det_bits2f((b & 0x007fffffu) | 0x3f000000u);
    if (m < 0.707106781186547524f) { e -= 1; m = (m + m) - 1.0f; }
    else { m = m - 1.0f; }
    float z = m * m;
    float y = 7.0376836292E-2f;
    y = fmaf(y, m, -1.1514610310E-1f);
    y = fmaf(y, m, 1.1676998740E-1f);
    y = fmaf(y, m, -1.2420140846E-1f);
    y = fmaf(y, m, 1.4249322787E-1f);
    y = fmaf(y, m, -1.6668057665E-1f);
    y = fmaf(y, m, 2.0000714765E-1f);
    y = fmaf(y, m, -2.4999993993E-1f);
    y = fmaf(y, m, 3.3333331174E-1f);
    y = (y * m) * z;
    float fe = (float)e;
    if (e != 0) y = fmaf(DET_LN2_LO, fe, y);
    y = fmaf(-0.5f, z, y);
    float r = m + y;
    if (e != 0) r = fmaf(DET_LN2_HI, fe, r);
    return r;
}

static inline float det_log1pf(float y) {
    float u = 1.0f + y;
    if (u == 1.0f) return y;
    return det_logf(u) * (y / (u - 1.0f));
}

static inline float det_lse2(float a, float b) {
    if (a == -INFINITY && b == -INFINITY) return -INFINITY;
    float hi = a > b ? a : b;
    float lo = a > b ? b : a;
    return hi + det_log1pf(det_expf(lo - hi));
}
#endif
