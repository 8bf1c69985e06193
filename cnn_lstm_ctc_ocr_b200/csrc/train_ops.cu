// Training-step kernels of the recognizer (sm_100a): everything train.py's graph adds around the dense contractions.
//
//   batch-norm with BATCH statistics + moving-average update     model.py:118-123 (training=True), train.py:116-118
//   gradients of ReLU / bias-add / max-pool / pool8+squeeze        model.py:97-116,145-147 (TensorFlow's registered gradients)
//   LSTMCell frames that keep what back-propagation through time needs, and the BPTT frame loop
//                                                                   model_bu.py:167-199 (bidirectional_dynamic_rnn)
//   weight-gradient contractions on tcgen05 (split-K over the long pixel / frame dimension, operands transposed so
//   that the contraction index is contiguous; the 9 filter taps are 9 shifted views of ONE padded planar copy)
//   Adam (tf.train.AdamOptimizer) over one flat parameter buffer     train.py:128-137
//
// The memory-bound kernels are coalesced grid-stride loops; per-channel reductions accumulate float partials per thread,
// combine them per CTA in shared memory and finish in double-precision atomics (order-independent to ~1e-16).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <type_traits>

#include "gemm_tf32.cuh"

namespace ocr {

static inline int grid_cap(long long total, int threads = 256, int waves = 16) {
    long long g = (total + threads - 1) / threads;
    const long long cap = 148LL * waves;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + __expf(-x)); }

// ---------------------------------------------------------------------------------------------------------------
// transposes: out[c][r] = src[r + src_shift][c] (zero when r + src_shift is outside [0, rows)), src [rows, cols]
// row-major, out row pitch ld.  PAD: `rows` counts the pixels of a zero-ringed grid [B, H+2, Wp] (Wp = W+2 rounded up to
// a multiple of 4, so that whole-row shifts of the planar copy stay 16-byte aligned for TMA); ring pixels read as zero,
// interior pixels come from the dense NHWC tensor.  blockIdx.z selects one of the pre-shifted copies (src_shift + z).
template <bool PAD>
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ in, long long rows, int cols, int ld_in, float* __restrict__ out, long long ld, int H, int W, int Wp,
                 long long src_shift, long long copy_stride)
{
    __shared__ float tile[32][33];
    const long long r0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    src_shift += blockIdx.z;
    out += (size_t)blockIdx.z * copy_stride;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long r = r0 + ty + i * 8 + src_shift;
        const int c = c0 + tx;
        float v = 0.0f;
        if (r >= 0 && r < rows && c < cols) {
            if (PAD) {       // rows < 2^31 (checked by the caller): 32-bit divisions
                const unsigned Hp = H + 2, ru = (unsigned)r;
                const unsigned q = ru / (unsigned)Wp, xp = ru - q * Wp;
                const unsigned b = q / Hp, yp = q - b * Hp;
                if (xp >= 1 && xp <= (unsigned)W && yp >= 1 && yp <= (unsigned)H) v = __ldg(in + (((size_t)b * H + (yp - 1)) * W + (xp - 1)) * ld_in + c);
            } else {
                v = __ldg(in + (size_t)r * ld_in + c);
            }
        }
        tile[ty + i * 8][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + i * 8;
        const long long r = r0 + tx;
        if (c < cols && r < rows) out[(size_t)c * ld + r] = tile[tx][ty + i * 8];
    }
}

// Zero-ringed channel-planar copies, the fast path of ocr_nhwc_to_planar_pad: one CTA = 128 padded pixels x 32 channels.
// The tile (with a one-pixel halo on both sides) is read ONCE with float4 loads along the channels and written NCOPY
// times (pixel shifts -1, 0, +1, or just 0) as 512-byte runs along the pixels.
// BLOCKED: the K-blocked layout of gemm_wgrad(blocked) -- element (channel row, r) at ((r / 32) * ld + row) * 32 + r % 32, where ld
// is then the operand's total row count and copy k starts copy_stride ROWS further down.
template <int NCOPY, bool BLOCKED = false>
__global__ void __launch_bounds__(256)
planar_pad_kernel(const float* __restrict__ in, unsigned rows, int C, float* __restrict__ out, long long ld, int H, int W, int Wp,
                  long long copy_stride)
{
    constexpr int TP = 128, PITCH = 133;   // odd pitch: the channel-major writes below hit distinct banks
    __shared__ float tile[32][PITCH];
    const unsigned r0 = blockIdx.x * TP;
    const int c0 = blockIdx.y * 32;
    const int q = threadIdx.x & 7;          // channel quad
    const int pl = threadIdx.x >> 3;        // pixel lane 0..31
    const unsigned Hp = H + 2;
    // tile column j holds padded pixel r0 - 1 + j, j = 0 .. TP + 1
    for (int j = pl; j < TP + 2; j += 32) {
        const long long r = (long long)r0 - 1 + j;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int c = c0 + 4 * q;
        if (r >= 0 && r < (long long)rows && c < C) {
            const unsigned ru = (unsigned)r;
            const unsigned qq = ru / (unsigned)Wp, xp = ru - qq * Wp;
            const unsigned b = qq / Hp, yp = qq - b * Hp;
            if (xp >= 1 && xp <= (unsigned)W && yp >= 1 && yp <= (unsigned)H)
                v = __ldg(reinterpret_cast<const float4*>(in + (((size_t)b * H + (yp - 1)) * W + (xp - 1)) * C + c));
        }
        tile[4 * q + 0][j] = v.x; tile[4 * q + 1][j] = v.y; tile[4 * q + 2][j] = v.z; tile[4 * q + 3][j] = v.w;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NCOPY; ++k) {
        const int sh = (NCOPY == 3) ? k : 1;     // copy k holds pixel r + k - 1 at r  -> tile column (r - r0) + k
        float* o = BLOCKED ? out : out + (size_t)k * copy_stride;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = warp + 8 * i;
            if (c0 + c >= C) continue;
#pragma unroll
            for (int j = 0; j < TP / 32; ++j) {
                const unsigned r = r0 + j * 32 + lane;
                if (BLOCKED) {
                    if (r < rows) o[((size_t)(r >> 5) * ld + (size_t)k * copy_stride + c0 + c) * 32 + lane] = tile[c][j * 32 + lane + sh];
                } else {
                    if (r < rows) o[(size_t)(c0 + c) * ld + r] = tile[c][j * 32 + lane + sh];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// per-channel reductions over the rows of x [rows, C].  One CTA = 8 warps x 32 channels (blockIdx.y = channel group);
// warp w walks rows w, w + 8*gridDim.x, ...; lane = channel.  MODE selects what is summed:
//   0: s0 = sum x, s1 = sum x^2                                        (batch-norm statistics)
//   1: s0 = sum x                                                      (bias gradient)
//   2: dz = g * (z > 0), z = gamma*(x-mean)*inv_std+beta: s0 = sum dz, s1 = sum dz * xhat   (batch-norm backward)
//   3: dy = g * (x > 0) written to out; s0 = sum dy                    (ReLU backward + bias gradient; x = layer output)
template <int MODE>
__global__ void __launch_bounds__(256)
channel_reduce_kernel(const float* __restrict__ x, const float* __restrict__ g, long long rows, int C, int ldx, const float* __restrict__ mean,
                      const float* __restrict__ inv_std, const float* __restrict__ gamma, const float* __restrict__ beta,
                      float* __restrict__ out, double* __restrict__ sums /*[2][C]*/)
{
    __shared__ float red[2][8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.y * 32 + lane;
    const bool ok = c < C;
    float mu = 0.f, is = 0.f, ga = 0.f, be = 0.f;
    if (MODE == 2 && ok) { mu = mean[c]; is = inv_std[c]; ga = gamma[c]; be = beta[c]; }
    float s0 = 0.f, s1 = 0.f, k0 = 0.f, k1 = 0.f;   // Kahan-compensated partial sums
    const long long stride = (long long)gridDim.x * 8;
    for (long long rb = (long long)blockIdx.x * 8 + warp; rb < rows; rb += 4 * stride) {
        if (!ok) continue;
        float xv[4], gv[4];
        bool live[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {     // four independent rows in flight
            const long long r = rb + u * stride;
            live[u] = r < rows;
            const size_t o = (size_t)(live[u] ? r : rb) * ldx + c;
            xv[u] = x[o];
            gv[u] = (MODE == 2 || MODE == 3) ? g[o] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (!live[u]) continue;
            float a0 = 0.f, a1 = 0.f;
            if (MODE == 0) { a0 = xv[u]; a1 = xv[u] * xv[u]; }
            if (MODE == 1) { a0 = xv[u]; }
            if (MODE == 2) {
                const float xh = (xv[u] - mu) * is;
                const float dz = (ga * xh + be > 0.f) ? gv[u] : 0.f;
                a0 = dz; a1 = dz * xh;
            }
            if (MODE == 3) {
                const float dy = xv[u] > 0.f ? gv[u] : 0.f;
                out[(size_t)(rb + u * stride) * ldx + c] = dy;
                a0 = dy;
            }
            { const float y = a0 - k0; const float t = s0 + y; k0 = (t - s0) - y; s0 = t; }
            if (MODE == 0 || MODE == 2) { const float y = a1 - k1; const float t = s1 + y; k1 = (t - s1) - y; s1 = t; }
        }
    }
    red[0][warp][lane] = s0;
    red[1][warp][lane] = s1;
    __syncthreads();
    if (warp == 0 && ok) {
        double t0 = 0.0, t1 = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { t0 += (double)red[0][w][lane]; t1 += (double)red[1][w][lane]; }
        atomicAdd(sums + c, t0);
        if (MODE == 0 || MODE == 2) atomicAdd(sums + C + c, t1);
    }
}

// batch statistics -> mean, inv_std; moving averages (tf.layers.batch_normalization, momentum 0.99; the fused kernel
// feeds the UNBIASED batch variance to the moving variance).
__global__ void bn_finalize_kernel(const double* __restrict__ sums, long long n, int C, float eps, float momentum, float* __restrict__ mean,
                                   float* __restrict__ inv_std, float* __restrict__ moving_mean, float* __restrict__ moving_var)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double m = sums[c] / (double)n;
    double var = sums[C + c] / (double)n - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    inv_std[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (moving_mean != nullptr) {
        const double unb = n > 1 ? var * ((double)n / (double)(n - 1)) : var;
        moving_mean[c] = (float)((double)momentum * moving_mean[c] + (1.0 - (double)momentum) * m);
        moving_var[c] = (float)((double)momentum * moving_var[c] + (1.0 - (double)momentum) * unb);
    }
}
__global__ void sums_to_float_kernel(const double* __restrict__ sums, int n, float* __restrict__ o0, float* __restrict__ o1)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    if (o0) o0[c] = (float)sums[c];
    if (o1) o1[c] = (float)sums[n + c];
}

// out = relu(gamma * (y - mean) * inv_std + beta), float4 over channels
__global__ void __launch_bounds__(256)
bn_relu_apply_kernel(const float* __restrict__ y, long long rows, int C, const float* __restrict__ mean, const float* __restrict__ inv_std,
                     const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ out)
{
    const int c4n = C >> 2;
    const long long total = rows * c4n;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % c4n) * 4;
        const float4 v = reinterpret_cast<const float4*>(y)[idx];
        const float4 mu = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(inv_std + c);
        const float4 ga = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
        float4 o;
        o.x = fmaxf(ga.x * ((v.x - mu.x) * is.x) + be.x, 0.f);
        o.y = fmaxf(ga.y * ((v.y - mu.y) * is.y) + be.y, 0.f);
        o.z = fmaxf(ga.z * ((v.z - mu.z) * is.z) + be.z, 0.f);
        o.w = fmaxf(ga.w * ((v.w - mu.w) * is.w) + be.w, 0.f);
        reinterpret_cast<float4*>(out)[idx] = o;
    }
}

// z = relu(batch-norm(y)) written to `out` AND the max-pool that follows the layer (window 2x2, stride (2, sw), 'valid';
// model.py:105-116) written to `pooled`, in one pass over y: the thread of a window's top-left pixel normalises the window's
// other three pixels too (their y values are its neighbours' own loads: cache hits) and stores the maximum.
__global__ void __launch_bounds__(256)
bn_relu_apply_pool_kernel(const float* __restrict__ y, int B, int H, int W, int C, const float* __restrict__ mean, const float* __restrict__ inv_std,
                          const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ out, int sw, int Hp, int Wp,
                          float* __restrict__ pooled)
{
    const int c4n = C >> 2;
    const long long total = (long long)B * H * W * c4n;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % c4n), c = c4 * 4;
        long long p = idx / c4n;
        const int x = (int)(p % W); p /= W;
        const int yy = (int)(p % H);
        const int b = (int)(p / H);
        const float4 mu = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(inv_std + c);
        const float4 ga = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
        auto act = [&](const float4 v) {
            float4 o;
            o.x = fmaxf(ga.x * ((v.x - mu.x) * is.x) + be.x, 0.f);
            o.y = fmaxf(ga.y * ((v.y - mu.y) * is.y) + be.y, 0.f);
            o.z = fmaxf(ga.z * ((v.z - mu.z) * is.z) + be.z, 0.f);
            o.w = fmaxf(ga.w * ((v.w - mu.w) * is.w) + be.w, 0.f);
            return o;
        };
        const float4 z0 = act(reinterpret_cast<const float4*>(y)[idx]);
        reinterpret_cast<float4*>(out)[idx] = z0;
        const int py = yy >> 1, px = sw == 2 ? x >> 1 : x;
        if ((yy & 1) == 0 && py < Hp && px < Wp && (sw == 1 || (x & 1) == 0)) {     // top-left pixel of a window
            const float4 z1 = act(__ldg(reinterpret_cast<const float4*>(y) + idx + c4n));
            const float4 z2 = act(__ldg(reinterpret_cast<const float4*>(y) + idx + (long long)W * c4n));
            const float4 z3 = act(__ldg(reinterpret_cast<const float4*>(y) + idx + (long long)W * c4n + c4n));
            float4 m;
            m.x = fmaxf(fmaxf(z0.x, z1.x), fmaxf(z2.x, z3.x)); m.y = fmaxf(fmaxf(z0.y, z1.y), fmaxf(z2.y, z3.y));
            m.z = fmaxf(fmaxf(z0.z, z1.z), fmaxf(z2.z, z3.z)); m.w = fmaxf(fmaxf(z0.w, z1.w), fmaxf(z2.w, z3.w));
            reinterpret_cast<float4*>(pooled)[(((size_t)b * Hp + py) * Wp + px) * c4n + c4] = m;
        }
    }
}

// The same pass WITHOUT the full-size activation: pooled [B,Hp,Wp,C] and, per pooled element, which of its window's four
// pixels is the FIRST maximum in row-major window order (TensorFlow's MaxPoolGrad rule; 2 bits per channel, the four channels of
// a float4 in one byte: arg [B,Hp,Wp,C/4]).  With the argument of the maximum kept, the pool's gradient needs neither the
// activation nor a tensor of its own (pool_arg_grad below): the layer's backward pass reads y and writes dy, nothing else of
// that size.  One thread per pooled float4; the y values of a window are read once (stride 2) or twice (stride 1, cache hits).
__global__ void __launch_bounds__(256)
bn_relu_pool_arg_kernel(const float* __restrict__ y, int B, int H, int W, int C, const float* __restrict__ mean, const float* __restrict__ inv_std,
                        const float* __restrict__ gamma, const float* __restrict__ beta, int sw, int Hp, int Wp,
                        float* __restrict__ pooled, unsigned char* __restrict__ arg)
{
    const int c4n = C >> 2;
    const long long total = (long long)B * Hp * Wp * c4n;
    const float4* y4 = reinterpret_cast<const float4*>(y);
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % c4n), c = c4 * 4;
        long long p = idx / c4n;
        const int px = (int)(p % Wp); p /= Wp;
        const int py = (int)(p % Hp);
        const int b = (int)(p / Hp);
        const float4 mu = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(inv_std + c);
        const float4 ga = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
        auto act = [&](const float4 v) {
            float4 o;
            o.x = fmaxf(ga.x * ((v.x - mu.x) * is.x) + be.x, 0.f);
            o.y = fmaxf(ga.y * ((v.y - mu.y) * is.y) + be.y, 0.f);
            o.z = fmaxf(ga.z * ((v.z - mu.z) * is.z) + be.z, 0.f);
            o.w = fmaxf(ga.w * ((v.w - mu.w) * is.w) + be.w, 0.f);
            return o;
        };
        const long long i00 = (((long long)b * H + 2 * py) * W + (long long)px * sw) * c4n + c4;
        const float4 z0 = act(__ldg(y4 + i00)), z1 = act(__ldg(y4 + i00 + c4n));
        const float4 z2 = act(__ldg(y4 + i00 + (long long)W * c4n)), z3 = act(__ldg(y4 + i00 + (long long)W * c4n + c4n));
        float4 m = z0;
        unsigned a = 0;
#define OCR_PA(f, sh_)                                              \
        {                                                          \
            unsigned k = 0;                                        \
            if (z1.f > m.f) { m.f = z1.f; k = 1; }                 \
            if (z2.f > m.f) { m.f = z2.f; k = 2; }                 \
            if (z3.f > m.f) { m.f = z3.f; k = 3; }                 \
            a |= k << sh_;                                         \
        }
        OCR_PA(x, 0) OCR_PA(y, 2) OCR_PA(z, 4) OCR_PA(w, 6)
#undef OCR_PA
        reinterpret_cast<float4*>(pooled)[idx] = m;
        arg[idx] = (unsigned char)a;
    }
}

// gradient of the 2x2 / stride (2, sw) max-pool w.r.t. its input pixel (b, yy, x), channels 4*c4 .. 4*c4+3, from the pooled
// gradient and the arguments of the maxima: the sum over the windows that contain the pixel (one, or two when sw == 1, in
// ascending window order as maxpool_bwd_kernel adds them) of dpool where the window's first maximum is this pixel.
__device__ __forceinline__ float4 pool_arg_grad(const float4* __restrict__ dpool4, const unsigned char* __restrict__ arg, int b, int yy, int x,
                                                int c4, int c4n, int sw, int Hp, int Wp)
{
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int oy = yy >> 1;
    if (oy >= Hp) return acc;
    const unsigned dy = (unsigned)(yy & 1) * 2;
    auto add = [&](int ox, unsigned code) {
        const long long o = (((long long)b * Hp + oy) * Wp + ox) * c4n + c4;
        const unsigned a = arg[o];
        const float4 d = __ldg(dpool4 + o);
        if ((a & 3u) == code) acc.x += d.x;
        if (((a >> 2) & 3u) == code) acc.y += d.y;
        if (((a >> 4) & 3u) == code) acc.z += d.z;
        if (((a >> 6) & 3u) == code) acc.w += d.w;
    };
    if (sw == 2) {
        const int ox = x >> 1;
        if (ox < Wp) add(ox, dy + (unsigned)(x & 1));
    } else {
        if (x >= 1 && x - 1 < Wp) add(x - 1, dy + 1u);
        if (x < Wp) add(x, dy);
    }
    return acc;
}

// ocr_bn_relu_bwd_sums for a layer whose output went through the pool: the gradient w.r.t. the activation is pool_arg_grad,
// never materialised.  float4 form (a thread keeps its four channels: 256 % (C/4) == 0), Kahan-compensated float partials per
// thread, one shared-memory pass per CTA, double atomics per channel.
__global__ void __launch_bounds__(256)
bn_bwd_sums_pool_kernel(const float* __restrict__ y, const float* __restrict__ dpool, const unsigned char* __restrict__ arg, int B, int H, int W,
                        int C, int sw, int Hp, int Wp, const float* __restrict__ mean, const float* __restrict__ inv_std,
                        const float* __restrict__ gamma, const float* __restrict__ beta, double* __restrict__ sums /*[2][C]*/)
{
    const int c4n = C >> 2;
    const long long total = (long long)B * H * W * c4n;
    const long long first = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    const int c4 = (int)(first % c4n), c = c4 * 4;
    const float4 mu = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(inv_std + c);
    const float4 ga = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
    const float4* y4 = reinterpret_cast<const float4*>(y);
    const float4* dpool4 = reinterpret_cast<const float4*>(dpool);
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f}, k0[4] = {0.f, 0.f, 0.f, 0.f}, k1[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long idx = first; idx < total; idx += stride) {
        unsigned p = (unsigned)(idx / c4n);                 // pixel index < 2^31 (checked by the caller)
        const int x = (int)(p % (unsigned)W); p /= (unsigned)W;
        const int yy = (int)(p % (unsigned)H);
        const int b = (int)(p / (unsigned)H);
        const float4 yv = y4[idx];
        const float4 gv = pool_arg_grad(dpool4, arg, b, yy, x, c4, c4n, sw, Hp, Wp);
#define OCR_BS(f, j)                                                                   \
        {                                                                              \
            const float xh = (yv.f - mu.f) * is.f;                                     \
            const float dz = (ga.f * xh + be.f > 0.f) ? gv.f : 0.f;                    \
            { const float v = dz - k0[j]; const float t = s0[j] + v; k0[j] = (t - s0[j]) - v; s0[j] = t; }       \
            { const float v = dz * xh - k1[j]; const float t = s1[j] + v; k1[j] = (t - s1[j]) - v; s1[j] = t; }  \
        }
        OCR_BS(x, 0) OCR_BS(y, 1) OCR_BS(z, 2) OCR_BS(w, 3)
#undef OCR_BS
    }
    __shared__ float4 red0[256], red1[256];
    red0[threadIdx.x] = make_float4(s0[0], s0[1], s0[2], s0[3]);
    red1[threadIdx.x] = make_float4(s1[0], s1[1], s1[2], s1[3]);
    __syncthreads();
    if ((int)threadIdx.x < c4n) {
        double t0[4] = {0.0, 0.0, 0.0, 0.0}, t1[4] = {0.0, 0.0, 0.0, 0.0};
        for (int k = threadIdx.x; k < 256; k += c4n) {
            t0[0] += (double)red0[k].x; t0[1] += (double)red0[k].y; t0[2] += (double)red0[k].z; t0[3] += (double)red0[k].w;
            t1[0] += (double)red1[k].x; t1[1] += (double)red1[k].y; t1[2] += (double)red1[k].z; t1[3] += (double)red1[k].w;
        }
        const int cc = (int)threadIdx.x * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) { atomicAdd(sums + cc + j, t0[j]); atomicAdd(sums + C + cc + j, t1[j]); }
    }
}

// dy = gamma * inv_std * (dz - dbeta/n - xhat * dgamma/n), dz = g * (z > 0)
// BIAS: also the per-channel sums of dy (the bias gradient of the convolution in front of the batch-norm: one pass less over
// dy).  When C/4 divides 256 the grid stride is a multiple of C/4, so a thread keeps ITS four channels for the whole loop:
// the per-channel constants (eight of them from float64 sums) are set up once per thread instead of once per element, two
// elements are in flight per thread, the float partials stay in registers, one shared-memory pass per CTA, double atomics per
// channel (as channel_reduce_kernel does).  The atomics are C per CTA whatever the tensor size: the BIAS launch uses four waves
// of CTAs, not sixteen (at per-GPU batch 32 the sixteen-wave grid spent 60-100 us per layer on 1.2 M atomics to 512 addresses).
struct BnBwdConst { float4 mu, is, ga, be, db, dg; };
__device__ __forceinline__ BnBwdConst bn_bwd_const(int c, int C, const float* __restrict__ mean, const float* __restrict__ inv_std,
                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                   const double* __restrict__ sums, double inv_n)
{
    BnBwdConst k;
    k.mu = *reinterpret_cast<const float4*>(mean + c); k.is = *reinterpret_cast<const float4*>(inv_std + c);
    k.ga = *reinterpret_cast<const float4*>(gamma + c); k.be = *reinterpret_cast<const float4*>(beta + c);
    k.db = make_float4((float)(sums[c] * inv_n), (float)(sums[c + 1] * inv_n), (float)(sums[c + 2] * inv_n), (float)(sums[c + 3] * inv_n));
    k.dg = make_float4((float)(sums[C + c] * inv_n), (float)(sums[C + c + 1] * inv_n), (float)(sums[C + c + 2] * inv_n), (float)(sums[C + c + 3] * inv_n));
    return k;
}
__device__ __forceinline__ float4 bn_bwd_one(const float4 yv, const float4 gv, const BnBwdConst& k)
{
    float4 o;
#define OCR_BNB(f)                                                       \
    {                                                                    \
        const float xh = (yv.f - k.mu.f) * k.is.f;                       \
        const float dz = (k.ga.f * xh + k.be.f > 0.f) ? gv.f : 0.f;      \
        o.f = k.ga.f * k.is.f * (dz - k.db.f - xh * k.dg.f);             \
    }
    OCR_BNB(x) OCR_BNB(y) OCR_BNB(z) OCR_BNB(w)
#undef OCR_BNB
    return o;
}
template <bool BIAS>
__global__ void __launch_bounds__(256)
bn_relu_bwd_apply_kernel(const float* __restrict__ y, const float* __restrict__ g, long long rows, long long n, int C, const float* __restrict__ mean,
                         const float* __restrict__ inv_std, const float* __restrict__ gamma, const float* __restrict__ beta,
                         const double* __restrict__ sums, float* __restrict__ dy, double* __restrict__ dy_sums)
{
    const int c4n = C >> 2;
    const long long total = rows * c4n;
    const double inv_n = 1.0 / (double)n;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long first = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    const float4* y4 = reinterpret_cast<const float4*>(y);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    if ((256 % c4n) == 0) {        // BIAS launches always (checked by the caller)
        const BnBwdConst k = bn_bwd_const((int)(first % c4n) * 4, C, mean, inv_std, gamma, beta, sums, inv_n);
        long long idx = first;
        for (; idx + stride < total; idx += 2 * stride) {
            const float4 y0 = y4[idx], g0 = g4[idx], y1 = y4[idx + stride], g1 = g4[idx + stride];
            const float4 o0 = bn_bwd_one(y0, g0, k), o1 = bn_bwd_one(y1, g1, k);
            reinterpret_cast<float4*>(dy)[idx] = o0;
            reinterpret_cast<float4*>(dy)[idx + stride] = o1;
            if (BIAS) { acc.x += o0.x; acc.y += o0.y; acc.z += o0.z; acc.w += o0.w; acc.x += o1.x; acc.y += o1.y; acc.z += o1.z; acc.w += o1.w; }
        }
        if (idx < total) {
            const float4 o = bn_bwd_one(y4[idx], g4[idx], k);
            reinterpret_cast<float4*>(dy)[idx] = o;
            if (BIAS) { acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w; }
        }
    } else {
        for (long long idx = first; idx < total; idx += stride) {
            const BnBwdConst k = bn_bwd_const((int)(idx % c4n) * 4, C, mean, inv_std, gamma, beta, sums, inv_n);
            reinterpret_cast<float4*>(dy)[idx] = bn_bwd_one(y4[idx], g4[idx], k);
        }
    }
    if (BIAS) {
        __shared__ float4 red[256];
        red[threadIdx.x] = acc;
        __syncthreads();
        if ((int)threadIdx.x < c4n) {          // threads t, t + c4n, t + 2 c4n, ... hold the same four channels (256 % c4n == 0)
            double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
            for (int k = threadIdx.x; k < 256; k += c4n) { t0 += (double)red[k].x; t1 += (double)red[k].y; t2 += (double)red[k].z; t3 += (double)red[k].w; }
            const int c = (int)threadIdx.x * 4;
            atomicAdd(dy_sums + c, t0); atomicAdd(dy_sums + c + 1, t1); atomicAdd(dy_sums + c + 2, t2); atomicAdd(dy_sums + c + 3, t3);
        }
    }
}

// bn_relu_bwd_apply_kernel<true> with the gradient w.r.t. the activation taken from the pool (pool_arg_grad)
__global__ void __launch_bounds__(256)
bn_relu_bwd_apply_pool_kernel(const float* __restrict__ y, const float* __restrict__ dpool, const unsigned char* __restrict__ arg, int B, int H,
                              int W, int C, int sw, int Hp, int Wp, long long n, const float* __restrict__ mean,
                              const float* __restrict__ inv_std, const float* __restrict__ gamma, const float* __restrict__ beta,
                              const double* __restrict__ sums, float* __restrict__ dy, double* __restrict__ dy_sums)
{
    const int c4n = C >> 2;
    const long long total = (long long)B * H * W * c4n;
    const double inv_n = 1.0 / (double)n;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long first = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    const float4* y4 = reinterpret_cast<const float4*>(y);
    const float4* dpool4 = reinterpret_cast<const float4*>(dpool);
    const int c4 = (int)(first % c4n);
    const BnBwdConst k = bn_bwd_const(c4 * 4, C, mean, inv_std, gamma, beta, sums, inv_n);
    for (long long idx = first; idx < total; idx += stride) {
        unsigned p = (unsigned)(idx / c4n);
        const int x = (int)(p % (unsigned)W); p /= (unsigned)W;
        const int yy = (int)(p % (unsigned)H);
        const int b = (int)(p / (unsigned)H);
        const float4 o = bn_bwd_one(y4[idx], pool_arg_grad(dpool4, arg, b, yy, x, c4, c4n, sw, Hp, Wp), k);
        reinterpret_cast<float4*>(dy)[idx] = o;
        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    __shared__ float4 red[256];
    red[threadIdx.x] = acc;
    __syncthreads();
    if ((int)threadIdx.x < c4n) {
        double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
        for (int kk = threadIdx.x; kk < 256; kk += c4n) { t0 += (double)red[kk].x; t1 += (double)red[kk].y; t2 += (double)red[kk].z; t3 += (double)red[kk].w; }
        const int c = (int)threadIdx.x * 4;
        atomicAdd(dy_sums + c, t0); atomicAdd(dy_sums + c + 1, t1); atomicAdd(dy_sums + c + 2, t2); atomicAdd(dy_sums + c + 3, t3);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// max-pool gradient (gather form, no atomics): the gradient of a window goes to its FIRST maximum in row-major window
// order (TensorFlow's MaxPoolGrad / the arg-max of the forward pass).
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const float* __restrict__ in, const float* __restrict__ dout, int B, int H, int W, int C, int ph, int pw, int sh, int sw,
                   int Hp, int Wp, float* __restrict__ din)
{
    const int c4n = C >> 2;
    const long long total = (long long)B * H * W * c4n;
    const float4* in4 = reinterpret_cast<const float4*>(in);
    const float4* dout4 = reinterpret_cast<const float4*>(dout);
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const unsigned c4 = (unsigned)(idx % c4n);
        unsigned p = (unsigned)(idx / c4n);                 // pixel index < 2^31 (checked by the caller)
        const int x = (int)(p % (unsigned)W); p /= (unsigned)W;
        const int y = (int)(p % (unsigned)H);
        const int b = (int)(p / (unsigned)H);
        const float4 v = in4[idx];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        // windows (oy, ox) that contain (y, x)
        const int oy_lo = max(0, (y - ph + sh) / sh), oy_hi = min(Hp - 1, y / sh);
        const int ox_lo = max(0, (x - pw + sw) / sw), ox_hi = min(Wp - 1, x / sw);
        for (int oy = oy_lo; oy <= oy_hi; ++oy)
            for (int ox = ox_lo; ox <= ox_hi; ++ox) {
                if (oy * sh + ph <= y || ox * sw + pw <= x) continue;
                // per channel: is (y, x) the first maximum of this window?
                bool fx = true, fy = true, fz = true, fw = true;
                for (int dy = 0; dy < ph; ++dy)
                    for (int dx = 0; dx < pw; ++dx) {
                        const int yy = oy * sh + dy, xx = ox * sw + dx;
                        if (yy == y && xx == x) continue;
                        const float4 u = __ldg(in4 + (((size_t)b * H + yy) * W + xx) * c4n + c4);
                        const bool before = (yy < y) || (yy == y && xx < x);
                        fx = fx && !(u.x > v.x || (before && u.x == v.x));
                        fy = fy && !(u.y > v.y || (before && u.y == v.y));
                        fz = fz && !(u.z > v.z || (before && u.z == v.z));
                        fw = fw && !(u.w > v.w || (before && u.w == v.w));
                    }
                const float4 g = __ldg(dout4 + (((size_t)b * Hp + oy) * Wp + ox) * c4n + c4);
                if (fx) acc.x += g.x;
                if (fy) acc.y += g.y;
                if (fz) acc.z += g.z;
                if (fw) acc.w += g.w;
            }
        reinterpret_cast<float4*>(din)[idx] = acc;
    }
}

// pool8 + squeeze + time-major transpose backward: in [B,H,W,C], dseq [W,B,C] -> din (first maximum over the H rows)
__global__ void __launch_bounds__(256)
rows_max_bwd_kernel(const float* __restrict__ in, const float* __restrict__ dseq, int B, int H, int W, int C, float* __restrict__ din)
{
    const long long total = (long long)B * W * C;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C);
        long long p = idx / C;
        const int x = (int)(p % W);
        const int b = (int)(p / W);
        int best = 0;
        float bv = in[(((size_t)b * H) * W + x) * C + c];
        for (int y = 1; y < H; ++y) {
            const float u = in[(((size_t)b * H + y) * W + x) * C + c];
            if (u > bv) { bv = u; best = y; }
        }
        const float gd = dseq[((size_t)x * B + b) * C + c];
        for (int y = 0; y < H; ++y) din[(((size_t)b * H + y) * W + x) * C + c] = (y == best) ? gd : 0.f;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// LSTM frames for training.  Same state-row convention as model_ops.cu (rows [0,B) forward, [B,2B) backward; frame of
// step s: forward t = s, backward t = len-1-s).  The pre-activations in xp [T*B, 8H] are REPLACED by the gate
// activations (i, tanh j, f, o) and the cell state of every visited frame is kept in cs [T,B,2H].
__global__ void __launch_bounds__(256)
lstm_cell_train_kernel(const float* __restrict__ gh, float* __restrict__ xp, const int32_t* __restrict__ seq_len, int s, int T, int B, int H,
                       float* __restrict__ h, float* __restrict__ c, float* __restrict__ out /*[T,B,2H]*/, float* __restrict__ cs /*[T,B,2H]*/)
{
    const int total = 2 * B * H;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int j = idx % H;
        const int r = idx / H;
        const int dir = r >= B, b = dir ? r - B : r;
        const int len = min(seq_len[b], T);
        if (s >= len) continue;
        const int t = dir ? len - 1 - s : s;
        const float* g = gh + (size_t)r * 4 * H;          // [2][B][4H]: this direction's own product only
        float* x = xp + ((size_t)t * B + b) * 8 * H + dir * 4 * H;
        const float gi = sigm(g[j] + x[j]), gj = tanhf(g[H + j] + x[H + j]);
        const float gf = sigm(g[2 * H + j] + x[2 * H + j] + 1.0f), go = sigm(g[3 * H + j] + x[3 * H + j]);
        const float cn = gf * c[idx] + gi * gj;
        const float hn = go * tanhf(cn);
        x[j] = gi; x[H + j] = gj; x[2 * H + j] = gf; x[3 * H + j] = go;
        c[idx] = cn;
        h[idx] = hn;
        const size_t o = ((size_t)t * B + b) * 2 * H + dir * H + j;
        out[o] = hn;
        cs[o] = cn;
    }
}

// One BPTT frame.  act [T*B, 8H]: gate activations in, gate PRE-activation gradients out (in place).  dh_rec
// [2][splits][B][H] holds the split-K partial products dG_step * W_h of the step processed just before; dh, dc [2B,H] carry the
// state gradients; dgs [2B, 4H] receives this step's gate gradients as the A operand of the next recurrent product
// (zero rows for examples that are past their length).
// GS = float: dgs float32 (TF32 product); GS = __nv_bfloat16: dgs bfloat16 (float32's exponent range -- gradients span many
// orders of magnitude -- with 8 mantissa bits; the 16-bit product, gemm_plan_dirs_h16).
template <typename GS>
__global__ void __launch_bounds__(256)
lstm_cell_bwd_kernel(float* __restrict__ act, const float* __restrict__ cs, const float* __restrict__ dout, const float* __restrict__ dh_rec,
                     const int32_t* __restrict__ seq_len, int s, int last, int T, int B, int H, int splits, float* __restrict__ dh,
                     float* __restrict__ dc, GS* __restrict__ dgs)
{
    auto cvt = [](float v) -> GS {
        if constexpr (std::is_same<GS, float>::value) return v;
        else return __float2bfloat16_rn(v);
    };
    // programmatic dependent launch: scheduled while the recurrent product of the previous step was finishing (see the frame loop)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // four consecutive units per thread: 16-byte accesses throughout (the scalar form took 7.5 us per frame at B = 256, a third of
    // the frame)
    const int H4 = H >> 2;
    const int total = 2 * B * H4;
    auto st4 = [&](GS* p, const float4 v) {
        if constexpr (std::is_same<GS, float>::value) {
            *reinterpret_cast<float4*>(p) = v;
        } else {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const unsigned*>(&lo), *reinterpret_cast<const unsigned*>(&hi));
        }
    };
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int j = (idx % H4) * 4;
        const int r = idx / H4;
        const int dir = r >= B, b = dir ? r - B : r;
        const int len = min(seq_len[b], T);
        GS* gs = dgs + (size_t)r * 4 * H;
        const size_t e = (size_t)r * H + j;          // element index in dh / dc
        // Every load of the frame is issued before the first use: the split-K partials as eight predicated loads (a loop over a
        // run-time `splits` kept one load in flight per trip), the activations of a live row beside them -- the kernel is a chain of
        // memory round trips (6.5 us per frame at B = 256 for 20 MB), and there were five of them in a row.
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool live = s < len, rec = !last && s + 1 < len;
        const int t = live ? (dir ? len - 1 - s : s) : 0;
        const size_t o = ((size_t)t * B + b) * 2 * H + dir * H + j;
        float* a = act + ((size_t)t * B + b) * 8 * H + dir * 4 * H;
        const float* pr = dh_rec + ((size_t)dir * splits * B + b) * H + j;     // [2][splits][B][H] split-K partials
        float4 part[8];
#pragma unroll
        for (int z = 0; z < 8; ++z) part[z] = (rec && z < splits) ? *reinterpret_cast<const float4*>(pr + (size_t)z * B * H) : zero;
        const float4 dcv = rec ? *reinterpret_cast<const float4*>(dc + e) : zero;
        float4 gi = zero, gj = zero, gf = zero, go = zero, cn = zero, dO = zero, cp = zero;
        if (live) {
            gi = *reinterpret_cast<const float4*>(a + j); gj = *reinterpret_cast<const float4*>(a + H + j);
            gf = *reinterpret_cast<const float4*>(a + 2 * H + j); go = *reinterpret_cast<const float4*>(a + 3 * H + j);
            cn = *reinterpret_cast<const float4*>(cs + o); dO = *reinterpret_cast<const float4*>(dout + o);
            if (s > 0) { const int tp = dir ? t + 1 : t - 1; cp = *reinterpret_cast<const float4*>(cs + ((size_t)tp * B + b) * 2 * H + dir * H + j); }
        }
        // gradient flowing into h_s from step s+1 (only if that step was live for this example): the partials in split order
        float4 dhv = zero;
#pragma unroll
        for (int z = 0; z < 8; ++z) { dhv.x += part[z].x; dhv.y += part[z].y; dhv.z += part[z].z; dhv.w += part[z].w; }
        if (!live) { st4(gs + j, zero); st4(gs + H + j, zero); st4(gs + 2 * H + j, zero); st4(gs + 3 * H + j, zero); continue; }
        float4 d_i, d_j, d_f, d_o, ndc;
#define OCR_CB1(f_)                                                   \
        {                                                             \
            const float tc = tanhf(cn.f_);                            \
            const float dht = dO.f_ + dhv.f_;                         \
            d_o.f_ = dht * tc * go.f_ * (1.f - go.f_);                \
            const float dct = dcv.f_ + dht * go.f_ * (1.f - tc * tc); \
            d_i.f_ = dct * gj.f_ * gi.f_ * (1.f - gi.f_);             \
            d_j.f_ = dct * gi.f_ * (1.f - gj.f_ * gj.f_);             \
            d_f.f_ = dct * cp.f_ * gf.f_ * (1.f - gf.f_);             \
            ndc.f_ = dct * gf.f_;                                     \
        }
        OCR_CB1(x) OCR_CB1(y) OCR_CB1(z) OCR_CB1(w)
#undef OCR_CB1
        *reinterpret_cast<float4*>(a + j) = d_i; *reinterpret_cast<float4*>(a + H + j) = d_j;
        *reinterpret_cast<float4*>(a + 2 * H + j) = d_f; *reinterpret_cast<float4*>(a + 3 * H + j) = d_o;
        st4(gs + j, d_i); st4(gs + H + j, d_j); st4(gs + 2 * H + j, d_f); st4(gs + 3 * H + j, d_o);
        *reinterpret_cast<float4*>(dc + e) = ndc;
        // (d h_t itself is not stored: the recurrent product is taken from dgs; the store was 1 MB per frame at B = 256)
    }
}

// ---------------------------------------------------------------------------------------------------------------
// GRU frames for training (tf.contrib.rnn.GRUCell, model.py:173-180): r,u = sigmoid([x,h] Wg + bg);
// c = tanh([x, r*h] Wc + bc); h' = u*h + (1-u)*c.  act [T*B, 6H] (per direction r | u | c) holds the input projections
// on entry and the activations afterwards; rh_all [T,B,2H] keeps r*h_{prev} (the operand of the candidate kernel's gradient).
__global__ void __launch_bounds__(256)
gru_gates_train_kernel(const float* __restrict__ gh /*[2][B][2H]*/, float* __restrict__ act, const int32_t* __restrict__ seq_len, int s, int T,
                       int B, int H, const float* __restrict__ h, float* __restrict__ rh /*[2B,H]*/, float* __restrict__ rh_all)
{
    const int total = 2 * B * H;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int j = idx % H;
        const int r = idx / H;
        const int dir = r >= B, b = dir ? r - B : r;
        const int len = min(seq_len[b], T);
        if (s >= len) { rh[idx] = 0.0f; continue; }
        const int t = dir ? len - 1 - s : s;
        const float* g = gh + (size_t)r * 2 * H;
        float* a = act + ((size_t)t * B + b) * 6 * H + dir * 3 * H;
        const float rr = sigm(g[j] + a[j]), uu = sigm(g[H + j] + a[H + j]);
        a[j] = rr; a[H + j] = uu;
        const float v = rr * h[idx];
        rh[idx] = v;
        rh_all[((size_t)t * B + b) * 2 * H + dir * H + j] = v;
    }
}
__global__ void __launch_bounds__(256)
gru_cell_train_kernel(const float* __restrict__ ch /*[2][B][H]*/, float* __restrict__ act, const int32_t* __restrict__ seq_len, int s, int T,
                      int B, int H, float* __restrict__ h, float* __restrict__ out)
{
    const int total = 2 * B * H;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int j = idx % H;
        const int r = idx / H;
        const int dir = r >= B, b = dir ? r - B : r;
        const int len = min(seq_len[b], T);
        if (s >= len) continue;
        const int t = dir ? len - 1 - s : s;
        float* a = act + ((size_t)t * B + b) * 6 * H + dir * 3 * H;
        const float cand = tanhf(ch[idx] + a[2 * H + j]);
        a[2 * H + j] = cand;
        const float uu = a[H + j];
        const float hn = uu * h[idx] + (1.0f - uu) * cand;
        h[idx] = hn;
        out[((size_t)t * B + b) * 2 * H + dir * H + j] = hn;
    }
}
// BPTT, first half of a frame: gradients of u and the candidate.  carry = d loss / d h_s from step s+1:
// dhd (through u*h) + dhr (through r*h) + the split-K partials of [dzr,dzu] * Wg_h.
__global__ void __launch_bounds__(256)
gru_bwd1_kernel(float* __restrict__ act, const float* __restrict__ out, const float* __restrict__ dout, const int32_t* __restrict__ seq_len, int s,
                int last, int T, int B, int H, float* __restrict__ dhd, const float* __restrict__ dhr, const float* __restrict__ dhg, int splits,
                float* __restrict__ dzc_step /*[2B,H]*/, float* __restrict__ dzg_step /*[2B,2H]*/, float* __restrict__ dht /*[2B,H]*/)
{
    const int total = 2 * B * H;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int j = idx % H;
        const int r = idx / H;
        const int dir = r >= B, b = dir ? r - B : r;
        const int len = min(seq_len[b], T);
        float carry = 0.f;
        if (!last && s + 1 < len) {
            carry = dhd[idx] + dhr[idx];
            const float* pr = dhg + ((size_t)dir * splits * B + b) * H + j;
            for (int z = 0; z < splits; ++z) carry += pr[(size_t)z * B * H];
        }
        if (s >= len) { dzc_step[idx] = 0.f; dzg_step[(size_t)r * 2 * H + H + j] = 0.f; dht[idx] = 0.f; continue; }
        const int t = dir ? len - 1 - s : s;
        const size_t o = ((size_t)t * B + b) * 2 * H + dir * H + j;
        float* a = act + ((size_t)t * B + b) * 6 * H + dir * 3 * H;
        const float uu = a[H + j], cand = a[2 * H + j];
        float hprev = 0.f;
        if (s > 0) { const int tp = dir ? t + 1 : t - 1; hprev = out[((size_t)tp * B + b) * 2 * H + dir * H + j]; }
        const float d = dout[o] + carry;
        const float dzc = d * (1.f - uu) * (1.f - cand * cand);
        const float dzu = d * (hprev - cand) * uu * (1.f - uu);
        a[H + j] = dzu; a[2 * H + j] = dzc;
        dzc_step[idx] = dzc;
        dzg_step[(size_t)r * 2 * H + H + j] = dzu;
        dhd[idx] = d * uu;
        dht[idx] = hprev;      // kept for the second half
    }
}
// second half: through r*h.  drh [2][splits][B][H] = split-K partials of dzc * Wc_h
__global__ void __launch_bounds__(256)
gru_bwd2_kernel(float* __restrict__ act, const float* __restrict__ drh, int splits, const int32_t* __restrict__ seq_len, int s, int T, int B, int H,
                const float* __restrict__ hprev, float* __restrict__ dhr, float* __restrict__ dzg_step)
{
    const int total = 2 * B * H;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int j = idx % H;
        const int r = idx / H;
        const int dir = r >= B, b = dir ? r - B : r;
        const int len = min(seq_len[b], T);
        if (s >= len) { dzg_step[(size_t)r * 2 * H + j] = 0.f; continue; }
        const int t = dir ? len - 1 - s : s;
        float* a = act + ((size_t)t * B + b) * 6 * H + dir * 3 * H;
        float d = 0.f;
        const float* pr = drh + ((size_t)dir * splits * B + b) * H + j;
        for (int z = 0; z < splits; ++z) d += pr[(size_t)z * B * H];
        const float rr = a[j];
        const float dzr = d * hprev[idx] * rr * (1.f - rr);
        a[j] = dzr;
        dzg_step[(size_t)r * 2 * H + j] = dzr;
        dhr[idx] = d * rr;
    }
}

// rows (t, b) with t >= len_b were never visited: their slots still hold input-projection values; their gradient is 0.
// One warp per row: a row inside its sequence costs one length compare (the element-wise form walked all T*B*cols4 float4s
// -- 78 us per layer at B = 256 -- to find, usually, nothing to zero).
__global__ void __launch_bounds__(256)
zero_past_len_kernel(float* __restrict__ a, const int32_t* __restrict__ seq_len, int T, int B, int cols4)
{
    const int lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    const long long rows = (long long)T * B;
    for (long long row = (long long)blockIdx.x * wpc + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpc) {
        const int b = (int)(row % B), t = (int)(row / B);
        if (t < seq_len[b]) continue;
        float4* p = reinterpret_cast<float4*>(a) + row * cols4;
        for (int c = lane; c < cols4; c += 32) p[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// conv1 weight gradient: dw[tap, co] = sum_pixels dy[pix, co] * img[pix + tap].  One input channel, so this is a
// reduction, not a contraction: lane = output channel (blockIdx.y = 32-channel group), a warp walks pixels.
template <bool kU8>
__global__ void __launch_bounds__(256)
conv1_wgrad_kernel(const void* __restrict__ in_, int B, int H, int W, const float* __restrict__ dy, int Co, double* __restrict__ sums /*[9][Co]*/)
{
    // per iteration the CTA stages the 3x3 neighbourhoods of 256 consecutive output pixels in shared memory (one pixel
    // per thread, coalesced byte loads), then every warp sweeps 32 of them: lane = filter, dy row coalesced (128 B),
    // image taps broadcast from shared memory
    __shared__ float taps[9][256 + 1];
    __shared__ float red[8][9][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int co = blockIdx.y * 32 + lane;
    const int Ho = H - 2, Wo = W - 2;
    const long long npix = (long long)B * Ho * Wo;
    float acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.f;
    for (long long p0 = (long long)blockIdx.x * 256; p0 < npix; p0 += (long long)gridDim.x * 256) {
        const long long p = p0 + threadIdx.x;
        if (p < npix) {
            const unsigned pp = (unsigned)p;
            const unsigned q = pp / (unsigned)Wo, x = pp - q * Wo;
            const unsigned b = q / (unsigned)Ho, y = q - b * Ho;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int jx = 0; jx < 3; ++jx) {
                    const size_t o = ((size_t)b * H + (y + i)) * W + (x + jx);
                    float v;
                    if (kU8) v = (float)__ldg(reinterpret_cast<const unsigned char*>(in_) + o) * (float)(1.0 / 255.0) - 0.5f;
                    else v = __ldg(reinterpret_cast<const float*>(in_) + o);
                    taps[i * 3 + jx][threadIdx.x] = v;
                }
        }
        __syncthreads();
        const int n = (int)min((long long)256, npix - p0);
        float g[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {    // eight dy rows in flight per warp
            const int pl = warp * 32 + u;
            g[u] = (pl < n && co < Co) ? dy[(size_t)(p0 + pl) * Co + co] : 0.f;
        }
#pragma unroll
        for (int u0 = 0; u0 < 32; u0 += 8) {
            float gn[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int pl = warp * 32 + u0 + 8 + u;
                gn[u] = (u0 + 8 < 32 && pl < n && co < Co) ? dy[(size_t)(p0 + pl) * Co + co] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int pl = warp * 32 + u0 + u;
                if (pl < n) {
#pragma unroll
                    for (int k = 0; k < 9; ++k) acc[k] = fmaf(g[u], taps[k][pl], acc[k]);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) g[u] = gn[u];
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) red[warp][k][lane] = acc[k];
    __syncthreads();
    for (int k = warp; k < 9; k += 8) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += (double)red[w][k][lane];
        if (co < Co) atomicAdd(sums + (size_t)k * Co + co, t);
    }
}

// filter layouts derived from the master HWIO tensor w [3,3,C,Co]:
//   w_fwd [Co, 9*C]:  w_fwd[co, tap*C + c]        = w[tap, c, co]     (K-major operand of the forward implicit GEMM)
//   w_dgr [C, 9*Co]:  w_dgr[c, (8-tap)*Co + co]   = w[tap, c, co]     (the input gradient is a 3x3 'same' convolution of dy
//                                                                        with the filter rotated by 180 degrees)
__global__ void __launch_bounds__(256)
conv_filter_layouts_kernel(const float* __restrict__ w, int C, int Co, float* __restrict__ w_fwd, float* __restrict__ w_dgr)
{
    const int total = 9 * C * Co;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int co = idx % Co;
        const int c = (idx / Co) % C;
        const int tap = idx / (Co * C);
        const float v = w[idx];
        if (w_fwd) w_fwd[(size_t)co * 9 * C + tap * C + c] = v;
        if (w_dgr) w_dgr[(size_t)c * 9 * Co + (8 - tap) * Co + co] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// tf.train.AdamOptimizer over a flat buffer: m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr_t m / (sqrt(v) + eps),
// lr_t = lr sqrt(1-b2^t) / (1-b1^t) computed by the host.  float4, grid-stride.
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n, float lr_t,
            const float* __restrict__ lr_t_dev, float b1, float b2, float eps, float gscale)
{
    if (lr_t_dev != nullptr) lr_t = *lr_t_dev;    // a captured CUDA graph replays with a fresh step size
    const long long n4 = n >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
#define OCR_ADAM1(f)                                              \
        {                                                         \
            const float gr = gg.f * gscale;                       \
            mm.f = b1 * mm.f + (1.f - b1) * gr;                   \
            vv.f = b2 * vv.f + (1.f - b2) * gr * gr;              \
            pp.f -= lr_t * mm.f / (sqrtf(vv.f) + eps);            \
        }
        OCR_ADAM1(x) OCR_ADAM1(y) OCR_ADAM1(z) OCR_ADAM1(w)
#undef OCR_ADAM1
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    // tail
    for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gr = g[i] * gscale;
        const float mn = b1 * m[i] + (1.f - b1) * gr, vn = b2 * v[i] + (1.f - b2) * gr * gr;
        m[i] = mn; v[i] = vn;
        p[i] -= lr_t * mn / (sqrtf(vn) + eps);
    }
}

// dz = g * (z > 0) for a ReLU whose OUTPUT z is given (logits layer), float4
__global__ void __launch_bounds__(256)
relu_bwd_kernel(const float* __restrict__ z, const float* __restrict__ g, long long n, float* __restrict__ dz)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dz[i] = z[i] > 0.f ? g[i] : 0.f;
}

}  // namespace ocr

using namespace ocr;

#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int ocr_transpose(const float* in, long long rows, int cols, int ld_in, float* out, long long ld_out, long long src_shift,
                             ocr_stream_t stream)
{
    OCR_CHECK_ARG(rows >= 0 && cols >= 1 && ld_in >= cols && ld_out >= rows, "ocr_transpose: bad shape rows=%lld cols=%d ld_in=%d ld_out=%lld", rows, cols, ld_in, ld_out);
    if (rows == 0) return OCR_OK;
    OCR_CHECK_ARG(in && out, "ocr_transpose: NULL argument");
    dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32));
    transpose_kernel<false><<<grid, 256, 0, ST(stream)>>>(in, rows, cols, ld_in, out, ld_out, 0, 0, 0, src_shift, 0);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

extern "C" int ocr_planar_pad_pitch(int W) { return (W + 2 + 3) & ~3; }

extern "C" int ocr_nhwc_to_planar_pad(const float* in, int B, int H, int W, int C, float* out, long long ld_out, int ncopies,
                                      long long copy_stride, ocr_stream_t stream)
{
    const int Wp = ocr_planar_pad_pitch(W);
    const long long rows = (long long)B * (H + 2) * Wp;
    OCR_CHECK_ARG(rows < 0x7fffffffLL, "ocr_nhwc_to_planar_pad: too many pixels");
    OCR_CHECK_ARG(B >= 0 && H >= 1 && W >= 1 && C >= 1 && ld_out >= rows && (ncopies == 1 || ncopies == 3) && (ncopies == 1 || copy_stride >= (long long)C * ld_out),
                  "ocr_nhwc_to_planar_pad: bad shape B=%d H=%d W=%d C=%d ld=%lld ncopies=%d", B, H, W, C, ld_out, ncopies);
    if (B == 0) return OCR_OK;
    OCR_CHECK_ARG(in && out, "ocr_nhwc_to_planar_pad: NULL argument");
    if ((C % 4) == 0 && ((uintptr_t)in % 16) == 0) {
        dim3 grid((unsigned)((rows + 127) / 128), (unsigned)((C + 31) / 32));
        if (ncopies == 3) planar_pad_kernel<3><<<grid, 256, 0, ST(stream)>>>(in, (unsigned)rows, C, out, ld_out, H, W, Wp, copy_stride);
        else planar_pad_kernel<1><<<grid, 256, 0, ST(stream)>>>(in, (unsigned)rows, C, out, ld_out, H, W, Wp, copy_stride);
    } else {
        dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)ncopies);
        transpose_kernel<true><<<grid, 256, 0, ST(stream)>>>(in, rows, C, C, out, ld_out, H, W, Wp, ncopies == 3 ? -1 : 0, copy_stride);
    }
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

// K-blocked form of the planar copies (ocr_gemm_tf32_wgrad_blocked): pitch = W + 1 rounded up to 32 -- ONE zero column between
// image rows serves as the right ring of one row and the left ring of the next, and the vertical taps (+- one pitch) are whole
// blocks of 32 pixels.  out: [R / 32][rows_total][32] with R = B * (H + 2) * pitch; copy k (pixel shift k - 1; ncopies 1 or 3)
// occupies rows row0 + k * C .. of the rows_total rows.
extern "C" int ocr_planar_pad_pitch32(int W) { return (W + 1 + 31) & ~31; }

extern "C" int ocr_nhwc_to_planar_blocked(const float* in, int B, int H, int W, int C, float* out, int rows_total, int row0, int ncopies,
                                          ocr_stream_t stream)
{
    const int Wp = ocr_planar_pad_pitch32(W);
    const long long rows = (long long)B * (H + 2) * Wp;
    OCR_CHECK_ARG(rows < 0x7fffffffLL, "ocr_nhwc_to_planar_blocked: too many pixels");
    OCR_CHECK_ARG(B >= 0 && H >= 1 && W >= 1 && C >= 4 && (C % 4) == 0 && (ncopies == 1 || ncopies == 3) && row0 >= 0 && row0 + ncopies * C <= rows_total,
                  "ocr_nhwc_to_planar_blocked: bad shape B=%d H=%d W=%d C=%d rows_total=%d row0=%d ncopies=%d", B, H, W, C, rows_total, row0, ncopies);
    if (B == 0) return OCR_OK;
    OCR_CHECK_ARG(in && out && ((uintptr_t)in % 16) == 0, "ocr_nhwc_to_planar_blocked: NULL or misaligned argument");
    dim3 grid((unsigned)((rows + 127) / 128), (unsigned)((C + 31) / 32));
    float* o = out + (size_t)row0 * 32;
    if (ncopies == 3) planar_pad_kernel<3, true><<<grid, 256, 0, ST(stream)>>>(in, (unsigned)rows, C, o, rows_total, H, W, Wp, C);
    else planar_pad_kernel<1, true><<<grid, 256, 0, ST(stream)>>>(in, (unsigned)rows, C, o, rows_total, H, W, Wp, C);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

extern "C" int ocr_gemm_tf32_wgrad_blocked(const float* At, long long a_rows, const float* Wt, long long w_rows, float* D, int ldd,
                                           long long batch_stride, int M, int N, long long R, int nbatch, const int32_t* a_shift_host,
                                           const int32_t* a_row_host, void* scratch, size_t scratch_bytes, ocr_stream_t stream)
{
    OCR_CHECK_ARG(M >= 1 && N >= 1 && R >= 1 && nbatch >= 1 && nbatch <= 9 && a_rows >= M && w_rows >= N, "ocr_gemm_tf32_wgrad_blocked: bad shape");
    if (scratch == nullptr || scratch_bytes < gemm_wgrad_scratch_floats(M, N, R, nbatch) * sizeof(float)) {
        set_error("ocr_gemm_tf32_wgrad_blocked: scratch too small");
        return OCR_EWORKSPACE;
    }
    int sh[9] = {0}, ar[9] = {0};
    for (int i = 0; i < nbatch; ++i) {
        sh[i] = a_shift_host ? a_shift_host[i] : 0;
        ar[i] = a_row_host ? a_row_host[i] : 0;
        OCR_CHECK_ARG((sh[i] % 32) == 0, "ocr_gemm_tf32_wgrad_blocked: a_shift[%d] = %d is not a multiple of 32", i, sh[i]);
        OCR_CHECK_ARG(ar[i] >= 0 && ar[i] + M <= a_rows, "ocr_gemm_tf32_wgrad_blocked: a_row[%d] = %d outside the operand", i, ar[i]);
    }
    return gemm_wgrad(At, 0, Wt, 0, D, ldd, batch_stride, M, N, R, nbatch, sh, ar, a_rows, reinterpret_cast<float*>(scratch), ST(stream), 1, w_rows);
}

extern "C" int ocr_gemm_wgrad_scratch_bytes(int M, int N, long long R, int nbatch, size_t* bytes)
{
    OCR_CHECK_ARG(bytes && M >= 1 && N >= 1 && R >= 1 && nbatch >= 1 && nbatch <= 9, "ocr_gemm_wgrad_scratch_bytes: bad argument");
    *bytes = gemm_wgrad_scratch_floats(M, N, R, nbatch) * sizeof(float);
    return OCR_OK;
}

extern "C" int ocr_gemm_tf32_wgrad(const float* At, long long lda, const float* Wt, long long ldw, float* D, int ldd, long long batch_stride,
                                   int M, int N, long long R, int nbatch, const int32_t* a_shift_host, const int32_t* a_row_host,
                                   long long a_rows, void* scratch, size_t scratch_bytes, ocr_stream_t stream)
{
    OCR_CHECK_ARG(M >= 1 && N >= 1 && R >= 1 && nbatch >= 1 && nbatch <= 9, "ocr_gemm_tf32_wgrad: bad shape");
    if (scratch == nullptr || scratch_bytes < gemm_wgrad_scratch_floats(M, N, R, nbatch) * sizeof(float)) {
        set_error("ocr_gemm_tf32_wgrad: scratch too small");
        return OCR_EWORKSPACE;
    }
    int sh[9] = {0}, ar[9] = {0};
    for (int i = 0; i < nbatch; ++i) {
        sh[i] = a_shift_host ? a_shift_host[i] : 0;
        ar[i] = a_row_host ? a_row_host[i] : 0;
        OCR_CHECK_ARG((sh[i] % 4) == 0, "ocr_gemm_tf32_wgrad: a_shift[%d] = %d is not a multiple of 4 (TMA needs 16-byte aligned box origins)", i, sh[i]);
        OCR_CHECK_ARG(ar[i] >= 0 && ar[i] + M <= (a_rows > 0 ? a_rows : M), "ocr_gemm_tf32_wgrad: a_row[%d] = %d outside the operand", i, ar[i]);
    }
    return gemm_wgrad(At, lda, Wt, ldw, D, ldd, batch_stride, M, N, R, nbatch, sh, ar, a_rows > 0 ? a_rows : M, reinterpret_cast<float*>(scratch), ST(stream));
}

template <int MODE>
static int launch_reduce(const float* x, const float* g, long long rows, int C, int ldx, const float* mean, const float* inv_std,
                         const float* gamma, const float* beta, float* out, double* sums, cudaStream_t st)
{
    OCR_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
    const int gy = (C + 31) / 32;
    long long gx = (rows + 7) / 8;
    const long long cap = (148LL * 8 + gy - 1) / gy;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    channel_reduce_kernel<MODE><<<dim3((unsigned)gx, (unsigned)gy), 256, 0, st>>>(x, g, rows, C, ldx, mean, inv_std, gamma, beta, out, sums);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

// sums [2][C] doubles: per-channel sum and sum of squares of THIS replica's rows.  A data-parallel job that wants the
// statistics of the global batch all-reduces `sums` (and adds up the row counts) before ocr_bn_finalize.
extern "C" int ocr_bn_batch_sums(const float* y, long long rows, int C, void* sums, ocr_stream_t stream)
{
    OCR_CHECK_ARG(rows >= 1 && C >= 1 && y && sums, "ocr_bn_batch_sums: bad argument");
    return launch_reduce<0>(y, nullptr, rows, C, C, nullptr, nullptr, nullptr, nullptr, nullptr, reinterpret_cast<double*>(sums), ST(stream));
}

extern "C" int ocr_bn_finalize(const void* sums, long long n, int C, float eps, float momentum, float* mean, float* inv_std,
                               float* moving_mean, float* moving_var, ocr_stream_t stream)
{
    OCR_CHECK_ARG(n >= 1 && C >= 1 && sums && mean && inv_std && ((moving_mean == nullptr) == (moving_var == nullptr)), "ocr_bn_finalize: bad argument");
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, ST(stream)>>>(reinterpret_cast<const double*>(sums), n, C, eps, momentum, mean, inv_std, moving_mean, moving_var);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

extern "C" int ocr_bn_relu_apply(const float* y, long long rows, int C, const float* mean, const float* inv_std, const float* gamma,
                                 const float* beta, float* out, ocr_stream_t stream)
{
    OCR_CHECK_ARG(rows >= 1 && C >= 4 && (C % 4) == 0 && y && mean && inv_std && gamma && beta && out, "ocr_bn_relu_apply: bad argument");
    bn_relu_apply_kernel<<<grid_cap(rows * (C / 4)), 256, 0, ST(stream)>>>(y, rows, C, mean, inv_std, gamma, beta, out);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

// ocr_bn_relu_apply followed by ocr_maxpool(2, 2, 2, stride_w) in one pass: out [B,H,W,C] = relu(bn(y)) (kept for the pool's
// gradient), pooled [B, (H-2)/2+1, (W-2)/stride_w+1, C]
extern "C" int ocr_bn_relu_apply_pool(const float* y, int B, int H, int W, int C, const float* mean, const float* inv_std, const float* gamma,
                                      const float* beta, float* out, int stride_w, float* pooled, ocr_stream_t stream)
{
    OCR_CHECK_ARG(B >= 1 && H >= 2 && W >= 2 && C >= 4 && (C % 4) == 0 && (stride_w == 1 || stride_w == 2) && y && mean && inv_std && gamma && beta && out && pooled,
                  "ocr_bn_relu_apply_pool: bad argument");
    const int Hp = (H - 2) / 2 + 1, Wp = (W - 2) / stride_w + 1;
    bn_relu_apply_pool_kernel<<<grid_cap((long long)B * H * W * (C / 4)), 256, 0, ST(stream)>>>(y, B, H, W, C, mean, inv_std, gamma, beta, out, stride_w, Hp, Wp, pooled);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

// sums [2][C] doubles out: sum dz and sum dz*xhat over this replica's rows (dz = dout where the ReLU was active);
// dgamma = sum dz*xhat, dbeta = sum dz (this replica's share of the parameter gradients).
extern "C" int ocr_bn_relu_bwd_sums(const float* y, const float* dout, long long rows, int C, const float* mean, const float* inv_std,
                                    const float* gamma, const float* beta, void* sums, float* dgamma, float* dbeta, ocr_stream_t stream)
{
    OCR_CHECK_ARG(rows >= 1 && C >= 1 && y && dout && mean && inv_std && gamma && beta && dgamma && dbeta && sums, "ocr_bn_relu_bwd_sums: bad argument");
    int rc = launch_reduce<2>(y, dout, rows, C, C, mean, inv_std, gamma, beta, nullptr, reinterpret_cast<double*>(sums), ST(stream));
    if (rc != OCR_OK) return rc;
    sums_to_float_kernel<<<(C + 127) / 128, 128, 0, ST(stream)>>>(reinterpret_cast<const double*>(sums), C, dbeta, dgamma);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

// dy = gamma * inv_std * (dz - sums[0]/n - xhat * sums[1]/n); n = rows the statistics were taken over (global batch when
// `sums` was all-reduced)
extern "C" int ocr_bn_relu_bwd_apply(const float* y, const float* dout, long long rows, long long n, int C, const float* mean,
                                     const float* inv_std, const float* gamma, const float* beta, const void* sums, float* dy,
                                     ocr_stream_t stream)
{
    OCR_CHECK_ARG(rows >= 1 && n >= rows && C >= 1 && y && dout && mean && inv_std && gamma && beta && sums && dy, "ocr_bn_relu_bwd_apply: bad argument");
    OCR_CHECK_ARG((C % 4) == 0, "ocr_bn_relu_bwd_apply: C must be a multiple of 4");
    bn_relu_bwd_apply_kernel<false><<<grid_cap(rows * (C / 4)), 256, 0, ST(stream)>>>(y, dout, rows, n, C, mean, inv_std, gamma, beta, reinterpret_cast<const double*>(sums), dy, nullptr);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

// the same, and dbias [C] = per-channel sums of dy (what ocr_colsum(dy) gives: the bias gradient of the convolution whose output
// the batch-norm normalises) from the same pass; scratch: C doubles.  Needs 256 % (C / 4) == 0 (C = 32 .. 1024 in powers of two).
extern "C" int ocr_bn_relu_bwd_apply_bias(const float* y, const float* dout, long long rows, long long n, int C, const float* mean,
                                          const float* inv_std, const float* gamma, const float* beta, const void* sums, float* dy,
                                          float* dbias, void* scratch, ocr_stream_t stream)
{
    OCR_CHECK_ARG(rows >= 1 && n >= rows && C >= 4 && y && dout && mean && inv_std && gamma && beta && sums && dy && dbias && scratch, "ocr_bn_relu_bwd_apply_bias: bad argument");
    OCR_CHECK_ARG((C % 4) == 0 && (256 % (C / 4)) == 0, "ocr_bn_relu_bwd_apply_bias: C / 4 must divide 256 (C = %d)", C);
    double* ds = reinterpret_cast<double*>(scratch);
    OCR_CHECK_CUDA(cudaMemsetAsync(ds, 0, sizeof(double) * C, ST(stream)));
    bn_relu_bwd_apply_kernel<true><<<grid_cap(rows * (C / 4), 256, 4), 256, 0, ST(stream)>>>(y, dout, rows, n, C, mean, inv_std, gamma, beta, reinterpret_cast<const double*>(sums), dy, ds);
    OCR_CHECK_LAUNCH();
    sums_to_float_kernel<<<(C + 127) / 128, 128, 0, ST(stream)>>>(ds, C, dbias, nullptr);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

// ---- the pooled batch-norm layers without the full-size activation and without a pool-gradient tensor (see bn_relu_pool_arg_kernel)
#define OCR_POOL_ARG_SHAPE(fn)                                                                                                              \
    OCR_CHECK_ARG(B >= 1 && H >= 2 && W >= 2 && C >= 4 && (C % 4) == 0 && (256 % (C / 4)) == 0 && (stride_w == 1 || stride_w == 2) &&       \
                  (long long)B * H * W < 0x7fffffffLL, fn ": bad shape B=%d H=%d W=%d C=%d stride_w=%d (C / 4 must divide 256)", B, H, W, C, stride_w)

extern "C" int ocr_bn_relu_apply_pool_arg(const float* y, int B, int H, int W, int C, const float* mean, const float* inv_std, const float* gamma,
                                          const float* beta, int stride_w, float* pooled, void* arg, ocr_stream_t stream)
{
    OCR_POOL_ARG_SHAPE("ocr_bn_relu_apply_pool_arg");
    OCR_CHECK_ARG(y && mean && inv_std && gamma && beta && pooled && arg, "ocr_bn_relu_apply_pool_arg: NULL argument");
    const int Hp = (H - 2) / 2 + 1, Wp = (W - 2) / stride_w + 1;
    bn_relu_pool_arg_kernel<<<grid_cap((long long)B * Hp * Wp * (C / 4)), 256, 0, ST(stream)>>>(y, B, H, W, C, mean, inv_std, gamma, beta, stride_w, Hp, Wp, pooled,
                                                                                            reinterpret_cast<unsigned char*>(arg));
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

extern "C" int ocr_bn_relu_bwd_sums_pool(const float* y, const float* dpooled, const void* arg, int B, int H, int W, int C, int stride_w,
                                         const float* mean, const float* inv_std, const float* gamma, const float* beta, void* sums,
                                         float* dgamma, float* dbeta, ocr_stream_t stream)
{
    OCR_POOL_ARG_SHAPE("ocr_bn_relu_bwd_sums_pool");
    OCR_CHECK_ARG(y && dpooled && arg && mean && inv_std && gamma && beta && sums && dgamma && dbeta, "ocr_bn_relu_bwd_sums_pool: NULL argument");
    const int Hp = (H - 2) / 2 + 1, Wp = (W - 2) / stride_w + 1;
    OCR_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, ST(stream)));
    bn_bwd_sums_pool_kernel<<<grid_cap((long long)B * H * W * (C / 4), 256, 4), 256, 0, ST(stream)>>>(
        y, dpooled, reinterpret_cast<const unsigned char*>(arg), B, H, W, C, stride_w, Hp, Wp, mean, inv_std, gamma, beta, reinterpret_cast<double*>(sums));
    OCR_CHECK_LAUNCH();
    sums_to_float_kernel<<<(C + 127) / 128, 128, 0, ST(stream)>>>(reinterpret_cast<const double*>(sums), C, dbeta, dgamma);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

extern "C" int ocr_bn_relu_bwd_apply_bias_pool(const float* y, const float* dpooled, const void* arg, int B, int H, int W, int C, int stride_w,
                                               long long n, const float* mean, const float* inv_std, const float* gamma, const float* beta,
                                               const void* sums, float* dy, float* dbias, void* scratch, ocr_stream_t stream)
{
    OCR_POOL_ARG_SHAPE("ocr_bn_relu_bwd_apply_bias_pool");
    OCR_CHECK_ARG(n >= (long long)B * H * W && y && dpooled && arg && mean && inv_std && gamma && beta && sums && dy && dbias && scratch,
                  "ocr_bn_relu_bwd_apply_bias_pool: bad argument");
    const int Hp = (H - 2) / 2 + 1, Wp = (W - 2) / stride_w + 1;
    double* ds = reinterpret_cast<double*>(scratch);
    OCR_CHECK_CUDA(cudaMemsetAsync(ds, 0, sizeof(double) * C, ST(stream)));
    bn_relu_bwd_apply_pool_kernel<<<grid_cap((long long)B * H * W * (C / 4), 256, 4), 256, 0, ST(stream)>>>(
        y, dpooled, reinterpret_cast<const unsigned char*>(arg), B, H, W, C, stride_w, Hp, Wp, n, mean, inv_std, gamma, beta,
        reinterpret_cast<const double*>(sums), dy, ds);
    OCR_CHECK_LAUNCH();
    sums_to_float_kernel<<<(C + 127) / 128, 128, 0, ST(stream)>>>(ds, C, dbias, nullptr);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}
#undef OCR_POOL_ARG_SHAPE

extern "C" int ocr_copy_2d(const float* src, long long ld_src, float* dst, long long ld_dst, long long rows, long long cols, ocr_stream_t stream)
{
    OCR_CHECK_ARG(rows >= 0 && cols >= 0 && ld_src >= cols && ld_dst >= cols, "ocr_copy_2d: bad shape");
    if (rows == 0 || cols == 0) return OCR_OK;
    OCR_CHECK_ARG(src && dst, "ocr_copy_2d: NULL argument");
    OCR_CHECK_CUDA(cudaMemcpy2DAsync(dst, (size_t)ld_dst * 4, src, (size_t)ld_src * 4, (size_t)cols * 4, (size_t)rows, cudaMemcpyDeviceToDevice, ST(stream)));
    return OCR_OK;
}

extern "C" int ocr_relu_bwd_bias(const float* out, const float* dout, long long rows, int C, float* dy, float* dbias, void* scratch,
                                 ocr_stream_t stream)
{
    OCR_CHECK_ARG(rows >= 1 && C >= 1 && out && dout && dy && dbias && scratch, "ocr_relu_bwd_bias: bad argument");
    double* sums = reinterpret_cast<double*>(scratch);
    int rc = launch_reduce<3>(out, dout, rows, C, C, nullptr, nullptr, nullptr, nullptr, dy, sums, ST(stream));
    if (rc != OCR_OK) return rc;
    sums_to_float_kernel<<<(C + 127) / 128, 128, 0, ST(stream)>>>(sums, C, dbias, nullptr);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

extern "C" int ocr_colsum(const float* x, long long rows, int C, int ldx, float* out, void* scratch, ocr_stream_t stream)
{
    OCR_CHECK_ARG(rows >= 1 && C >= 1 && ldx >= C && x && out && scratch, "ocr_colsum: bad argument");
    double* sums = reinterpret_cast<double*>(scratch);
    int rc = launch_reduce<1>(x, nullptr, rows, C, ldx, nullptr, nullptr, nullptr, nullptr, nullptr, sums, ST(stream));
    if (rc != OCR_OK) return rc;
    sums_to_float_kernel<<<(C + 127) / 128, 128, 0, ST(stream)>>>(sums, C, out, nullptr);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

extern "C" int ocr_relu_bwd(const float* z, const float* g, long long n, float* dz, ocr_stream_t stream)
{
    OCR_CHECK_ARG(n >= 0 && (n == 0 || (z && g && dz)), "ocr_relu_bwd: bad argument");
    if (n == 0) return OCR_OK;
    relu_bwd_kernel<<<grid_cap(n), 256, 0, ST(stream)>>>(z, g, n, dz);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

extern "C" int ocr_maxpool_bwd(const float* in, const float* dout, int B, int H, int W, int C, int pool_h, int pool_w, int stride_h, int stride_w,
                               float* din, ocr_stream_t stream)
{
    OCR_CHECK_ARG(B >= 1 && C >= 1 && pool_h >= 1 && pool_w >= 1 && stride_h >= 1 && stride_w >= 1 && H >= pool_h && W >= pool_w && in && dout && din,
                  "ocr_maxpool_bwd: bad argument");
    OCR_CHECK_ARG((C % 4) == 0 && (long long)B * H * W < 0x7fffffffLL, "ocr_maxpool_bwd: C must be a multiple of 4");
    const int Hp = (H - pool_h) / stride_h + 1, Wp = (W - pool_w) / stride_w + 1;
    maxpool_bwd_kernel<<<grid_cap((long long)B * H * W * (C / 4)), 256, 0, ST(stream)>>>(in, dout, B, H, W, C, pool_h, pool_w, stride_h, stride_w, Hp, Wp, din);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

extern "C" int ocr_rows_max_to_seq_bwd(const float* in, const float* dseq, int B, int H, int W, int C, float* din, ocr_stream_t stream)
{
    OCR_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && C >= 1 && in && dseq && din, "ocr_rows_max_to_seq_bwd: bad argument");
    rows_max_bwd_kernel<<<grid_cap((long long)B * W * C), 256, 0, ST(stream)>>>(in, dseq, B, H, W, C, din);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

extern "C" int ocr_conv1_wgrad(const void* in, int in_is_u8, int B, int H, int W, const float* dy, int Cout, float* dw, void* scratch,
                               ocr_stream_t stream)
{
    OCR_CHECK_ARG(B >= 1 && H >= 3 && W >= 3 && Cout >= 1 && in && dy && dw && scratch, "ocr_conv1_wgrad: bad argument");
    double* sums = reinterpret_cast<double*>(scratch);
    OCR_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * 9 * Cout, ST(stream)));
    const int gy = (Cout + 31) / 32;
    const long long npix = (long long)B * (H - 2) * (W - 2);
    long long gx = (npix + 255) / 256;
    if (gx > 148 * 6) gx = 148 * 6;
    OCR_CHECK_ARG(npix < 0x7fffffffLL, "ocr_conv1_wgrad: too many pixels");
    if (in_is_u8) conv1_wgrad_kernel<true><<<dim3((unsigned)gx, (unsigned)gy), 256, 0, ST(stream)>>>(in, B, H, W, dy, Cout, sums);
    else conv1_wgrad_kernel<false><<<dim3((unsigned)gx, (unsigned)gy), 256, 0, ST(stream)>>>(in, B, H, W, dy, Cout, sums);
    OCR_CHECK_LAUNCH();
    sums_to_float_kernel<<<(9 * Cout + 127) / 128, 128, 0, ST(stream)>>>(sums, 9 * Cout, dw, nullptr);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

extern "C" int ocr_conv_filter_layouts(const float* w_hwio, int C, int Cout, float* w_fwd, float* w_dgrad, ocr_stream_t stream)
{
    OCR_CHECK_ARG(C >= 1 && Cout >= 1 && w_hwio && (w_fwd || w_dgrad), "ocr_conv_filter_layouts: bad argument");
    conv_filter_layouts_kernel<<<grid_cap(9LL * C * Cout), 256, 0, ST(stream)>>>(w_hwio, C, Cout, w_fwd, w_dgrad);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

extern "C" int ocr_adam_step(float* params, const float* grads, float* m, float* v, long long n, float lr_t, const float* lr_t_device,
                             float beta1, float beta2, float eps, float grad_scale, ocr_stream_t stream)
{
    OCR_CHECK_ARG(n >= 0 && (n == 0 || (params && grads && m && v)), "ocr_adam_step: bad argument");
    OCR_CHECK_ARG(((uintptr_t)params % 16) == 0 && ((uintptr_t)grads % 16) == 0 && ((uintptr_t)m % 16) == 0 && ((uintptr_t)v % 16) == 0,
                  "ocr_adam_step: buffers must be 16-byte aligned");
    if (n == 0) return OCR_OK;
    adam_kernel<<<grid_cap((n + 3) / 4, 256, 8), 256, 0, ST(stream)>>>(params, grads, m, v, n, lr_t, lr_t_device, beta1, beta2, eps, grad_scale);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

namespace ocr {  // lstm_persistent.cu / model_ops.cu
bool lstm_persistent_supported(int T, int B, int H);
size_t lstm_persistent_workspace_floats(int B, int H);
int lstm_persistent_run(const float* xp, const float* wh, const float* wh_perm, const int32_t* seq_len, int T, int B, int H,
                        float* out, float* ws, cudaStream_t st, float* gates_out, float* cs_out);
extern int g_birnn_path;
// lstm_bptt_persistent.cu
bool lstm_bptt_supported(int T, int B, int H);
size_t lstm_bptt_workspace_floats(int B, int H);
int lstm_bptt_run(const float* dout, int T, int B, int H, const int32_t* seq_len, float* act, const float* cstate, const float* wh_rows,
                  float* ws, cudaStream_t st);
}

// tile width of the per-frame recurrent products: enough CTAs for the 148 SMs, as wide as that allows
static int recurrent_bn(int M, int N, int splits) {
    const long long mt = (M + kGemmBM - 1) / kGemmBM;
    int bn = 32;
    if (N > 32) bn = 64;
    if (N > 64 && 2 * mt * ((N + 127) / 128) * splits >= 120) bn = 128;
    if (N > 128 && 2 * mt * ((N + 255) / 256) * splits >= 120) bn = 256;
    return bn;
}

// ---------------------------------------------------------------------------------------------------------------
// bidirectional LSTM layer, training form (frame by frame; keeps gate activations and cell states)
static int g_proj_f16 = 1;   // binary16 operands in the training-side input projections (0: TF32)
extern "C" int ocr_debug_proj_f16(int on) { g_proj_f16 = on ? 1 : 0; return OCR_OK; }

extern "C" int ocr_birnn_lstm_train_workspace_bytes(int T, int B, int H, size_t* bytes)
{
    OCR_CHECK_ARG(bytes && T >= 0 && B >= 0 && H >= 1, "ocr_birnn_lstm_train_workspace_bytes: bad argument");
    // gh [2B, 8H] (backward: split-K partials of dh_rec, at most 8 x [2B, H]) + h, c, dh, dc [2B,H] + dgs [2B,4H]
    // + the bfloat16 copy of wh_rows [2H, 4H] for the 16-bit frame-by-frame product (4 H^2 floats)
    size_t fl = (size_t)2 * B * 8 * H + (size_t)2 * B * H * 4 + (size_t)2 * B * 4 * H + (size_t)4 * H * H + 64;
    if (T >= 1 && B >= 1 && lstm_persistent_supported(T, B, H) && lstm_persistent_workspace_floats(B, H) > fl) fl = lstm_persistent_workspace_floats(B, H);
    if (T >= 1 && B >= 1 && lstm_bptt_supported(T, B, H) && lstm_bptt_workspace_floats(B, H) > fl) fl = lstm_bptt_workspace_floats(B, H);
    // binary16 copies of the layer input [T*B, I] and of the input kernels [8H, I] for the input projection (I <= 2H: the
    // recognizer's layers; wider inputs take the TF32 product): (T*B + 8H) * 2H halves, used before anything else touches the workspace
    const size_t f16 = ((size_t)T * B + (size_t)8 * H) * H + 64;
    if (f16 > fl) fl = f16;
    *bytes = sizeof(float) * fl + 256;
    return OCR_OK;
}

extern "C" int ocr_birnn_lstm_train_fwd(const float* x, int T, int B, int I, int H, const int32_t* seq_len, const float* wx, const float* wh,
                                        const float* bias, float* out, float* gates, float* cstate, void* workspace, size_t workspace_bytes,
                                        ocr_stream_t stream)
{
    OCR_CHECK_ARG(T >= 1 && B >= 1 && I >= 4 && (I % 4) == 0 && H >= 4 && (H % 4) == 0, "ocr_birnn_lstm_train_fwd: bad shape T=%d B=%d I=%d H=%d", T, B, I, H);
    OCR_CHECK_ARG(x && seq_len && wx && wh && bias && out && gates && cstate, "ocr_birnn_lstm_train_fwd: NULL argument");
    size_t need = 0;
    ocr_birnn_lstm_train_workspace_bytes(T, B, H, &need);
    if (workspace == nullptr || workspace_bytes < need) { set_error("ocr_birnn_lstm_train_fwd: workspace too small"); return OCR_EWORKSPACE; }
    cudaStream_t st = ST(stream);
    float* ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    float* gh = ws;
    float* h = gh + (size_t)2 * B * 8 * H;
    float* c = h + (size_t)2 * B * H;
    int rc;
    if (g_proj_f16 && I <= 2 * H && (I % 8) == 0) {
        // input projection with binary16 operands (the 10 mantissa bits TF32 keeps; layer inputs are ReLU outputs / tanh-bounded
        // layer outputs, far inside binary16's range): twice the tensor rate, half the operand bytes.  The copies live at the
        // head of the workspace, which nothing else uses before the product has run (same stream).
        unsigned short* x16 = reinterpret_cast<unsigned short*>(ws);
        unsigned short* w16 = x16 + (size_t)T * B * I;
        rc = ocr_float_to_half(x, x16, (long long)T * B * I, stream);
        if (rc != OCR_OK) return rc;
        rc = ocr_float_to_half(wx, w16, (long long)8 * H * I, stream);
        if (rc != OCR_OK) return rc;
        rc = ocr_gemm_f16(x16, I, w16, I, bias, gates, 8 * H, T * B, 8 * H, I, 0, stream);
    } else {
        rc = ocr_gemm_tf32(x, I, wx, I, bias, gates, 8 * H, T * B, 8 * H, I, 0, stream);
    }
    if (rc != OCR_OK) return rc;
    if ((g_birnn_path == 0 || g_birnn_path == 2 || g_birnn_path == 3) && lstm_persistent_supported(T, B, H)) {
        // all T frames of both directions in ONE cooperative launch, W_h resident in shared memory (lstm_persistent.cu)
        OCR_CHECK_CUDA(cudaMemsetAsync(cstate, 0, sizeof(float) * (size_t)T * B * 2 * H, st));
        return lstm_persistent_run(gates, wh, nullptr, seq_len, T, B, H, out, ws, st, gates, cstate);
    }
    OCR_CHECK_CUDA(cudaMemsetAsync(h, 0, sizeof(float) * (size_t)2 * B * H * 2, st));
    OCR_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)T * B * 2 * H, st));
    OCR_CHECK_CUDA(cudaMemsetAsync(cstate, 0, sizeof(float) * (size_t)T * B * 2 * H, st));
    OCR_CHECK_CUDA(cudaMemsetAsync(gh, 0, sizeof(float) * (size_t)2 * B * 8 * H, st));
    GemmPlan p1;
    // both directions in one launch: gh[d] [B, 4H] = h[d] [B, H] * wh[d] [4H, H]^T
    rc = gemm_plan_dirs(&p1, h, H, wh, H, gh, B, 4 * H, H, 2, 1, recurrent_bn(B, 4 * H, 1));
    if (rc != OCR_OK) return rc;
    const int cg = grid_cap((long long)2 * B * H);
    for (int s = 0; s < T; ++s) {
        if (s > 0) { rc = gemm_run(p1, st); if (rc != OCR_OK) return rc; }
        lstm_cell_train_kernel<<<cg, 256, 0, st>>>(gh, gates, seq_len, s, T, B, H, h, c, out, cstate);
        OCR_CHECK_LAUNCH();
    }
    return OCR_OK;
}

// wh_rows [2H, 4H]: the h-part of the TensorFlow kernels (rows = hidden unit, columns = gates i,j,f,o), forward
// direction's H rows then the backward direction's.  gates: activations in, d(pre-activation) out.
static int g_bptt_pdl = 1;   // programmatic dependent launch on the frame-by-frame BPTT chain (0: ordinary launches)
static int g_bptt_bn = 0;    // tile width of the frame-by-frame recurrent product (0: automatic)
static int g_bptt_h16 = 1;   // bfloat16 operands (dG, W_h) in the frame-by-frame recurrent product where 4H % 64 == 0
__global__ void to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = __float2bfloat16_rn(in[i]);
}
extern "C" int ocr_debug_bptt_pdl(int on) {
    g_bptt_pdl = (on & 1) ? 1 : 0;
    g_bptt_h16 = (on & 2) ? 0 : 1;
    g_bptt_bn = on >> 4 << 4;             // tuning: on = 1 + 64 / 128 / 256 overrides the tile width
    return OCR_OK;
}
namespace ocr { int lstm_bptt_set_copies(int on); }
// Tuning aid: the persistent BPTT kernel's row copies at small batches (1, default) / one copy per row (0); same bits either way.
extern "C" int ocr_debug_bptt_copies(int on) { return lstm_bptt_set_copies(on); }

extern "C" int ocr_birnn_lstm_bwd(const float* dout, int T, int B, int H, const int32_t* seq_len, float* gates, const float* cstate,
                                  const float* wh_rows, void* workspace, size_t workspace_bytes, ocr_stream_t stream)
{
    OCR_CHECK_ARG(T >= 1 && B >= 1 && H >= 4 && (H % 4) == 0, "ocr_birnn_lstm_bwd: bad shape T=%d B=%d H=%d", T, B, H);
    OCR_CHECK_ARG(dout && seq_len && gates && cstate && wh_rows, "ocr_birnn_lstm_bwd: NULL argument");
    size_t need = 0;
    ocr_birnn_lstm_train_workspace_bytes(T, B, H, &need);
    if (workspace == nullptr || workspace_bytes < need) { set_error("ocr_birnn_lstm_bwd: workspace too small"); return OCR_EWORKSPACE; }
    cudaStream_t st = ST(stream);
    float* ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    float* dh_rec = ws;                               // [2][splits][B][H], splits <= 8
    float* dh = dh_rec + (size_t)2 * B * 8 * H;
    float* dc = dh + (size_t)2 * B * H;
    float* dgs = dc + (size_t)2 * B * H * 3;          // [2B, 4H] float32 or bfloat16
    __nv_bfloat16* wh16 = reinterpret_cast<__nv_bfloat16*>(dgs + (size_t)2 * B * 4 * H);   // [2H, 4H] bfloat16 copy of wh_rows
    OCR_CHECK_CUDA(cudaMemsetAsync(dh, 0, sizeof(float) * (size_t)2 * B * H * 2, st));
    zero_past_len_kernel<<<grid_cap((long long)T * B * 32), 256, 0, st>>>(gates, seq_len, T, B, 2 * H);
    OCR_CHECK_LAUNCH();
    // all frames in one cooperative launch, W_h slices resident on chip.  Its partial-sum exchange grows with the batch
    // (512 KB per CTA per frame at 128 rows): measured faster than the per-frame launches up to B = 64 (12-16 % of the step)
    if (((g_birnn_path == 0 && B <= 64) || g_birnn_path == 3) && lstm_bptt_supported(T, B, H))
        return lstm_bptt_run(dout, T, B, H, seq_len, gates, cstate, wh_rows, ws, st);
    GemmPlan p1;
    // dh_rec[d][b, n] = sum_g dgs[d*B + b, g] * wh_rows[d*H + n, g]: K = 4H is long and the tile count small, so the
    // contraction is split over K across the SMs; the cell kernel of the next step adds the partials up
    const int bn = g_bptt_bn ? g_bptt_bn : (H > 32 ? 64 : 32);
    int want = 148 / (2 * ((B + 127) / 128) * ((H + bn - 1) / bn));     // one full wave of CTAs, no ragged second wave
    if (want > 8) want = 8;
    if (want < 1) want = 1;
    // 16-bit operands: the product is bound by the operand bytes an SM pulls through L2 (A tile re-read by every N tile, W by
    // every M tile: 48 MB per frame at B = 256 in float32); bfloat16 gate gradients x bfloat16 weights halve them
    const bool h16 = g_bptt_h16 && (H % 16) == 0 && bn <= 64;
    int rc;
    if (h16) {
        to_bf16_kernel<<<grid_cap((long long)8 * H * H), 256, 0, st>>>(wh_rows, wh16, (long long)8 * H * H);
        OCR_CHECK_LAUNCH();
        rc = gemm_plan_dirs_h16(&p1, dgs, 4 * H, wh16, 4 * H, dh_rec, B, H, 4 * H, 2, want, bn);
    } else {
        rc = gemm_plan_dirs(&p1, dgs, 4 * H, wh_rows, 4 * H, dh_rec, B, H, 4 * H, 2, want, bn);
    }
    if (rc != OCR_OK) return rc;
    const int splits = p1.splits;
    const int cg = grid_cap((long long)2 * B * (H / 4));
    // The 2T launches of this loop are one dependent chain of short kernels: each is launched with programmatic stream
    // serialization, so its CTAs are scheduled (and the product's prologue -- barriers, TMEM allocation -- runs) while the
    // predecessor drains; both kernels order themselves with griddepcontrol.wait before their first global access.
    p1.pdl = g_bptt_pdl;
    for (int s = T - 1; s >= 0; --s) {
        if (g_bptt_pdl && s < T - 1) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cg);
            cfg.blockDim = dim3(256);
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            if (h16) OCR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, lstm_cell_bwd_kernel<__nv_bfloat16>, gates, (const float*)cstate, dout, (const float*)dh_rec, seq_len, s, 0, T, B, H, splits, dh, dc, reinterpret_cast<__nv_bfloat16*>(dgs)));
            else OCR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, lstm_cell_bwd_kernel<float>, gates, (const float*)cstate, dout, (const float*)dh_rec, seq_len, s, 0, T, B, H, splits, dh, dc, dgs));
            count_launch();
        } else {
            if (h16) lstm_cell_bwd_kernel<__nv_bfloat16><<<cg, 256, 0, st>>>(gates, cstate, dout, dh_rec, seq_len, s, s == T - 1 ? 1 : 0, T, B, H, splits, dh, dc, reinterpret_cast<__nv_bfloat16*>(dgs));
            else lstm_cell_bwd_kernel<float><<<cg, 256, 0, st>>>(gates, cstate, dout, dh_rec, seq_len, s, s == T - 1 ? 1 : 0, T, B, H, splits, dh, dc, dgs);
            OCR_CHECK_LAUNCH();
        }
        if (s > 0) { rc = gemm_run(p1, st); if (rc != OCR_OK) return rc; }
    }
    return OCR_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// bidirectional GRU layer, training form (model.py:167-199) and its back-propagation through time
static int gru_splits(int B, int H, int K) {
    int want = 148 / (2 * ((B + 127) / 128) * ((H + 63) / 64));
    if (want > 8) want = 8;
    if (want > K / 64) want = K / 64;     // at least two k-steps of work per split
    return want < 1 ? 1 : want;
}

extern "C" int ocr_birnn_gru_train_workspace_bytes(int T, int B, int H, size_t* bytes)
{
    OCR_CHECK_ARG(bytes && T >= 0 && B >= 0 && H >= 1, "ocr_birnn_gru_train_workspace_bytes: bad argument");
    // gh [2B,2H], ch, h, rh, dzc, dhd, dhr, hprev [2B,H] each, dzg [2B,2H], two partial buffers of <= 8 x [2B,H]
    *bytes = sizeof(float) * (size_t)2 * B * H * (2 + 7 + 2 + 16) + 256;
    return OCR_OK;
}

extern "C" int ocr_birnn_gru_train_fwd(const float* x, int T, int B, int I, int H, const int32_t* seq_len, const float* wx, const float* whg,
                                       const float* whc, const float* bias, float* out, float* act, float* rh_all, void* workspace,
                                       size_t workspace_bytes, ocr_stream_t stream)
{
    OCR_CHECK_ARG(T >= 1 && B >= 1 && I >= 4 && (I % 4) == 0 && H >= 4 && (H % 4) == 0, "ocr_birnn_gru_train_fwd: bad shape T=%d B=%d I=%d H=%d", T, B, I, H);
    OCR_CHECK_ARG(x && seq_len && wx && whg && whc && bias && out && act && rh_all, "ocr_birnn_gru_train_fwd: NULL argument");
    size_t need = 0;
    ocr_birnn_gru_train_workspace_bytes(T, B, H, &need);
    if (workspace == nullptr || workspace_bytes < need) { set_error("ocr_birnn_gru_train_fwd: workspace too small"); return OCR_EWORKSPACE; }
    cudaStream_t st = ST(stream);
    float* ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    const size_t n = (size_t)2 * B * H;
    float* gh = ws;            // [2][B][2H]
    float* ch = gh + 2 * n;    // [2][B][H]
    float* h = ch + n;
    float* rh = h + n;
    int rc = ocr_gemm_tf32(x, I, wx, I, bias, act, 6 * H, T * B, 6 * H, I, 0, stream);
    if (rc != OCR_OK) return rc;
    OCR_CHECK_CUDA(cudaMemsetAsync(gh, 0, sizeof(float) * 5 * n, st));
    OCR_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)T * B * 2 * H, st));
    OCR_CHECK_CUDA(cudaMemsetAsync(rh_all, 0, sizeof(float) * (size_t)T * B * 2 * H, st));
    GemmPlan p1, p2;
    rc = gemm_plan_dirs(&p1, h, H, whg, H, gh, B, 2 * H, H, 2, 1, recurrent_bn(B, 2 * H, 1));
    if (rc != OCR_OK) return rc;
    rc = gemm_plan_dirs(&p2, rh, H, whc, H, ch, B, H, H, 2, 1, recurrent_bn(B, H, 1));
    if (rc != OCR_OK) return rc;
    const int cg = grid_cap((long long)n);
    for (int s = 0; s < T; ++s) {
        if (s > 0) { rc = gemm_run(p1, st); if (rc != OCR_OK) return rc; }
        gru_gates_train_kernel<<<cg, 256, 0, st>>>(gh, act, seq_len, s, T, B, H, h, rh, rh_all);
        OCR_CHECK_LAUNCH();
        if (s > 0) { rc = gemm_run(p2, st); if (rc != OCR_OK) return rc; }
        gru_cell_train_kernel<<<cg, 256, 0, st>>>(ch, act, seq_len, s, T, B, H, h, out);
        OCR_CHECK_LAUNCH();
    }
    return OCR_OK;
}

// wg_rows [2H, 2H]: rows I.. of the forward cell's gates kernel (TensorFlow layout [H, 2H]) then the backward cell's;
// wc_rows [2H, H]: the same for the candidate kernels.  act: activations in, gradients of the pre-activations out.
extern "C" int ocr_birnn_gru_bwd(const float* dout, int T, int B, int H, const int32_t* seq_len, float* act, const float* out,
                                 const float* wg_rows, const float* wc_rows, void* workspace, size_t workspace_bytes, ocr_stream_t stream)
{
    OCR_CHECK_ARG(T >= 1 && B >= 1 && H >= 4 && (H % 4) == 0, "ocr_birnn_gru_bwd: bad shape T=%d B=%d H=%d", T, B, H);
    OCR_CHECK_ARG(dout && seq_len && act && out && wg_rows && wc_rows, "ocr_birnn_gru_bwd: NULL argument");
    size_t need = 0;
    ocr_birnn_gru_train_workspace_bytes(T, B, H, &need);
    if (workspace == nullptr || workspace_bytes < need) { set_error("ocr_birnn_gru_bwd: workspace too small"); return OCR_EWORKSPACE; }
    cudaStream_t st = ST(stream);
    float* ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    const size_t n = (size_t)2 * B * H;
    float* dzg = ws;               // [2B, 2H]
    float* dzc = dzg + 2 * n;      // [2B, H]
    float* dhd = dzc + n;
    float* dhr = dhd + n;
    float* hprev = dhr + n;
    float* drh = hprev + n + 2 * n;   // <= 8 partials [2][B][H]
    float* dhg = drh + 8 * n;         // <= 8 partials
    OCR_CHECK_CUDA(cudaMemsetAsync(dhd, 0, sizeof(float) * 2 * n, st));
    zero_past_len_kernel<<<grid_cap((long long)T * B * 32), 256, 0, st>>>(act, seq_len, T, B, 6 * H / 4);
    OCR_CHECK_LAUNCH();
    GemmPlan pa, pb;
    // drh[d][b, n] = sum_k dzc[d*B + b, k] * wc_rows[d*H + n, k];  dhg[d][b, n] = sum_k dzg[d*B + b, k] * wg_rows[d*H + n, k]
    int rc = gemm_plan_dirs(&pa, dzc, H, wc_rows, H, drh, B, H, H, 2, gru_splits(B, H, H), H > 32 ? 64 : 32);
    if (rc != OCR_OK) return rc;
    rc = gemm_plan_dirs(&pb, dzg, 2 * H, wg_rows, 2 * H, dhg, B, H, 2 * H, 2, gru_splits(B, H, 2 * H), H > 32 ? 64 : 32);
    if (rc != OCR_OK) return rc;
    const int cg = grid_cap((long long)n);
    for (int s = T - 1; s >= 0; --s) {
        gru_bwd1_kernel<<<cg, 256, 0, st>>>(act, out, dout, seq_len, s, s == T - 1 ? 1 : 0, T, B, H, dhd, dhr, dhg, pb.splits, dzc, dzg, hprev);
        OCR_CHECK_LAUNCH();
        rc = gemm_run(pa, st);
        if (rc != OCR_OK) return rc;
        gru_bwd2_kernel<<<cg, 256, 0, st>>>(act, drh, pa.splits, seq_len, s, T, B, H, hprev, dhr, dzg);
        OCR_CHECK_LAUNCH();
        if (s > 0) { rc = gemm_run(pb, st); if (rc != OCR_OK) return rc; }
    }
    return OCR_OK;
}
