// The recognizer's layers around the tensor-core GEMM (sm_100a), INFER mode.
//
//   conv1 (3x3 'valid', one input channel) with the image preprocessing fused in       model.py:47,84-109; validate.py:56-68
//   im2col for the 3x3 'same' convolutions, with the preceding max-pool fused into the gather   model.py:111-116,134-146
//   pool8 (3x1) + squeeze + time-major transpose                                       model.py:145-147,212
//   the frame loop of the bidirectional LSTM / GRU layers (tf.nn.bidirectional_dynamic_rnn with per-example
//   sequence lengths)                                                                   model.py:167-199, model_bu.py:167-199
//
// The convolutions themselves, the RNN input projections and the logits layer are ocr_gemm_tf32 calls
// (tcgen05); everything here is the memory-bound glue, written as coalesced float4 kernels.
// Batch-norm in INFER mode is an affine map per channel and is folded into the filters and biases by the host
// (w' = w * gamma / sqrt(var + eps), b' = (b - mean) * gamma / sqrt(var + eps) + beta), so conv -> BN -> ReLU is one
// GEMM epilogue.
#include "gemm_tf32.cuh"

namespace ocr {

// ---------------------------------------------------------------------------------------------------------------
// tf.image.convert_image_dtype(uint8 -> float32) multiplies by the float32 constant 1/255 (scale = 1. / dtype.max)
__device__ constexpr float kInv255 = (float)(1.0 / 255.0);

// mjsynth._preprocess_image + the zero padding of the batcher (mjsynth.py:185-194, 56, 69): u8 rows -> float32, first row
// duplicated on top, 0.0 (not -0.5) right of each crop's own width.  One thread = one output pixel.
__global__ void __launch_bounds__(256)
preprocess_train_kernel(const unsigned char* __restrict__ in, int B, int Hin, int W, const int32_t* __restrict__ widths, float* __restrict__ out)
{
    const long long total = (long long)B * (Hin + 1) * W;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % W);
        long long p = idx / W;
        const int r = (int)(p % (Hin + 1));
        const int b = (int)(p / (Hin + 1));
        const int src = r > 0 ? r - 1 : 0;
        float v = 0.0f;
        if (x < __ldg(widths + b)) v = __fsub_rn(__fmul_rn((float)__ldg(in + ((size_t)b * Hin + src) * W + x), kInv255), 0.5f);   // two roundings, as TF's multiply then subtract (no FMA contraction)
        out[idx] = v;
    }
}

// conv1: out[b,y,x,:] = relu(bias + sum_{i,j} w[i,j,:] * in[b,y+i,x+j]),   in = u8 * (1/255) - 0.5 or float
// one thread = one output pixel x 4 channels.  kFixed (Co/4 divides 256 and the pixel count fits 32 bits): the grid stride is a
// multiple of Co/4, so a thread keeps ITS four channels -- bias and the nine filter taps live in registers -- and the pixel
// coordinates come from 32-bit divisions (the general form decodes a 64-bit index with five 64-bit divisions and reloads ten
// float4 per pixel: 165 us for cfg3's 250 MB output).  Same operations in the same order: same bits.
template <bool kU8, bool kFixed>
__global__ void __launch_bounds__(256)
conv1_kernel(const void* __restrict__ in_, int B, int H, int W, const float* __restrict__ w /*[3,3,1,Co]*/,
             const float* __restrict__ bias, int Co, float* __restrict__ out)
{
    const int Ho = H - 2, Wo = W - 2, c4n = Co >> 2;
    const long long total = (long long)B * Ho * Wo * c4n;
    auto tap = [&](size_t o) -> float {
        if (kU8) return __fsub_rn(__fmul_rn((float)__ldg(reinterpret_cast<const unsigned char*>(in_) + o), kInv255), 0.5f);
        return __ldg(reinterpret_cast<const float*>(in_) + o);
    };
    if (kFixed) {
        const unsigned first = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
        const unsigned c4 = first % (unsigned)c4n;
        const float4 bv = __ldg(reinterpret_cast<const float4*>(bias) + c4);
        float4 wv[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) wv[t] = __ldg(reinterpret_cast<const float4*>(w + t * Co) + c4);
        for (unsigned idx = first; idx < (unsigned)total; idx += stride) {
            unsigned p = idx / (unsigned)c4n;
            const unsigned x = p % (unsigned)Wo; p /= (unsigned)Wo;
            const unsigned y = p % (unsigned)Ho;
            const unsigned b = p / (unsigned)Ho;
            float4 acc = bv;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const float v = tap(((size_t)b * H + (y + i)) * W + (x + j));
                    const float4 ww = wv[i * 3 + j];
                    acc.x = fmaf(v, ww.x, acc.x); acc.y = fmaf(v, ww.y, acc.y); acc.z = fmaf(v, ww.z, acc.z); acc.w = fmaf(v, ww.w, acc.w);
                }
            acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
            reinterpret_cast<float4*>(out)[idx] = acc;
        }
        return;
    }
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % c4n);
        long long p = idx / c4n;
        const int x = (int)(p % Wo); p /= Wo;
        const int y = (int)(p % Ho);
        const int b = (int)(p / Ho);
        float4 acc = __ldg(reinterpret_cast<const float4*>(bias) + c4);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float v = tap(((size_t)b * H + (y + i)) * W + (x + j));
                const float4 ww = __ldg(reinterpret_cast<const float4*>(w + (i * 3 + j) * Co) + c4);
                acc.x = fmaf(v, ww.x, acc.x); acc.y = fmaf(v, ww.y, acc.y); acc.z = fmaf(v, ww.z, acc.z); acc.w = fmaf(v, ww.w, acc.w);
            }
        acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
        reinterpret_cast<float4*>(out)[idx] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// im2col of pool(in): in [B,H,W,C] NHWC; pooled dims Hp = (H-ph)/sh+1, Wp = (W-pw)/sw+1; 3x3 'same' patches:
// out[(b,y,x), (i,j,c)] = pooled[b, y+i-1, x+j-1, c] (0 outside).  One thread = one float4 of channels.
__global__ void __launch_bounds__(256)
im2col3x3_kernel(const float* __restrict__ in, int B, int H, int W, int C, int ph, int pw, int sh, int sw, int Hp, int Wp,
                 float* __restrict__ out)
{
    const int c4n = C >> 2;
    const long long total = (long long)B * Hp * Wp * 9 * c4n;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % c4n);
        long long p = idx / c4n;
        const int tap = (int)(p % 9); p /= 9;
        const int x = (int)(p % Wp); p /= Wp;
        const int y = (int)(p % Hp);
        const int b = (int)(p / Hp);
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (yy >= 0 && yy < Hp && xx >= 0 && xx < Wp) {
            v = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            for (int dy = 0; dy < ph; ++dy)
                for (int dx = 0; dx < pw; ++dx) {
                    const float4 u = __ldg(reinterpret_cast<const float4*>(in + (((size_t)b * H + (yy * sh + dy)) * W + (xx * sw + dx)) * C) + c4);
                    v.x = fmaxf(v.x, u.x); v.y = fmaxf(v.y, u.y); v.z = fmaxf(v.z, u.z); v.w = fmaxf(v.w, u.w);
                }
        }
        reinterpret_cast<float4*>(out)[idx] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// pool8 + squeeze + transpose: in [B,H,W,C] -> out [W,B,C], max over the H rows (H = 3 for 32-pixel-high crops)
__global__ void __launch_bounds__(256)
rows_max_to_seq_kernel(const float* __restrict__ in, int B, int H, int W, int C, float* __restrict__ out)
{
    const int c4n = C >> 2;
    const long long total = (long long)W * B * c4n;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % c4n);
        long long p = idx / c4n;
        const int b = (int)(p % B);
        const int x = (int)(p / B);
        float4 v = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        for (int y = 0; y < H; ++y) {
            const float4 u = __ldg(reinterpret_cast<const float4*>(in + (((size_t)b * H + y) * W + x) * C) + c4);
            v.x = fmaxf(v.x, u.x); v.y = fmaxf(v.y, u.y); v.z = fmaxf(v.z, u.z); v.w = fmaxf(v.w, u.w);
        }
        reinterpret_cast<float4*>(out)[idx] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Recurrent cells.  State rows: [0,B) forward direction, [B,2B) backward direction.  The h * W_h^T product of
// BOTH directions is one stacked GEMM ([2B,H] x [2*G*H,H]^T, the off-diagonal blocks are ignored), so the cell
// kernels read row r, column block dir.  Frame of step s: forward t = s, backward t = len-1-s (reverse_sequence);
// examples with s >= len keep their state and emit nothing (the output buffer is pre-zeroed).
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// LSTMCell, gate order i, j, f, o, forget_bias 1.0 (TF): gh [2B, 8H], xp [T*B, 8H] (bias already added)
__global__ void __launch_bounds__(256)
lstm_cell_kernel(const float* __restrict__ gh, const float* __restrict__ xp, const int32_t* __restrict__ seq_len, int s, int T, int B, int H,
                 float* __restrict__ h, float* __restrict__ c, float* __restrict__ out /*[T,B,2H]*/)
{
    const int total = 2 * B * H;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int j = idx % H;
        const int r = idx / H;  // state row
        const int dir = r >= B, b = dir ? r - B : r;
        const int len = min(seq_len[b], T);
        if (s >= len) continue;
        const int t = dir ? len - 1 - s : s;
        const float* g = gh + (size_t)r * 8 * H + dir * 4 * H;
        const float* x = xp + ((size_t)t * B + b) * 8 * H + dir * 4 * H;
        const float zi = g[j] + x[j], zj = g[H + j] + x[H + j], zf = g[2 * H + j] + x[2 * H + j], zo = g[3 * H + j] + x[3 * H + j];
        const float cn = sigmoidf_(zf + 1.0f) * c[idx] + sigmoidf_(zi) * tanhf(zj);
        const float hn = sigmoidf_(zo) * tanhf(cn);
        c[idx] = cn;
        h[idx] = hn;
        out[((size_t)t * B + b) * 2 * H + dir * H + j] = hn;
    }
}

// GRUCell part 1: r,u = sigmoid(gates); writes r*h (the A operand of the candidate GEMM) and u.
// gh [2B, 4H] (per direction 2H: r | u), xp [T*B, 6H] (per direction 3H: r | u | candidate)
__global__ void __launch_bounds__(256)
gru_gates_kernel(const float* __restrict__ gh, const float* __restrict__ xp, const int32_t* __restrict__ seq_len, int s, int T, int B, int H,
                 const float* __restrict__ h, float* __restrict__ rh, float* __restrict__ u)
{
    const int total = 2 * B * H;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int j = idx % H;
        const int r = idx / H;
        const int dir = r >= B, b = dir ? r - B : r;
        const int len = min(seq_len[b], T);
        if (s >= len) { rh[idx] = 0.0f; continue; }
        const int t = dir ? len - 1 - s : s;
        const float* g = gh + (size_t)r * 4 * H + dir * 2 * H;
        const float* x = xp + ((size_t)t * B + b) * 6 * H + dir * 3 * H;
        const float rr = sigmoidf_(g[j] + x[j]);
        u[idx] = sigmoidf_(g[H + j] + x[H + j]);
        rh[idx] = rr * h[idx];
    }
}
// GRUCell part 2: c = tanh(candidate); h' = u*h + (1-u)*c.   ch [2B, 2H] (per direction H)
__global__ void __launch_bounds__(256)
gru_cell_kernel(const float* __restrict__ ch, const float* __restrict__ xp, const float* __restrict__ u, const int32_t* __restrict__ seq_len,
                int s, int T, int B, int H, float* __restrict__ h, float* __restrict__ out)
{
    const int total = 2 * B * H;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int j = idx % H;
        const int r = idx / H;
        const int dir = r >= B, b = dir ? r - B : r;
        const int len = min(seq_len[b], T);
        if (s >= len) continue;
        const int t = dir ? len - 1 - s : s;
        const float cand = tanhf(ch[(size_t)r * 2 * H + dir * H + j] + xp[((size_t)t * B + b) * 6 * H + dir * 3 * H + 2 * H + j]);
        const float uu = u[idx];
        const float hn = uu * h[idx] + (1.0f - uu) * cand;
        h[idx] = hn;
        out[((size_t)t * B + b) * 2 * H + dir * H + j] = hn;
    }
}

// lstm_persistent.cu
bool lstm_persistent_supported(int T, int B, int H);
size_t lstm_persistent_workspace_floats(int B, int H);
int lstm_permute_wh(const float* wh, int H, float* whp, cudaStream_t st);
int lstm_persistent_run(const float* xp, const float* wh, const float* wh_perm, const int32_t* seq_len, int T, int B, int H,
                        float* out, float* ws, cudaStream_t st, float* gates_out = nullptr, float* cs_out = nullptr);

// gru_persistent.cu
bool gru_persistent_supported(int T, int B, int H);
size_t gru_persistent_workspace_floats(int B, int H);
int gru_persistent_run(const float* xp, const float* whg, const float* whc, const int32_t* seq_len, int T, int B, int H,
                       float* out, float* ws, cudaStream_t st);

int lstm_set_timeline(long long* buf);
int lstm_bptt_set_timeline(long long* buf);
int lstm_set_operands(int f16);
int gru_set_operands(int f16);

static inline int grid_for(long long total, int threads = 256) {
    long long g = (total + threads - 1) / threads;
    const long long cap = 148LL * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace ocr

using namespace ocr;

extern "C" int ocr_conv1_3x3_valid(const void* in, int in_is_u8, int B, int H, int W, const float* w, const float* bias, int Cout,
                                   float* out, ocr_stream_t stream)
{
    OCR_CHECK_ARG(B >= 0 && H >= 3 && W >= 3 && Cout >= 4 && (Cout % 4) == 0, "ocr_conv1_3x3_valid: bad shape B=%d H=%d W=%d Cout=%d", B, H, W, Cout);
    if (B == 0) return OCR_OK;
    OCR_CHECK_ARG(in && w && bias && out, "ocr_conv1_3x3_valid: NULL argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long total = (long long)B * (H - 2) * (W - 2) * (Cout / 4);
    const bool fixed = (256 % (Cout / 4)) == 0 && total < 0x7fffffffLL;
    if (in_is_u8) {
        if (fixed) conv1_kernel<true, true><<<grid_for(total), 256, 0, st>>>(in, B, H, W, w, bias, Cout, out);
        else conv1_kernel<true, false><<<grid_for(total), 256, 0, st>>>(in, B, H, W, w, bias, Cout, out);
    } else {
        if (fixed) conv1_kernel<false, true><<<grid_for(total), 256, 0, st>>>(in, B, H, W, w, bias, Cout, out);
        else conv1_kernel<false, false><<<grid_for(total), 256, 0, st>>>(in, B, H, W, w, bias, Cout, out);
    }
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

extern "C" int ocr_preprocess_train(const unsigned char* in, int B, int Hin, int W, const int32_t* widths, float* out, ocr_stream_t stream)
{
    OCR_CHECK_ARG(B >= 0 && Hin >= 1 && W >= 1, "ocr_preprocess_train: bad shape B=%d Hin=%d W=%d", B, Hin, W);
    if (B == 0) return OCR_OK;
    OCR_CHECK_ARG(in && widths && out, "ocr_preprocess_train: NULL argument");
    preprocess_train_kernel<<<grid_for((long long)B * (Hin + 1) * W), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, B, Hin, W, widths, out);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

extern "C" int ocr_im2col3x3_same(const float* in, int B, int H, int W, int C, int pool_h, int pool_w, int stride_h, int stride_w,
                                  float* out, ocr_stream_t stream)
{
    OCR_CHECK_ARG(B >= 0 && C >= 4 && (C % 4) == 0 && pool_h >= 1 && pool_w >= 1 && stride_h >= 1 && stride_w >= 1 && H >= pool_h && W >= pool_w,
                  "ocr_im2col3x3_same: bad shape B=%d H=%d W=%d C=%d pool %dx%d stride %dx%d", B, H, W, C, pool_h, pool_w, stride_h, stride_w);
    if (B == 0) return OCR_OK;
    OCR_CHECK_ARG(in && out, "ocr_im2col3x3_same: NULL argument");
    const int Hp = (H - pool_h) / stride_h + 1, Wp = (W - pool_w) / stride_w + 1;
    const long long total = (long long)B * Hp * Wp * 9 * (C / 4);
    im2col3x3_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, B, H, W, C, pool_h, pool_w, stride_h, stride_w, Hp, Wp, out);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

extern "C" int ocr_rows_max_to_seq(const float* in, int B, int H, int W, int C, float* out, ocr_stream_t stream)
{
    OCR_CHECK_ARG(B >= 0 && H >= 1 && W >= 1 && C >= 4 && (C % 4) == 0, "ocr_rows_max_to_seq: bad shape B=%d H=%d W=%d C=%d", B, H, W, C);
    if (B == 0) return OCR_OK;
    OCR_CHECK_ARG(in && out, "ocr_rows_max_to_seq: NULL argument");
    rows_max_to_seq_kernel<<<grid_for((long long)W * B * (C / 4)), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, B, H, W, C, out);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

// 0 = automatic (persistent LSTM kernel when the shape allows), 1 = frame-by-frame launches only
namespace ocr { int g_birnn_path = 0; }
extern "C" int ocr_birnn_set_path(int path) {
    OCR_CHECK_ARG(path >= 0 && path <= 3, "ocr_birnn_set_path: path=%d outside [0,3]", path);
    g_birnn_path = path;
    return OCR_OK;
}

// Tuning aid: per-frame clock64() stamps of CTA 0 of the persistent LSTM kernel (8 int64 per frame; NULL = off).
extern "C" int ocr_debug_lstm_timeline(long long* device_buffer) {
    const int rc = lstm_set_timeline(device_buffer);
    return rc != OCR_OK ? rc : lstm_bptt_set_timeline(device_buffer);     // the persistent BPTT kernel stamps the same buffer
}

// Operand precision of the persistent recurrence kernels: 1 (default) = IEEE binary16 h and W_h where H % 64 == 0 (K = 16 per
// tcgen05.mma: half the instructions and half the h bytes per frame; same 10 mantissa bits as a TF32 operand), 0 = TF32.
// Takes effect for weights prepared (ocr_lstm_prepare_wh) and layers run after the call.
extern "C" int ocr_debug_lstm_operands(int f16) {
    const int rc = lstm_set_operands(f16);
    return rc != OCR_OK ? rc : gru_set_operands(f16);
}

extern "C" int ocr_birnn_workspace_bytes(int cell, int T, int B, int H, size_t* bytes)
{
    OCR_CHECK_ARG(bytes != nullptr && (cell == 0 || cell == 1) && T >= 0 && B >= 0 && H >= 1, "ocr_birnn_workspace_bytes: bad argument");
    const size_t G = (cell == 0) ? 4 : 3;
    // xp [T*B, 2*G*H] + gh [2B, 2*(cell?2:4)*H] + h, c/u, rh [2B,H] each + ch [2B,2H]
    size_t step = (size_t)2 * B * 8 * H + (size_t)2 * B * H * 3 + (size_t)2 * B * 2 * H;
    if (cell == 0 && lstm_persistent_supported(T, B, H)) step = step > lstm_persistent_workspace_floats(B, H) ? step : lstm_persistent_workspace_floats(B, H);
    if (cell == 1 && gru_persistent_supported(T, B, H)) step = step > gru_persistent_workspace_floats(B, H) ? step : gru_persistent_workspace_floats(B, H);
    *bytes = sizeof(float) * ((size_t)T * B * 2 * G * H + step) + 256;
    return OCR_OK;
}

// One bidirectional recurrent layer (time-major).
//   cell 0 = LSTM: wx [8H, I] (rows: fw i,j,f,o | bw i,j,f,o), wh [8H, H], bias [8H]
//   cell 1 = GRU : wx [6H, I] (rows per direction: r,u,cand), whg [4H, H] (r,u per direction), whc [2H, H], bias [6H]
extern "C" int ocr_lstm_prepare_wh(const float* wh, int H, float* wh_perm, ocr_stream_t stream)
{
    OCR_CHECK_ARG(wh && wh_perm && H >= 16 && (H % 16) == 0, "ocr_lstm_prepare_wh: bad argument (H=%d)", H);
    return lstm_permute_wh(wh, H, wh_perm, static_cast<cudaStream_t>(stream));
}

extern "C" int ocr_birnn_layer(int cell, const float* x, int T, int B, int I, int H, const int32_t* seq_len, const float* wx,
                               const float* wh, const float* wh2, const float* bias, float* out, void* workspace,
                               size_t workspace_bytes, ocr_stream_t stream)
{
    OCR_CHECK_ARG((cell == 0 || cell == 1) && T >= 1 && B >= 1 && I >= 4 && (I % 4) == 0 && H >= 4 && (H % 4) == 0,
                  "ocr_birnn_layer: bad shape cell=%d T=%d B=%d I=%d H=%d", cell, T, B, I, H);
    OCR_CHECK_ARG(x && seq_len && wx && wh && bias && out && (cell == 0 || wh2), "ocr_birnn_layer: NULL argument");
    size_t need = 0;
    ocr_birnn_workspace_bytes(cell, T, B, H, &need);
    if (workspace == nullptr || workspace_bytes < need) {
        set_error("ocr_birnn_layer: workspace too small (%zu < %zu)", workspace_bytes, need);
        return OCR_EWORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int G = (cell == 0) ? 4 : 3;
    float* ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    float* xp = ws;                                   // [T*B, 2*G*H]
    float* gh = xp + (size_t)T * B * 2 * G * H;       // [2B, 8H] (LSTM) / [2B, 4H] (GRU)
    float* h = gh + (size_t)2 * B * 8 * H;            // [2B, H]
    float* c = h + (size_t)2 * B * H;                 // LSTM c / GRU u
    float* rh = c + (size_t)2 * B * H;                // GRU r*h
    float* ch = rh + (size_t)2 * B * H;               // GRU candidate pre-activation [2B, 2H]
    // input projection of every frame and both directions, bias folded in
    int rc = ocr_gemm_tf32(x, I, wx, I, bias, xp, 2 * G * H, T * B, 2 * G * H, I, 0, stream);
    if (rc != OCR_OK) return rc;
    if (cell == 0 && g_birnn_path == 0 && lstm_persistent_supported(T, B, H))
        return lstm_persistent_run(xp, wh, cell == 0 ? wh2 : nullptr, seq_len, T, B, H, out, gh, st);   // one launch for all T frames
    if (cell == 1 && g_birnn_path == 0 && gru_persistent_supported(T, B, H))
        return gru_persistent_run(xp, wh, wh2, seq_len, T, B, H, out, gh, st);                          // likewise: two phases per frame
    OCR_CHECK_CUDA(cudaMemsetAsync(h, 0, sizeof(float) * (size_t)2 * B * H * 3, st));
    OCR_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)T * B * 2 * H, st));
    const int cg = grid_for((long long)2 * B * H);
    GemmPlan p1, p2;
    if (cell == 0) {
        rc = gemm_plan(&p1, h, H, wh, H, nullptr, gh, 8 * H, 2 * B, 8 * H, H, 0);
        if (rc != OCR_OK) return rc;
        for (int s = 0; s < T; ++s) {
            if (s > 0) { rc = gemm_run(p1, st); if (rc != OCR_OK) return rc; }
            else OCR_CHECK_CUDA(cudaMemsetAsync(gh, 0, sizeof(float) * (size_t)2 * B * 8 * H, st));  // h_0 = 0
            lstm_cell_kernel<<<cg, 256, 0, st>>>(gh, xp, seq_len, s, T, B, H, h, c, out);
            OCR_CHECK_LAUNCH();
        }
    } else {
        rc = gemm_plan(&p1, h, H, wh, H, nullptr, gh, 4 * H, 2 * B, 4 * H, H, 0);
        if (rc != OCR_OK) return rc;
        rc = gemm_plan(&p2, rh, H, wh2, H, nullptr, ch, 2 * H, 2 * B, 2 * H, H, 0);
        if (rc != OCR_OK) return rc;
        for (int s = 0; s < T; ++s) {
            if (s > 0) { rc = gemm_run(p1, st); if (rc != OCR_OK) return rc; }
            else OCR_CHECK_CUDA(cudaMemsetAsync(gh, 0, sizeof(float) * (size_t)2 * B * 4 * H, st));
            gru_gates_kernel<<<cg, 256, 0, st>>>(gh, xp, seq_len, s, T, B, H, h, rh, c);
            OCR_CHECK_LAUNCH();
            if (s > 0) { rc = gemm_run(p2, st); if (rc != OCR_OK) return rc; }
            else OCR_CHECK_CUDA(cudaMemsetAsync(ch, 0, sizeof(float) * (size_t)2 * B * 2 * H, st));
            gru_cell_kernel<<<cg, 256, 0, st>>>(ch, xp, c, seq_len, s, T, B, H, h, out);
            OCR_CHECK_LAUNCH();
        }
    }
    return OCR_OK;
}
