// CTC greedy decoder and edit distance for sm_100a.
//
// Greedy: replaces tf.nn.ctc_greedy_decoder(merge_repeated=True) called by the reference's
// validate._get_output (/root/reference/src/weinman/validate.py:81-92).  Semantics (SURVEY.md
// App. A.5): per frame first-maximum arg-max (ties -> lowest class index; they are real here
// because the logits come out of a ReLU, model.py:206,216), neg_sum_logits accumulated in frame
// order, emit unless blank or equal to the previous frame's class.
// One CTA (4 warps) per sequence: HBM-bound read-once scan of T*C*4 bytes per sequence; sixteen rows are kept
// in flight per CTA to cover DRAM latency.
//
// Edit distance: replaces tf.edit_distance(normalize=False) at src/weinman/test.py:90.  One warp
// per pair, anti-diagonal Levenshtein wavefront in shared memory.
#include "common.cuh"

namespace ocr {

constexpr int kGreedyWarps = 4;

// lexicographic (value desc, index asc) arg-max combine
__device__ __forceinline__ void argmax_combine(float& v, int& i, float ov, int oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

// One CTA (4 warps) per sequence.  Phase 1: the warps take the frames round-robin, four rows in flight each, and
// leave the per-frame arg-max (class, value) in shared memory.  Phase 2: warp 0 collapses repeats / drops blanks with
// ballot prefix counts (order-preserving compaction) and lane 0 adds the maxima in frame order (bit-exact float sum).
__global__ void __launch_bounds__(kGreedyWarps * 32)
ctc_greedy_kernel(const float* __restrict__ logits, int T, int B, int C, const int32_t* __restrict__ seq_len,
                  int merge_repeated, int64_t* __restrict__ decoded, int32_t* __restrict__ decoded_len,
                  float* __restrict__ neg_sum_logits)
{
    extern __shared__ unsigned char greedy_smem[];
    int* s_idx = reinterpret_cast<int*>(greedy_smem);          // [T]
    float* s_val = reinterpret_cast<float*>(s_idx + T);        // [T]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x;
    const int Tb = min(max(seq_len[b], 0), T);
    const int blank = C - 1;
    const size_t rstride = (size_t)B * C;
    const float* xb = logits + (size_t)b * C;
    int64_t* out = decoded + (size_t)b * T;
    constexpr int R = 4;  // rows in flight per warp
    for (int t0 = warp * R; t0 < Tb; t0 += kGreedyWarps * R) {
        float bv[R];
        int bi[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            bv[r] = -INFINITY;
            bi[r] = 0x7fffffff;
            const int t = t0 + r;
            if (t < Tb) {
                const float* row = xb + t * rstride;
                for (int k = lane; k < C; k += 32) {
                    float v = ld_stream(row + k);
                    if (v > bv[r] || bi[r] == 0x7fffffff) { bv[r] = v; bi[r] = k; }  // strict '>' keeps the first max
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                float ov = __shfl_xor_sync(kFullMask, bv[r], o);
                int oi = __shfl_xor_sync(kFullMask, bi[r], o);
                argmax_combine(bv[r], bi[r], ov, oi);
            }
            if (lane == 0 && t0 + r < Tb) { s_idx[t0 + r] = bi[r]; s_val[t0 + r] = bv[r]; }
        }
    }
    __syncthreads();
    if (warp != 0) return;
    int n = 0;
    for (int t0 = 0; t0 < Tb; t0 += 32) {
        const int t = t0 + lane;
        bool keep = false;
        int c = 0;
        if (t < Tb) {
            c = s_idx[t];
            const int prev = t > 0 ? s_idx[t - 1] : -1;
            keep = c != blank && !(merge_repeated && c == prev);
        }
        const unsigned m = __ballot_sync(kFullMask, keep);
        if (keep) out[n + __popc(m & ((1u << lane) - 1))] = c;
        n += __popc(m);
    }
    for (int t = n + lane; t < T; t += 32) out[t] = -1;
    if (lane == 0) {
        float acc = 0.0f;
        for (int t = 0; t < Tb; ++t) acc += -s_val[t];
        decoded_len[b] = n;
        neg_sum_logits[b] = acc;
    }
}

// One warp per (hyp, truth) pair.  Rows of the DP table live in shared memory; each lane owns
// columns j = lane, lane+32, ... and rows are processed sequentially (n, m are tens of symbols).
constexpr int kEditWarps = 4;
__global__ void __launch_bounds__(kEditWarps * 32)
edit_distance_kernel(const int64_t* __restrict__ hyp, int hyp_stride, const int32_t* __restrict__ hyp_len,
                     const int32_t* __restrict__ truth, const int32_t* __restrict__ truth_off, int B, int max_m,
                     float* __restrict__ dist)
{
    extern __shared__ int s_rows[];  // [kEditWarps][2][max_m+1]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.x * kEditWarps + w;
    if (b >= B) return;
    int* r0 = s_rows + (size_t)w * 2 * (max_m + 1);
    int* r1 = r0 + (max_m + 1);
    const int64_t* h = hyp + (size_t)b * hyp_stride;
    int n = 0;
    if (hyp_len) n = hyp_len[b];
    else {
        // -1 padded row: length = number of leading non-negative entries
        int cnt = 0;
        for (int i = lane; i < hyp_stride; i += 32) cnt += (h[i] >= 0);
        n = warp_sum_int(cnt);
    }
    const int32_t* g = truth + truth_off[b];
    const int m = truth_off[b + 1] - truth_off[b];
    for (int j = lane; j <= m; j += 32) r0[j] = j;
    __syncwarp();
    for (int i = 1; i <= n; ++i) {
        const int64_t hi = h[i - 1];
        // substitution / deletion terms are independent of the new row; the insertion term is a
        // running minimum (prefix-min of v[j]-j, plus j), resolved with a warp scan per 32 columns
        int carry = i;  // cur[0] = i
        if (lane == 0) r1[0] = i;
        for (int j0 = 1; j0 <= m; j0 += 32) {
            const int j = j0 + lane;
            int v = 0x3fffffff;
            if (j <= m) {
                int sub = r0[j - 1] + (hi != (int64_t)g[j - 1]);
                int del = r0[j] + 1;
                v = min(sub, del);
            }
            // cur[j] = min(v[j], cur[j-1]+1)  ==>  cur[j] = j + min(prefixmin_{k<=j}(v[k]-k), carry - (j0-1))
            int key = v - j;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int other = __shfl_up_sync(kFullMask, key, o);
                if (lane >= o) key = min(key, other);
            }
            int cur = j + min(key, carry - (j0 - 1));
            if (j <= m) r1[j] = cur;
            carry = __shfl_sync(kFullMask, cur, 31);
        }
        __syncwarp();
        int* tmp = r0; r0 = r1; r1 = tmp;
    }
    if (lane == 0) dist[b] = (float)r0[m];
}

}  // namespace ocr

using namespace ocr;

extern "C" int ocr_ctc_greedy_decode(const float* logits, int T, int B, int C, const int32_t* seq_len,
                                     int merge_repeated, int64_t* decoded, int32_t* decoded_len,
                                     float* neg_sum_logits, ocr_stream_t stream)
{
    OCR_CHECK_ARG(T >= 1 && B >= 0 && C >= 2, "ocr_ctc_greedy_decode: bad shape T=%d B=%d C=%d", T, B, C);
    if (B == 0) return OCR_OK;
    OCR_CHECK_ARG(logits && seq_len && decoded && decoded_len && neg_sum_logits, "ocr_ctc_greedy_decode: NULL argument");
    OCR_CHECK_ARG((size_t)T * 8 <= 48 * 1024, "ocr_ctc_greedy_decode: T=%d too large", T);
    ctc_greedy_kernel<<<B, kGreedyWarps * 32, (size_t)T * 8, static_cast<cudaStream_t>(stream)>>>(
        logits, T, B, C, seq_len, merge_repeated, decoded, decoded_len, neg_sum_logits);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

extern "C" int ocr_edit_distance(const int64_t* hyp, int hyp_stride, const int32_t* hyp_len, const int32_t* truth,
                                 const int32_t* truth_offsets, int B, int max_truth_len, float* dist,
                                 ocr_stream_t stream)
{
    OCR_CHECK_ARG(B >= 0 && hyp_stride >= 0 && max_truth_len >= 0, "ocr_edit_distance: bad shape");
    if (B == 0) return OCR_OK;
    OCR_CHECK_ARG(hyp && truth && truth_offsets && dist, "ocr_edit_distance: NULL argument");
    const size_t smem = (size_t)kEditWarps * 2 * (max_truth_len + 1) * sizeof(int);
    OCR_CHECK_ARG(smem <= 48 * 1024, "ocr_edit_distance: max_truth_len=%d too large", max_truth_len);
    const int grid = (B + kEditWarps - 1) / kEditWarps;
    edit_distance_kernel<<<grid, kEditWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
        hyp, hyp_stride, hyp_len, truth, truth_offsets, B, max_truth_len, dist);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}
