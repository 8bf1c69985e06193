// CTC loss + gradient for sm_100a.  Replaces tf.nn.ctc_loss as called by the reference's
// ctc_loss_layer (/root/reference/src/weinman/model.py:224-229); semantics follow upstream
// TensorFlow's CTCLossCalculator: blank = C-1, log-domain alpha/beta with
// beta excluding y_t, log p(z|x) = LSE_u(alpha(u,0)+beta(u,0)), gradient w.r.t. the logits
// = softmax - exp(LSE_{u:l'_u=k}(alpha+beta) - log p), zero rows past seq_len, "no valid path"
// -> loss=+inf, grad=softmax.
//
// Work decomposition (HBM/latency-bound, no tensor cores): one CTA per sequence, 4 warps.
//   phase 1  all warps: each warp streams whole logit rows [C] (coalesced, read once from HBM),
//            stages the row in shared memory, reduces max / sum-exp with shuffles and keeps only
//            what the lattice needs: the row's log-sum-exp and log y at the L label classes + blank.
//   phase 2  warp 0 runs the alpha chain forward in time while warp 1 runs the beta chain backward
//            (the two 1-D dependency chains are independent, so they overlap); lattice in smem.
//   phase 3  all warps: per frame, occupancy exp(alpha+beta-logp) scattered per class through a
//            per-class linked list (deterministic order, no atomics), then grad = y - occ written
//            once, coalesced.  Logits are re-read here; they are L2-resident (the CTA read them
//            microseconds ago), so DRAM sees each logit once and each grad element once.
// Algorithmic HBM bytes per sequence: 2*T*C*4 (+ labels).
#include "common.cuh"

namespace ocr {

constexpr int kCtcThreads = 128;
constexpr int kCtcWarps = kCtcThreads / 32;

struct CtcSmemLayout {
    // byte offsets into dynamic shared memory
    int lab, next, head, skip_a, rowlse, stage, lpl, alpha, beta, total;
    int Lp1, U;
};

__host__ __device__ inline int align16(int x) { return (x + 15) & ~15; }

__host__ __device__ inline CtcSmemLayout ctc_layout(int T, int C, int Lmax, bool lattice_in_smem) {
    CtcSmemLayout s;
    s.Lp1 = Lmax + 1;
    s.U = 2 * Lmax + 1;
    int o = 0;
    s.lab = o;    o = align16(o + 4 * (Lmax > 0 ? Lmax : 1));
    s.next = o;   o = align16(o + 4 * (Lmax > 0 ? Lmax : 1));
    s.head = o;   o = align16(o + 4 * C);
    s.skip_a = o; o = align16(o + s.U);  // 1 byte per state: may take the u-2 / u+2 transition
    s.rowlse = o; o = align16(o + 4 * T);
    s.stage = o;  o = align16(o + 4 * C * kCtcWarps);
    if (lattice_in_smem) {
        s.lpl = o;   o = align16(o + 4 * T * s.Lp1);
        s.alpha = o; o = align16(o + 4 * T * s.U);
        s.beta = o;  o = align16(o + 4 * T * s.U);
    } else {
        s.lpl = s.alpha = s.beta = -1;
    }
    s.total = o;
    return s;
}

__device__ __forceinline__ float lse3(float a, float b, float c) {
    float m = fmaxf(a, fmaxf(b, c));
    if (m == -INFINITY) return -INFINITY;
    return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

template <bool kLatticeInSmem>
__global__ void __launch_bounds__(kCtcThreads)
ctc_loss_kernel(const float* __restrict__ logits, int T, int B, int C, const int32_t* __restrict__ labels,
                const int32_t* __restrict__ label_offsets, const int32_t* __restrict__ seq_len, int Lmax,
                float* __restrict__ loss, float* __restrict__ grad, int32_t* __restrict__ status,
                float grad_scale, float* __restrict__ workspace)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const CtcSmemLayout lay = ctc_layout(T, C, Lmax, kLatticeInSmem);
    int* s_lab = reinterpret_cast<int*>(smem + lay.lab);
    int* s_next = reinterpret_cast<int*>(smem + lay.next);
    int* s_head = reinterpret_cast<int*>(smem + lay.head);
    unsigned char* s_skip = smem + lay.skip_a;
    float* s_rowlse = reinterpret_cast<float*>(smem + lay.rowlse);
    float* s_stage = reinterpret_cast<float*>(smem + lay.stage) + warp * C;
    const int Lp1 = lay.Lp1, Us = lay.U;  // strides (sized for Lmax)
    float *lpl, *A, *Bt;
    if (kLatticeInSmem) {
        lpl = reinterpret_cast<float*>(smem + lay.lpl);
        A = reinterpret_cast<float*>(smem + lay.alpha);
        Bt = reinterpret_cast<float*>(smem + lay.beta);
    } else {
        float* w = workspace + (size_t)b * ((size_t)T * (Lp1 + 2 * Us));
        lpl = w;
        A = w + (size_t)T * Lp1;
        Bt = A + (size_t)T * Us;
    }
    __shared__ float s_logp;
    __shared__ int s_bad;

    const int off = label_offsets[b];
    const int L = label_offsets[b + 1] - off;
    const int U = 2 * L + 1;
    const int Tb = seq_len[b];
    const int blank = C - 1;
    const size_t rstride = (size_t)B * C;
    const float* xb = logits + (size_t)b * C;
    float* gb = grad ? grad + (size_t)b * C : nullptr;

    // ---- phase 0: labels, per-class position lists, feasibility
    if (tid == 0) s_bad = 0;
    for (int k = tid; k < C; k += kCtcThreads) s_head[k] = -1;
    for (int s = tid; s < L; s += kCtcThreads) s_lab[s] = labels[off + s];
    __syncthreads();
    if (tid == 0) {
        int need = L, bad = 0;
        for (int s = L - 1; s >= 0; --s) {
            int l = s_lab[s];
            if (l < 0 || l >= blank) { bad = 3; l = 0; s_lab[s] = 0; }
            s_next[s] = s_head[l];
            s_head[l] = s;
            if (s > 0 && s_lab[s] == s_lab[s - 1]) need++;
        }
        if (!bad && need > Tb && Tb > 0) bad = 2;
        if (Tb < 0 || Tb > T) bad = 3;
        s_bad = bad;
    }
    for (int u = tid; u < U; u += kCtcThreads) {
        // state u (odd = label (u>>1)) may be entered from u-2 iff it is a label differing from the previous label
        s_skip[u] = (u & 1) && (u >= 3) && (labels[off + (u >> 1)] != labels[off + (u >> 1) - 1]);
    }
    __syncthreads();
    const int bad = s_bad;
    if (bad || Tb == 0) {
        // TF: zero-length sequence -> loss 0, grad 0.  Infeasible / invalid -> flagged, zero outputs.
        if (tid == 0) {
            loss[b] = 0.0f;
            if (status) status[b] = bad;
        }
        if (gb)
            for (int t = warp; t < T; t += kCtcWarps)
                for (int k = lane; k < C; k += 32) st_stream(gb + t * rstride + k, 0.0f);
        return;
    }

    // ---- phase 1: stream logit rows, keep row LSE and log y at the label classes + blank
    for (int t = warp; t < Tb; t += kCtcWarps) {
        const float* row = xb + t * rstride;
        float m = -INFINITY;
        for (int k = lane; k < C; k += 32) {
            float v = ld_stream(row + k);
            s_stage[k] = v;
            m = fmaxf(m, v);
        }
        m = warp_max(m);
        float se = 0.0f;
        for (int k = lane; k < C; k += 32) se += expf(s_stage[k] - m);
        se = warp_sum(se);
        const float lse = m + logf(se);
        __syncwarp();
        if (lane == 0) s_rowlse[t] = lse;
        for (int s = lane; s < L; s += 32) lpl[t * Lp1 + s] = s_stage[s_lab[s]] - lse;
        if (lane == 0) lpl[t * Lp1 + L] = s_stage[blank] - lse;
        __syncwarp();
    }
    __syncthreads();

    // ---- phase 2: alpha (warp 0, forward) and beta (warp 1, backward) chains, concurrently
    if (warp == 0) {
        for (int u = lane; u < U; u += 32)
            A[u] = (u == 0) ? lpl[L] : (u == 1 ? lpl[0] : -INFINITY);
        __syncwarp();
        for (int t = 1; t < Tb; ++t) {
            const float* prev = A + (size_t)(t - 1) * Us;
            float* cur = A + (size_t)t * Us;
            const float* lp = lpl + (size_t)t * Lp1;
            for (int u = lane; u < U; u += 32) {
                float a0 = prev[u];
                float a1 = (u > 0) ? prev[u - 1] : -INFINITY;
                float a2 = s_skip[u] ? prev[u - 2] : -INFINITY;
                float l = (u & 1) ? lp[u >> 1] : lp[L];
                cur[u] = l + lse3(a0, a1, a2);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        float* last = Bt + (size_t)(Tb - 1) * Us;
        for (int u = lane; u < U; u += 32) last[u] = (u >= U - 2) ? 0.0f : -INFINITY;
        __syncwarp();
        for (int t = Tb - 2; t >= 0; --t) {
            const float* nxt = Bt + (size_t)(t + 1) * Us;
            float* cur = Bt + (size_t)t * Us;
            const float* lp = lpl + (size_t)(t + 1) * Lp1;
            for (int u = lane; u < U; u += 32) {
                // successors u, u+1, u+2; (u+2) allowed iff state u+2 may be entered from u
                float b0 = nxt[u] + ((u & 1) ? lp[u >> 1] : lp[L]);
                float b1 = (u + 1 < U) ? nxt[u + 1] + (((u + 1) & 1) ? lp[(u + 1) >> 1] : lp[L]) : -INFINITY;
                float b2 = (u + 2 < U && s_skip[u + 2]) ? nxt[u + 2] + lp[(u + 2) >> 1] : -INFINITY;
                cur[u] = lse3(b0, b1, b2);
            }
            __syncwarp();
        }
    }
    __syncthreads();

    // ---- log p(z|x) = LSE_u(alpha(u,0) + beta(u,0))   (CalculateLoss)
    if (warp == 0) {
        float m = -INFINITY;
        for (int u = lane; u < U; u += 32) m = fmaxf(m, A[u] + Bt[u]);
        m = warp_max(m);
        float se = 0.0f;
        if (m != -INFINITY)
            for (int u = lane; u < U; u += 32) se += expf(A[u] + Bt[u] - m);
        se = warp_sum(se);
        if (lane == 0) {
            float lp = (m == -INFINITY) ? -INFINITY : m + logf(se);
            s_logp = lp;
            loss[b] = -lp;
            if (status) status[b] = (lp == -INFINITY) ? 1 : 0;
        }
    }
    __syncthreads();
    if (!gb) return;
    const float logp = s_logp;
    const bool novalid = (logp == -INFINITY);

    // ---- phase 3: gradient rows
    for (int t = warp; t < T; t += kCtcWarps) {
        float* grow = gb + t * rstride;
        if (t >= Tb) {
            for (int k = lane; k < C; k += 32) st_stream(grow + k, 0.0f);
            continue;
        }
        float* a = A + (size_t)t * Us;
        const float* bt = Bt + (size_t)t * Us;
        float bs = 0.0f;
        if (!novalid) {
            for (int u = lane; u < U; u += 32) {
                float e = expf(a[u] + bt[u] - logp);
                a[u] = e;
                if (!(u & 1)) bs += e;
            }
            bs = warp_sum(bs);
        }
        __syncwarp();
        const float lse = s_rowlse[t];
        const float* row = xb + t * rstride;
        for (int k = lane; k < C; k += 32) {
            float y = expf(__ldg(row + k) - lse);
            float occ = 0.0f;
            if (!novalid) {
                if (k == blank) occ = bs;
                else for (int s = s_head[k]; s >= 0; s = s_next[s]) occ += a[2 * s + 1];
            }
            st_stream(grow + k, (y - occ) * grad_scale);
        }
    }
}

}  // namespace ocr

using namespace ocr;

static size_t ctc_ws_floats_per_seq(int T, int Lmax) {
    return (size_t)T * ((size_t)(Lmax + 1) + 2 * (size_t)(2 * Lmax + 1));
}

extern "C" int ocr_ctc_loss_workspace_bytes(int T, int B, int C, int max_label_len, size_t* bytes)
{
    OCR_CHECK_ARG(bytes != nullptr, "ocr_ctc_loss_workspace_bytes: bytes is NULL");
    OCR_CHECK_ARG(T >= 0 && B >= 0 && C >= 2 && max_label_len >= 0, "ocr_ctc_loss_workspace_bytes: bad shape T=%d B=%d C=%d L=%d", T, B, C, max_label_len);
    CtcSmemLayout lay = ctc_layout(T, C, max_label_len, true);
    *bytes = (lay.total <= kMaxDynSmem) ? 0 : (size_t)B * ctc_ws_floats_per_seq(T, max_label_len) * sizeof(float);
    return OCR_OK;
}

extern "C" int ocr_ctc_loss(const float* logits, int T, int B, int C, const int32_t* labels,
                            const int32_t* label_offsets, const int32_t* seq_len, int max_label_len,
                            float* loss, float* grad, int32_t* status, float grad_scale, void* workspace,
                            size_t workspace_bytes, ocr_stream_t stream)
{
    OCR_CHECK_ARG(T >= 0 && B >= 0 && C >= 2 && max_label_len >= 0, "ocr_ctc_loss: bad shape T=%d B=%d C=%d L=%d", T, B, C, max_label_len);
    if (B == 0) return OCR_OK;
    OCR_CHECK_ARG(logits && labels && label_offsets && seq_len && loss, "ocr_ctc_loss: NULL argument");
    OCR_CHECK_ARG(T >= 1, "ocr_ctc_loss: T must be >= 1");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CtcSmemLayout lay = ctc_layout(T, C, max_label_len, true);
    if (lay.total <= kMaxDynSmem) {
        static int configured = -1;
        int dev = 0;
        OCR_CHECK_CUDA(cudaGetDevice(&dev));
        if (configured != dev) {
            OCR_CHECK_CUDA(cudaFuncSetAttribute(ctc_loss_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
            configured = dev;
        }
        ctc_loss_kernel<true><<<B, kCtcThreads, lay.total, st>>>(logits, T, B, C, labels, label_offsets, seq_len,
                                                                  max_label_len, loss, grad, status, grad_scale, nullptr);
    } else {
        size_t need = (size_t)B * ctc_ws_floats_per_seq(T, max_label_len) * sizeof(float);
        if (workspace == nullptr || workspace_bytes < need) {
            set_error("ocr_ctc_loss: workspace too small (%zu < %zu)", workspace_bytes, need);
            return OCR_EWORKSPACE;
        }
        CtcSmemLayout l2 = ctc_layout(T, C, max_label_len, false);
        OCR_CHECK_ARG(l2.total <= kMaxDynSmem, "ocr_ctc_loss: T=%d C=%d too large for shared memory bookkeeping", T, C);
        static int configured2 = -1;
        int dev = 0;
        OCR_CHECK_CUDA(cudaGetDevice(&dev));
        if (configured2 != dev) {
            OCR_CHECK_CUDA(cudaFuncSetAttribute(ctc_loss_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
            configured2 = dev;
        }
        ctc_loss_kernel<false><<<B, kCtcThreads, l2.total, st>>>(logits, T, B, C, labels, label_offsets, seq_len,
                                                                  max_label_len, loss, grad, status, grad_scale,
                                                                  static_cast<float*>(workspace));
    }
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}
