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
#include <algorithm>
#include "common.cuh"
#include "ctc_loss_fast.cuh"
#include "ctc_loss_stream.cuh"

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

// CTA barrier of the general routine: the whole CTA in its own kernel, the first 128 threads (named barrier 9; the fast
// kernel's pair barriers use 1..8) when it runs as the tail of the fast kernel
template <bool kInFast>
__device__ __forceinline__ void gen_sync() {
    if constexpr (kInFast) asm volatile("bar.sync 9, 128;" ::: "memory");
    else __syncthreads();
}

template <bool kLatticeInSmem, bool kInFast>
__device__ __forceinline__ void
ctc_general_one(unsigned char* smem, const int b, const float* __restrict__ logits, int T, int B, int C,
                const int32_t* __restrict__ labels, const int32_t* __restrict__ label_offsets,
                const int32_t* __restrict__ seq_len, int Lmax, float* __restrict__ loss, float* __restrict__ grad,
                int32_t* __restrict__ status, float grad_scale, float* __restrict__ workspace)
{
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
        float* w = workspace + (size_t)blockIdx.x * ((size_t)T * (Lp1 + 2 * Us));  // one slot per (persistent) CTA
        lpl = w;
        A = w + (size_t)T * Lp1;
        Bt = A + (size_t)T * Us;
    }
    __shared__ float s_logp;
    __shared__ int s_bad;
    gen_sync<kInFast>();  // previous sequence of this persistent CTA is done with shared memory

    const int off = label_offsets[b];
    const int Lraw = label_offsets[b + 1] - off;
    // the label arrays and the lattice strides are sized for the caller's max_label_len (Lp1 - 1): a label outside
    // [0, Lmax] is an invalid argument (flagged 3 below), never an overrun
    const bool bad_len = Lraw < 0 || Lraw > Lp1 - 1;
    const int L = bad_len ? 0 : Lraw;
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
    gen_sync<kInFast>();
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
        if (Tb < 0 || Tb > T || bad_len) bad = 3;
        s_bad = bad;
    }
    for (int u = tid; u < U; u += kCtcThreads) {
        // state u (odd = label (u>>1)) may be entered from u-2 iff it is a label differing from the previous label
        s_skip[u] = (u & 1) && (u >= 3) && (labels[off + (u >> 1)] != labels[off + (u >> 1) - 1]);
    }
    gen_sync<kInFast>();
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
    gen_sync<kInFast>();

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
    gen_sync<kInFast>();

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
    gen_sync<kInFast>();
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

// Persistent wrapper: CTA c walks sequences c, c+grid, ...; with only_flagged it recomputes just the
// sequences the fast kernel marked kCtcRedo (normally none: the loop is a flag read per sequence; a chunked gate with
// fewer CTAs scanning flags side by side measured slower at cfg2, 14.35 vs 13.76 us per call).
template <bool kLatticeInSmem>
__global__ void __launch_bounds__(kCtcThreads)
ctc_loss_kernel(const float* __restrict__ logits, int T, int B, int C, const int32_t* __restrict__ labels,
                const int32_t* __restrict__ label_offsets, const int32_t* __restrict__ seq_len, int Lmax,
                float* __restrict__ loss, float* __restrict__ grad, int32_t* __restrict__ status,
                float grad_scale, float* __restrict__ workspace, int only_flagged)
{
    extern __shared__ __align__(16) unsigned char smem[];
    // programmatic dependent launch: whatever follows in the stream may be scheduled now (it waits for this grid's
    // completion itself); this grid waits for its predecessor's results before the first global access
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        if (only_flagged && status[b] != kCtcRedo) continue;
        ctc_general_one<kLatticeInSmem, false>(smem, b, logits, T, B, C, labels, label_offsets, seq_len, Lmax, loss, grad,
                                        status, grad_scale, workspace);
    }
}

}  // namespace ocr

using namespace ocr;

static size_t ctc_ws_floats_per_seq(int T, int Lmax) {
    return (size_t)T * ((size_t)(Lmax + 1) + 2 * (size_t)(2 * Lmax + 1));
}

// path selection: 0 auto (streaming kernel where eligible, else the fast kernel, else the general one), 1 general kernel
// only, 2 fast kernel with LSU loads/stores (no TMA bulk copies), 3 fast kernel leaving its redo flags (diagnostics),
// 4 per-frame bulk copies, 5/6 tensor-map loads / stores only, 7 fast kernel as on path 0 but never the streaming kernel
static int g_ctc_path = 0;
extern "C" int ocr_ctc_loss_set_path(int path) {
    OCR_CHECK_ARG(path >= 0 && path <= 8, "ocr_ctc_loss_set_path: path=%d outside [0,8]", path);
    g_ctc_path = path;
    return OCR_OK;
}

// Tuning aid: buffer of clock64() stamps, kCtcTimelineSlots per warp of every CTA of the fast kernel
// (NULL switches it off).  The caller sizes it for ceil(B/G) CTAs x 2G warps.
static long long* g_ctc_timeline = nullptr;
extern "C" int ocr_debug_ctc_timeline(long long* device_buffer) {
    g_ctc_timeline = device_buffer;
    return OCR_OK;
}

struct FastPlan { int G, NP, CR, bulk, smem; };
static int g_ctc_group = 0;   // tuning override of the group size (0 = automatic)
extern "C" int ocr_debug_ctc_group(int G) {
    OCR_CHECK_ARG(G == 0 || G == 1 || G == 2 || G == 4 || G == 8, "ocr_debug_ctc_group: G=%d", G);
    g_ctc_group = G;
    return OCR_OK;
}

// Launch with the programmatic-stream-serialization attribute: the grid may be scheduled while its predecessor in the
// stream still runs (once every predecessor CTA executed griddepcontrol.launch_dependents or exited) and orders itself
// with griddepcontrol.wait.  The redo gate releases its successor at once, so the next call's fast kernel is resident
// when the gate ends (inside CUDA graphs too: the capture records programmatic edges): cfg2 14.75 -> 13.76 us per call.
// The fast kernel itself never releases early: letting the gate in before the fast grid has exited measured SLOWER
// (16.0 us with the release at its start, 15.6 us with the release before its gradient store).
static int g_ctc_pdl = 1;
extern "C" int ocr_debug_ctc_pdl(int on) {
    g_ctc_pdl = on;   // 2: single-wave fast grids also release their successor at once (tuning)
    return OCR_OK;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_ctc_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

static int g_ctc_speculate = 1;   // tensor-map path: request all boxes of a group before its sequence lengths are known
extern "C" int ocr_debug_ctc_speculate(int on) {
    g_ctc_speculate = on ? 1 : 0;
    return OCR_OK;
}
static int g_ctc_inline_redo = 1;   // flagged sequences are redone in the fast kernel's tail when the exact routine fits there (0: always the gate launch)
extern "C" int ocr_debug_ctc_inline_redo(int on) {
    g_ctc_inline_redo = on ? 1 : 0;
    return OCR_OK;
}

static int g_ctc_prefetch = -1;   // L2 prefetch distance of the fast kernel in CTAs (-1 = half the resident CTAs of the grid, 0 = off)
extern "C" int ocr_debug_ctc_prefetch(int stride) {
    OCR_CHECK_ARG(stride >= -1, "ocr_debug_ctc_prefetch: stride=%d", stride);
    g_ctc_prefetch = stride;
    return OCR_OK;
}

// Chooses the group size G of the fast kernel: the shared-memory footprint is ~G*(T*C + 2*T*(Lmax+1))*4
// bytes; prefer TMA-eligible groups (G*C*4 a multiple of 16 bytes, 16-byte aligned tensors) and as many
// resident sequences per SM as possible.  Returns false when no configuration fits (long T): general kernel.
static bool plan_fast(const void* logits, const void* grad, int T, int B, int C, int Lmax, FastPlan* out) {
    if (Lmax + 1 > 128 || T < 1) return false;
    const int NP = (Lmax + 1 <= 32) ? 1 : ((Lmax + 1 <= 64) ? 2 : 4);
    const int CR = (C <= 64) ? 64 : ((C <= 128) ? 128 : 0);  // softmax rows held in registers up to 128 classes
    const int maxG = (CR == 128) ? 4 : kFastMaxG;              // 128-register rows: 256-thread CTAs
    const bool ptr_ok = ((uintptr_t)logits % 16 == 0) && (grad == nullptr || (uintptr_t)grad % 16 == 0) &&
                        ((long long)B * C) % 4 == 0 && g_ctc_path != 2;  // path 2: LSU loads/stores
    int best = -1, best_score = -1, best_smem = 0, best_bulk = 0;
    for (int G = 1; G <= maxG; G *= 2) {
        if (g_ctc_group != 0 && G != g_ctc_group) continue;
        const FastLayout lay = fast_layout(T, C, Lmax, G);
        if (lay.total + 128 > kMaxDynSmem) break;
        const int bulk = ptr_ok && (G * C) % 4 == 0;
        int per_sm = (227 * 1024) / (lay.total + 1024);
        per_sm = per_sm < 1 ? 1 : per_sm;
        if (per_sm * G * 64 > 2048) per_sm = 2048 / (G * 64);
        int seqs = per_sm * G;
        // a single resident CTA cannot overlap its own loads with compute: weigh it down
        int score = (bulk ? 1000 : 0) + seqs * 4 + (per_sm >= 2 ? 2 : 0) + (G == 4 ? 1 : 0);
        if (score > best_score) { best_score = score; best = G; best_smem = lay.total; best_bulk = bulk; }
    }
    if (best < 0) return false;
    out->G = best; out->NP = NP; out->CR = CR; out->bulk = best_bulk; out->smem = best_smem;
    return true;
}

static int general_grid(int B) { return B < 148 * 4 ? B : 148 * 4; }

// workspace = [B int32 status scratch, 256-byte aligned][general kernel lattice slots, if they do not fit in smem]
static size_t ws_status_bytes(int B) { return (((size_t)B * 4) + 255) & ~(size_t)255; }
static size_t ws_general_bytes(int T, int B, int C, int Lmax) {
    CtcSmemLayout lay = ctc_layout(T, C, Lmax, true);
    return (lay.total <= kMaxDynSmem) ? 0 : (size_t)general_grid(B) * ctc_ws_floats_per_seq(T, Lmax) * sizeof(float);
}

extern "C" int ocr_ctc_loss_workspace_bytes(int T, int B, int C, int max_label_len, size_t* bytes)
{
    OCR_CHECK_ARG(bytes != nullptr, "ocr_ctc_loss_workspace_bytes: bytes is NULL");
    OCR_CHECK_ARG(T >= 0 && B >= 0 && C >= 2 && max_label_len >= 0, "ocr_ctc_loss_workspace_bytes: bad shape T=%d B=%d C=%d L=%d", T, B, C, max_label_len);
    *bytes = ws_status_bytes(B) + ws_general_bytes(T, B, C, max_label_len);
    return OCR_OK;
}

// [T, B*C] fp32 tensor, box {G*C floats, 16 frames}, no swizzle: the staging layout of the fast kernel when its row pitch
// is dense (RS == G*C)
static int ctc_tensor_map(CUtensorMap* tm, const float* base, int T, int B, int C, int G) {
    typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static Enc enc = nullptr;
    if (enc == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            enc = reinterpret_cast<Enc>(p);
    }
    if (enc == nullptr) return OCR_ECUDA;
    cuuint64_t dims[2] = {(cuuint64_t)B * C, (cuuint64_t)T};
    cuuint64_t strides[1] = {(cuuint64_t)B * C * 4};
    cuuint32_t box[2] = {(cuuint32_t)(G * C), (cuuint32_t)kTmRows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? OCR_OK : OCR_ECUDA;
}

static int sms_of(int dev) {
    static int sms = 0, of = -1;
    if (of != dev) {
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        of = dev;
    }
    return sms;
}

template <int NP, int CR>
static int launch_fast(const FastPlan& fp, const float* logits, int T, int B, int C, const int32_t* labels,
                       const int32_t* label_offsets, const int32_t* seq_len, int Lmax, float* loss, float* grad,
                       int32_t* status, float grad_scale, cudaStream_t st, int* redo_inlined)
{
    // tensor-map transfers (one request per 16 frames) when the staging rows are dense, a box fits the TMA limits and
    // whole boxes stay inside the tensor (T a multiple of 16: a partial last box would be clipped and deliver fewer bytes)
    CUtensorMap tmIn, tmOut;
    memset(&tmIn, 0, sizeof(tmIn));
    memset(&tmOut, 0, sizeof(tmOut));
    int bulk = fp.bulk;
    const FastLayout lay = fast_layout(T, C, Lmax, fp.G);
    if (bulk && g_ctc_path != 4 && lay.RS == fp.G * C && fp.G * C <= 256 && (T % kTmRows) == 0 && (B % fp.G) == 0 && grad != nullptr &&
        ctc_tensor_map(&tmIn, logits, T, B, C, fp.G) == OCR_OK && ctc_tensor_map(&tmOut, grad, T, B, C, fp.G) == OCR_OK)
        bulk = (g_ctc_path == 5 || g_ctc_path == 6) ? g_ctc_path : 2;
    static int configured = -1;
    int dev = 0;
    OCR_CHECK_CUDA(cudaGetDevice(&dev));
    if (configured != dev) {
        OCR_CHECK_CUDA(cudaFuncSetAttribute(ctc_loss_fast_kernel<NP, CR>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
        configured = dev;
    }
    const int grid = (B + fp.G - 1) / fp.G;
    // L2 prefetch distance: half the resident CTAs (measured at B = 65536, G = 4, 296 resident: 0 -> 2.33 TB/s,
    // 74..148 -> 2.49, 296 -> 2.43: far enough ahead to land before the successor starts, near enough that it has not started)
    int pf = g_ctc_prefetch;
    int resident = 0;
    if (pf < 0) {
        static int key_smem = -1, key_g = -1, per_sm = 0;   // per instantiation; re-queried when the plan changes
        const int sms = sms_of(dev);
        if (key_smem != fp.smem || key_g != fp.G) {
            OCR_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ctc_loss_fast_kernel<NP, CR>, 64 * fp.G, fp.smem + 128));
            key_smem = fp.smem; key_g = fp.G;
        }
        resident = per_sm * sms;
        pf = resident / 2;
    }
    if (grid <= (resident > 0 ? resident : pf)) pf = 0;   // one wave: nobody comes after
    // flagged sequences (lattice outside the float32 range: rare) are redone by the flagging CTA itself when the exact
    // routine's 128 threads and shared-memory layout fit it; otherwise (and on path 3) the caller launches the redo gate
    const int inl = (g_ctc_inline_redo && g_ctc_path != 3 && 64 * fp.G >= kCtcThreads && ctc_layout(T, C, Lmax, true).total <= fp.smem) ? 1 : 0;
    *redo_inlined = inl;
    if (g_ctc_timeline != nullptr && NP == 1 && CR == 64) {   // the instantiation with the phase marks compiled in (tuning aid)
        static int configured_tl = -1;
        if (configured_tl != dev) {
            OCR_CHECK_CUDA(cudaFuncSetAttribute(ctc_loss_fast_kernel<1, 64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
            configured_tl = dev;
        }
    OCR_CHECK_CUDA(launch_pdl(ctc_loss_fast_kernel<1, 64, true>, grid, 64 * fp.G, (size_t)fp.smem + 128, st, logits, T, B, C, labels, label_offsets,
                              seq_len, Lmax, fp.G, bulk, loss, grad, status, grad_scale, tmIn, tmOut, pf, inl | ((g_ctc_pdl == 2 && grid * 2 <= sms_of(dev)) ? 2 : 0) | (g_ctc_speculate ? 4 : 0),
                              g_ctc_timeline, lay));
        count_launch();
        return OCR_OK;
    }
    OCR_CHECK_CUDA(launch_pdl(ctc_loss_fast_kernel<NP, CR>, grid, 64 * fp.G, (size_t)fp.smem + 128, st, logits, T, B, C, labels, label_offsets,
                              seq_len, Lmax, fp.G, bulk, loss, grad, status, grad_scale, tmIn, tmOut, pf, inl | ((g_ctc_pdl == 2 && grid * 2 <= sms_of(dev)) ? 2 : 0) | (g_ctc_speculate ? 4 : 0),
                              g_ctc_timeline, lay));
    count_launch();
    return OCR_OK;
}

struct StreamPlan { int G, nbuf, smem; };
static int g_ctc_stream_nbuf = 0;   // tuning override of the ring depth (0 = automatic)
extern "C" int ocr_debug_ctc_stream_nbuf(int n) {
    OCR_CHECK_ARG(n >= 0 && n <= kStreamMaxBuf, "ocr_debug_ctc_stream_nbuf: n=%d", n);
    g_ctc_stream_nbuf = n;
    return OCR_OK;
}

// The streaming kernel takes TMA-eligible tensors (16-byte aligned, G*C*4 a multiple of 16 bytes, whole groups), labels up
// to 31 (one state pair per lane) and up to 67 classes (a row passes through 2 x 32 registers).  Per sequence it holds T*(Lmax+1) + T*(2*Lmax+5) floats.
static bool plan_stream(const void* logits, const void* grad, int T, int B, int C, int Lmax, int dev, StreamPlan* out) {
    if (g_ctc_path != 8 || Lmax + 1 > 32 || C > 67 || T < kTmRows) return false;   // two lanes x 8 float4 per row (+ up to 3 scalars)
    if (((uintptr_t)logits % 16) != 0 || (grad != nullptr && ((uintptr_t)grad % 16) != 0) || ((long long)B * C) % 4 != 0) return false;
    const int NC = (T + kTmRows - 1) / kTmRows;
    static const int order[3] = {4, 2, 8};
    for (int k = 0; k < 3; ++k) {
        const int G = order[k];
        if (g_ctc_group != 0 && G != g_ctc_group) continue;
        if ((G * C) % 4 != 0 || G * C > 256 || B % G != 0) continue;
        const int grid = B / G;
        // one wave: the whole sequence block in flight at once (nothing else competes for the SM's shared memory)
        int nbuf = g_ctc_stream_nbuf ? g_ctc_stream_nbuf : (grid <= sms_of(dev) ? NC : 2);
        nbuf = nbuf > kStreamMaxBuf ? kStreamMaxBuf : nbuf;
        nbuf = nbuf > NC ? NC : nbuf;
        if (nbuf < NC && nbuf > 3) nbuf = 3;   // a reused slot must stay with one front warp group (at most three of them work)
        for (; nbuf >= 1; --nbuf) {
            const StreamLayout lay = stream_layout(T, C, Lmax, G, nbuf);
            if (lay.total + 128 <= kMaxDynSmem) {
                out->G = G; out->nbuf = nbuf; out->smem = lay.total + 128;
                return true;
            }
        }
    }
    return false;
}

template <int G>
static int launch_stream(const StreamPlan& sp, const float* logits, int T, int B, int C, const int32_t* labels,
                         const int32_t* label_offsets, const int32_t* seq_len, int Lmax, float* loss, float* grad,
                         int32_t* status, float grad_scale, cudaStream_t st, int dev, int* redo_inlined, bool* launched)
{
    *launched = false;
    CUtensorMap tmIn, tmOut;
    memset(&tmIn, 0, sizeof(tmIn));
    memset(&tmOut, 0, sizeof(tmOut));
    if (ctc_tensor_map(&tmIn, logits, T, B, C, G) != OCR_OK || (grad != nullptr && ctc_tensor_map(&tmOut, grad, T, B, C, G) != OCR_OK)) return OCR_OK;
    static int configured = -1;
    if (configured != dev) {
        OCR_CHECK_CUDA(cudaFuncSetAttribute(ctc_loss_stream_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
        configured = dev;
    }
    const int grid = B / G, threads = 64 * G;
    int pf = g_ctc_prefetch, resident = 0;
    if (pf < 0) {
        static int key_smem = -1, per_sm = 0;
        if (key_smem != sp.smem) {
            OCR_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ctc_loss_stream_kernel<G>, threads, sp.smem));
            key_smem = sp.smem;
        }
        resident = per_sm * sms_of(dev);
        pf = resident / 2;
    }
    if (grid <= (resident > 0 ? resident : pf)) pf = 0;   // one wave: nobody comes after
    const int inl = (g_ctc_inline_redo && 64 * G >= kCtcThreads && ctc_layout(T, C, Lmax, true).total <= sp.smem - 128) ? 1 : 0;
    *redo_inlined = inl;
    OCR_CHECK_CUDA(launch_pdl(ctc_loss_stream_kernel<G>, grid, threads, (size_t)sp.smem, st, logits, T, B, C, labels, label_offsets,
                              seq_len, Lmax, sp.nbuf, loss, grad, status, grad_scale, tmIn, tmOut, pf, inl, g_ctc_timeline, stream_layout(T, C, Lmax, G, sp.nbuf)));
    count_launch();
    *launched = true;
    return OCR_OK;
}

static int launch_general(const float* logits, int T, int B, int C, const int32_t* labels, const int32_t* label_offsets,
                          const int32_t* seq_len, int Lmax, float* loss, float* grad, int32_t* status, float grad_scale,
                          float* lattice_ws, int only_flagged, cudaStream_t st)
{
    CtcSmemLayout lay = ctc_layout(T, C, Lmax, true);
    int dev = 0;
    OCR_CHECK_CUDA(cudaGetDevice(&dev));
    const int grid = general_grid(B);
    if (lay.total <= kMaxDynSmem) {
        static int configured = -1;
        if (configured != dev) {
            OCR_CHECK_CUDA(cudaFuncSetAttribute(ctc_loss_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
            configured = dev;
        }
        OCR_CHECK_CUDA(launch_pdl(ctc_loss_kernel<true>, grid, kCtcThreads, (size_t)lay.total, st, logits, T, B, C, labels, label_offsets,
                                  seq_len, Lmax, loss, grad, status, grad_scale, (float*)nullptr, only_flagged));
    } else {
        CtcSmemLayout l2 = ctc_layout(T, C, Lmax, false);
        OCR_CHECK_ARG(l2.total <= kMaxDynSmem, "ocr_ctc_loss: T=%d C=%d too large for shared memory bookkeeping", T, C);
        static int configured2 = -1;
        if (configured2 != dev) {
            OCR_CHECK_CUDA(cudaFuncSetAttribute(ctc_loss_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
            configured2 = dev;
        }
        OCR_CHECK_CUDA(launch_pdl(ctc_loss_kernel<false>, grid, kCtcThreads, (size_t)l2.total, st, logits, T, B, C, labels, label_offsets,
                                  seq_len, Lmax, loss, grad, status, grad_scale, lattice_ws, only_flagged));
    }
    count_launch();
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
    const size_t need = ws_status_bytes(B) + ws_general_bytes(T, B, C, max_label_len);
    if (workspace == nullptr || workspace_bytes < need) {
        set_error("ocr_ctc_loss: workspace too small (%zu < %zu)", workspace_bytes, need);
        return OCR_EWORKSPACE;
    }
    int32_t* st_buf = status ? status : static_cast<int32_t*>(workspace);
    float* lattice_ws = reinterpret_cast<float*>(static_cast<unsigned char*>(workspace) + ws_status_bytes(B));
    FastPlan fp;
    StreamPlan sp;
    int dev = 0;
    OCR_CHECK_CUDA(cudaGetDevice(&dev));
    if (grad_scale > 0.0f && plan_stream(logits, grad, T, B, C, max_label_len, dev, &sp)) {
        int rc = OCR_OK, inlined = 0;
        bool launched = false;
#define OCR_STREAM(G_) rc = launch_stream<G_>(sp, logits, T, B, C, labels, label_offsets, seq_len, max_label_len, loss, grad, st_buf, grad_scale, st, dev, &inlined, &launched)
        if (sp.G == 2) OCR_STREAM(2); else if (sp.G == 4) OCR_STREAM(4); else OCR_STREAM(8);
#undef OCR_STREAM
        if (rc != OCR_OK) return rc;
        if (launched) {
            if (inlined) return OCR_OK;
            return launch_general(logits, T, B, C, labels, label_offsets, seq_len, max_label_len, loss, grad, st_buf,
                                  grad_scale, lattice_ws, 1, st);
        }
        // no tensor map for this tensor: the fast kernel below takes it
    }
    // the fast kernel stages y * grad_scale and runs its lattice on it: it needs a positive scale
    if (g_ctc_path != 1 && grad_scale > 0.0f && plan_fast(logits, grad, T, B, C, max_label_len, &fp)) {
        int rc;
        int inlined = 0;
#define OCR_FAST(NP_, CR_) rc = launch_fast<NP_, CR_>(fp, logits, T, B, C, labels, label_offsets, seq_len, max_label_len, loss, grad, st_buf, grad_scale, st, &inlined)
        if (fp.CR == 64) { if (fp.NP == 1) OCR_FAST(1, 64); else if (fp.NP == 2) OCR_FAST(2, 64); else OCR_FAST(4, 64); }
        else if (fp.CR == 128) { if (fp.NP == 1) OCR_FAST(1, 128); else if (fp.NP == 2) OCR_FAST(2, 128); else OCR_FAST(4, 128); }
        else { if (fp.NP == 1) OCR_FAST(1, 0); else if (fp.NP == 2) OCR_FAST(2, 0); else OCR_FAST(4, 0); }
#undef OCR_FAST
        if (rc != OCR_OK || g_ctc_path == 3 || inlined) return rc;  // path 3 (diagnostics): leave kCtcRedo flags in status
        // sequences whose lattice left the float32 range of the fast kernel (status kCtcRedo): exact kernel
        return launch_general(logits, T, B, C, labels, label_offsets, seq_len, max_label_len, loss, grad, st_buf,
                              grad_scale, lattice_ws, 1, st);
    }
    return launch_general(logits, T, B, C, labels, label_offsets, seq_len, max_label_len, loss, grad, st_buf, grad_scale,
                          lattice_ws, 0, st);
}
