// Persistent bidirectional GRU recurrence for sm_100a: the frame loop of tf.nn.bidirectional_dynamic_rnn over
// tf.contrib.rnn.GRUCell (/root/reference/src/weinman/model.py:167-199 -- the cell the reference's server loads) as ONE
// kernel launch per layer.  Same construction as lstm_persistent.cu (weights resident in shared memory, h streamed by TMA,
// tcgen05.mma.kind::tf32 into TMEM, TMEM lane = batch row, per-frame grid barrier on a global counter), with the one
// thing the GRU adds: TensorFlow's GRUCell applies the reset gate BEFORE the candidate product,
//     r, u = sigmoid([x, h] W_g + b_g);   c = tanh([x, r*h] W_c + b_c);   h' = u*h + (1-u)*c,
// so a frame is TWO dependent products with a grid-wide exchange between them (every CTA needs all of r*h):
//   phase A  h_{t-1} (TMA) x the CTA's 2*hs gate rows of W_g  -> r, u for its hs units; r*h goes to global memory;
//   phase B  r*h (TMA, after the direction's CTAs have met) x its hs rows of W_c -> candidate, cell update, h_t, output.
// CTA (d, j) keeps its 3*hs weight rows (r | u | candidate, 96 KB for hs = 16, H = 512) resident for the whole sequence;
// u and h stay in registers between the phases.  Against the launch-per-frame path (two GEMM launches + two cell kernels
// per frame) the 61-frame recognizer goes from 2.45 to about half of that per batch of 32 (bench.py block inference_gru).
// F16: h, r*h and the weights as IEEE binary16 operands (tcgen05.mma.kind::f16, K = 16 per instruction), as in lstm_persistent.cu.
#include <cuda_fp16.h>

#include <type_traits>

#include "gemm_tf32.cuh"

namespace ocr {

// 16 consecutive state values of a batch row -> the operand buffer (float32 or binary16)
template <bool F16, int HS>
__device__ __forceinline__ void gru_store_operand(void* base, size_t off, const float (&v)[HS]) {
    if constexpr (F16) {
        __half* p = reinterpret_cast<__half*>(base) + off;
#pragma unroll
        for (int u = 0; u < HS; u += 8) {
            const __half2 p0 = __floats2half2_rn(v[u], v[u + 1]), p1 = __floats2half2_rn(v[u + 2], v[u + 3]);
            const __half2 p2 = __floats2half2_rn(v[u + 4], v[u + 5]), p3 = __floats2half2_rn(v[u + 6], v[u + 7]);
            *reinterpret_cast<uint4*>(p + u) = make_uint4(*reinterpret_cast<const unsigned*>(&p0), *reinterpret_cast<const unsigned*>(&p1),
                                                          *reinterpret_cast<const unsigned*>(&p2), *reinterpret_cast<const unsigned*>(&p3));
        }
    } else {
        float* p = reinterpret_cast<float*>(base) + off;
#pragma unroll
        for (int u = 0; u < HS; u += 4) *reinterpret_cast<float4*>(p + u) = make_float4(v[u], v[u + 1], v[u + 2], v[u + 3]);
    }
}

constexpr int kGruMaxStages = 32;
constexpr int kGruThreads = 192;

__device__ __forceinline__ float gru_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float gru_sigm(float x) { return fmaf(0.5f, gru_tanh(0.5f * x), 0.5f); }

// bounded spin on a global counter (acquire)
__device__ __forceinline__ void gru_wait_counter(const unsigned* ctr, unsigned target) {
    for (unsigned it = 0; it < (1u << 27); ++it) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        if (v >= target) return;
    }
    __trap();
}

template <int NC>
__device__ __forceinline__ void gru_tmem_ld(unsigned taddr, unsigned (&g)[NC]) {
#pragma unroll
    for (int c0 = 0; c0 < NC; c0 += 16) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(g[c0 + 0]), "=r"(g[c0 + 1]), "=r"(g[c0 + 2]), "=r"(g[c0 + 3]), "=r"(g[c0 + 4]), "=r"(g[c0 + 5]), "=r"(g[c0 + 6]),
              "=r"(g[c0 + 7]), "=r"(g[c0 + 8]), "=r"(g[c0 + 9]), "=r"(g[c0 + 10]), "=r"(g[c0 + 11]), "=r"(g[c0 + 12]),
              "=r"(g[c0 + 13]), "=r"(g[c0 + 14]), "=r"(g[c0 + 15])
            : "r"(taddr + (unsigned)c0) : "memory");
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int HS, bool F16>  // hidden units per CTA: 2*HS gate columns in phase A, HS candidate columns in phase B; F16: binary16 operands
__global__ void __launch_bounds__(kGruThreads, 1)
gru_persistent_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH00,
                      const __grid_constant__ CUtensorMap tmH01, const __grid_constant__ CUtensorMap tmH10,
                      const __grid_constant__ CUtensorMap tmH11, const __grid_constant__ CUtensorMap tmR0,
                      const __grid_constant__ CUtensorMap tmR1, const float* __restrict__ xp /*[T*B, 6H]*/,
                      const int32_t* __restrict__ seq_len, void* __restrict__ hbuf /*[2 parity][2 dir][B][H], float32 or binary16*/,
                      void* __restrict__ rhbuf /*[2 dir][B][H]*/, float* __restrict__ out /*[T,B,2H]*/,
                      unsigned* __restrict__ counters /*[2]*/, int T, int B, int H, int NS, int a_rows, int n_stages, int gc, int MT)
{
    constexpr int NA = 2 * HS, NB = HS, NWR = 3 * HS;   // MMA N of the two phases, resident weight rows per k-chunk
    constexpr int BK = F16 ? kGemmBKh : kGemmBK;   // elements per 128-byte swizzle row
    constexpr unsigned kRowB = 128;
    const int nk = H / BK;
    extern __shared__ unsigned char gru_smem_raw[];
    unsigned char* smem = gru_smem_raw + ((1024u - (g_smem_u32(gru_smem_raw) & 1023u)) & 1023u);
    const unsigned s_base = g_smem_u32(smem);
    const unsigned w_bytes = (unsigned)NWR * kRowB;        // one k-chunk of the weight slice: rows r | u | candidate
    const unsigned wc_off = (unsigned)NA * kRowB;          // the candidate rows inside it (a whole number of 8-row swizzle atoms)
    const unsigned a_bytes = (unsigned)a_rows * kRowB;     // one k-chunk of the operand (h or r*h): the a_rows real batch rows
    const unsigned s_w = s_base;
    const unsigned s_a = s_w + (unsigned)nk * w_bytes;
    const unsigned g_bytes = (unsigned)gc * a_bytes;
    const int ng = nk / gc;
    const unsigned s_bar = s_a + (unsigned)n_stages * g_bytes + (a_rows < kGemmBM ? kGemmBM * kRowB : 0);
    const unsigned bar_full = s_bar, bar_empty = s_bar + kGruMaxStages * 8, bar_w = bar_empty + kGruMaxStages * 8, bar_acc = bar_w + 8;   // bar_acc: [2]
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(smem + (s_bar - s_base) + (2 * kGruMaxStages + 3) * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = blockIdx.x % MT, j = (blockIdx.x / MT) % NS, d = blockIdx.x / (MT * NS);
    const int m0 = mt * kGemmBM;
    // two issuing threads when the whole operand block is one resident group (see lstm_persistent.cu): the producer lane issues
    // the odd k-chunks into a second accumulator, the epilogue adds the two
    const bool dual = n_stages == 1 && gc == nk && nk >= 2;
    // TMEM columns: phase A accumulators at 0 and NA, phase B accumulators at 2*NA and 2*NA + NB
    constexpr unsigned kTmemCols = 128;
    static_assert(2 * NA + 2 * NB <= 128, "TMEM columns");

    if (threadIdx.x == 0) {
        for (int s = 0; s < n_stages; ++s) { g_mbar_init(bar_full + s * 8, 1); g_mbar_init(bar_empty + s * 8, dual ? 2 : 1); }
        g_mbar_init(bar_w, 1);
        g_mbar_init(bar_acc, dual ? 2 : 1);
        g_mbar_init(bar_acc + 8, dual ? 2 : 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(g_smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_d = *tmem_slot;
    // instruction descriptors: float32 accumulator; operand formats TF32 (2) or binary16 (0); N, M
    constexpr unsigned kFmt = ((F16 ? 0u : 2u) << 7) | ((F16 ? 0u : 2u) << 10);
    const unsigned idescA = (1u << 4) | kFmt | ((unsigned)(NA >> 3) << 17) | ((unsigned)(kGemmBM >> 4) << 24);
    const unsigned idescB = (1u << 4) | kFmt | ((unsigned)(NB >> 3) << 17) | ((unsigned)(kGemmBM >> 4) << 24);
    auto mma = [](unsigned td, unsigned long long da, unsigned long long db, unsigned idesc, unsigned acc) {
        if constexpr (F16) umma_f16(td, da, db, idesc, acc);
        else umma_tf32(td, da, db, idesc, acc);
    };

    if (warp == 0) {
        if (lane == 0) {
            // resident weight slice: rows [(d*NS + j)*3*HS, +3*HS) of the permuted recurrent weights
            g_mbar_expect_tx(bar_w, (unsigned)nk * w_bytes);
            for (int k = 0; k < nk; ++k) tma_load_2d(s_w + k * w_bytes, &tmW, k * BK, (d * NS + j) * NWR, bar_w);
            int it = 0;
            for (int hf = 0; hf < 2 * T; ++hf) {   // half-frames: phase A of frame s = hf/2, then its phase B
                const int s = hf >> 1, ph = hf & 1;
                {   // the first group's stage is armed before the grid barrier (see lstm_persistent.cu)
                    const int st0 = it % n_stages;
                    if (it >= n_stages) g_mbar_wait(bar_empty + st0 * 8, ((it / n_stages) - 1) & 1);
                    g_mbar_expect_tx(bar_full + st0 * 8, g_bytes);
                }
                if (hf > 0) {
                    gru_wait_counter(counters + d, (unsigned)(NS * MT) * (unsigned)hf);   // every slice / batch tile of this direction published its part
                    asm volatile("fence.proxy.async.global;" ::: "memory");              // generic-proxy writes -> async-proxy (TMA) reads (global state space only)
                }
                const CUtensorMap* tm = ph ? (d ? &tmR1 : &tmR0) : ((s & 1) ? (d ? &tmH11 : &tmH10) : (d ? &tmH01 : &tmH00));
                for (int gi = 0; gi < ng; ++gi, ++it) {
                    const int st = it % n_stages;
                    if (gi > 0) {
                        if (it >= n_stages) g_mbar_wait(bar_empty + st * 8, ((it / n_stages) - 1) & 1);
                        g_mbar_expect_tx(bar_full + st * 8, g_bytes);
                    }
                    tma_load_3d(s_a + st * g_bytes, tm, 0, m0, gi * gc, bar_full + st * 8);
                }
                if (dual) {   // second MMA issuer: odd k-chunks -> the phase's second accumulator
                    if (hf == 0) g_mbar_wait(bar_w, 0);
                    g_mbar_wait(bar_full, hf & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const unsigned acc = tmem_d + (ph ? (unsigned)(2 * NA + NB) : (unsigned)NA);
                    for (int k = 1; k < nk; k += 2) {
                        const unsigned long long da = umma_desc_k128(s_a + k * a_bytes), db = umma_desc_k128(s_w + k * w_bytes + (ph ? wc_off : 0u));
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)   // 32 bytes of the swizzle row per MMA (8 tf32 / 16 binary16)
                            mma(acc, da + (unsigned long long)(kk * 2), db + (unsigned long long)(kk * 2), ph ? idescB : idescA, (k > 1 || kk) ? 1u : 0u);
                    }
                    umma_commit(bar_empty);
                    umma_commit(bar_acc + ph * 8);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            g_mbar_wait(bar_w, 0);
            int it = 0;
            for (int hf = 0; hf < 2 * T; ++hf) {
                const int ph = hf & 1;
                const unsigned acc = tmem_d + (ph ? (unsigned)(2 * NA) : 0u);
                for (int gi = 0; gi < ng; ++gi, ++it) {
                    const int st = it % n_stages;
                    g_mbar_wait(bar_full + st * 8, (it / n_stages) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int c = 0; c < gc; c += (dual ? 2 : 1)) {   // dual: even chunks here, odd chunks on the producer lane
                        const int k = gi * gc + c;
                        const unsigned long long da = umma_desc_k128(s_a + st * g_bytes + c * a_bytes), db = umma_desc_k128(s_w + k * w_bytes + (ph ? wc_off : 0u));
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            mma(acc, da + (unsigned long long)(kk * 2), db + (unsigned long long)(kk * 2), ph ? idescB : idescA, (k | kk) ? 1u : 0u);
                    }
                    umma_commit(bar_empty + st * 8);
                }
                umma_commit(bar_acc + ph * 8);   // the phase's pre-activations are in TMEM
            }
        }
    } else {
        const int q = warp & 3;
        const int r = m0 + q * 32 + lane;      // batch row (TMEM lane q*32 + lane of this batch tile)
        const bool live_row = r < B;
        const int len = live_row ? min(max(seq_len[r], 0), T) : 0;
        float h[HS], uu[HS];
#pragma unroll
        for (int u = 0; u < HS; ++u) { h[u] = 0.0f; uu[u] = 0.0f; }
        for (int s = 0; s < T; ++s) {
            // the input projection of this frame does not depend on the recurrent products: fetch it while the MMAs run
            const bool upd = live_row && s < len;
            const int t = d ? len - 1 - s : s;
            float xr[HS], xu[HS], xc[HS];
            if (upd) {
                const float* x = xp + ((size_t)t * B + r) * 6 * H + (size_t)d * 3 * H + j * HS;
#pragma unroll
                for (int u = 0; u < HS; u += 4) {
                    *reinterpret_cast<float4*>(xr + u) = __ldg(reinterpret_cast<const float4*>(x + u));
                    *reinterpret_cast<float4*>(xu + u) = __ldg(reinterpret_cast<const float4*>(x + H + u));
                    *reinterpret_cast<float4*>(xc + u) = __ldg(reinterpret_cast<const float4*>(x + 2 * H + u));
                }
            }
            // ---- phase A: r, u; publish r*h
            g_mbar_wait(bar_acc, s & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {
                unsigned g[NA];
                gru_tmem_ld<NA>(tmem_d + ((unsigned)(q * 32) << 16), g);
                if (dual) {
                    unsigned g2[NA];
                    gru_tmem_ld<NA>(tmem_d + ((unsigned)(q * 32) << 16) + (unsigned)NA, g2);
#pragma unroll
                    for (int i = 0; i < NA; ++i) g[i] = __float_as_uint(__uint_as_float(g[i]) + __uint_as_float(g2[i]));
                }
                if (live_row) {
                    float rh[HS];
#pragma unroll
                    for (int u = 0; u < HS; ++u) {
                        rh[u] = 0.0f;
                        if (upd) {
                            const float rr = gru_sigm(__uint_as_float(g[u]) + xr[u]);
                            uu[u] = gru_sigm(__uint_as_float(g[HS + u]) + xu[u]);
                            rh[u] = rr * h[u];
                        }
                    }
                    gru_store_operand<F16, HS>(rhbuf, ((size_t)d * B + r) * H + j * HS, rh);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps: their stores are ordered before...
            if (warp == 2 && lane == 0)                      // ...this gpu-scope release (cumulative) that publishes the slice
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counters + d) : "memory");
            // ---- phase B: candidate, cell update, h_t
            g_mbar_wait(bar_acc + 8, s & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {
                unsigned g[NB];
                gru_tmem_ld<NB>(tmem_d + ((unsigned)(q * 32) << 16) + (unsigned)(2 * NA), g);
                if (dual) {
                    unsigned g2[NB];
                    gru_tmem_ld<NB>(tmem_d + ((unsigned)(q * 32) << 16) + (unsigned)(2 * NA + NB), g2);
#pragma unroll
                    for (int i = 0; i < NB; ++i) g[i] = __float_as_uint(__uint_as_float(g[i]) + __uint_as_float(g2[i]));
                }
                if (live_row) {
                    if (upd) {
#pragma unroll
                        for (int u = 0; u < HS; ++u) {
                            const float cand = gru_tanh(__uint_as_float(g[u]) + xc[u]);
                            h[u] = fmaf(uu[u], h[u] - cand, cand);   // u*h + (1-u)*c
                        }
                    }
                    // carried or updated, the state is the next frame's operand
                    gru_store_operand<F16, HS>(hbuf, (((size_t)((s + 1) & 1) * 2 + d) * B + r) * H + j * HS, h);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (warp == 2 && lane == 0)
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counters + d) : "memory");
            if (upd) {   // the layer output leaves after the slice is published (the release waits for the stores before it)
                float* o = out + ((size_t)t * B + r) * 2 * H + (size_t)d * H + j * HS;
#pragma unroll
                for (int u = 0; u < HS; u += 4) *reinterpret_cast<float4*>(o + u) = make_float4(h[u], h[u + 1], h[u + 2], h[u + 3]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols) : "memory");
    }
}

// slice-major permutation of the recurrent weights: row ((d*NS + j)*3 + g)*HS + u  <-  g < 2: whg[d*2H + g*H + j*HS + u], g = 2: whc[d*H + j*HS + u]
template <typename TO>
__global__ void permute_gru_wh_kernel(const float* __restrict__ whg, const float* __restrict__ whc, TO* __restrict__ whp, int H, int HS, int NS)
{
    const long long total = (long long)6 * H * H;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(idx % H);
        long long row = idx / H;
        const int u = (int)(row % HS); row /= HS;
        const int g = (int)(row % 3); row /= 3;
        const int j = (int)(row % NS);
        const int d = (int)(row / NS);
        const float v = g < 2 ? whg[((size_t)d * 2 * H + (size_t)g * H + j * HS + u) * H + k] : whc[((size_t)d * H + j * HS + u) * H + k];
        if constexpr (std::is_same<TO, __half>::value) whp[idx] = __float2half_rn(v);
        else whp[idx] = v;
    }
}

}  // namespace ocr

using namespace ocr;

constexpr int kGruHS = 16;

namespace ocr {

static int g_gru_f16 = 1;   // binary16 recurrent operands (see lstm_persistent.cu)
int gru_set_operands(int f16) {
    g_gru_f16 = f16 ? 1 : 0;
    return OCR_OK;
}

static bool gru_f16(int H) { return g_gru_f16 && (H % kGemmBKh) == 0; }

bool gru_persistent_supported(int T, int B, int H) {
    if ((H % kGemmBK) != 0 || (H % kGruHS) != 0 || B < 1) return false;
    const int NS = H / kGruHS, MT = (B + kGemmBM - 1) / kGemmBM;
    if (2 * NS * MT > 148) return false;   // one CTA per SM, all co-resident
    const int bk = gru_f16(H) ? kGemmBKh : kGemmBK;
    if (H / bk > kGruMaxStages) return false;
    const size_t w = (size_t)H / bk * (3 * kGruHS) * 128;
    const size_t a = (size_t)(B >= kGemmBM ? kGemmBM : (B + 7) / 8 * 8) * 128;
    return w + 2 * a + kGemmBM * 128 + 1024 + 1024 <= (size_t)kMaxDynSmem && T >= 1;
}

size_t gru_persistent_workspace_floats(int B, int H) {
    // permuted weights [6H, H] + h double buffer [2][2][B][H] + r*h [2][B][H] + counters (64 floats); the binary16 forms use half of each
    return (size_t)6 * H * H + (size_t)6 * B * H + 64;
}

template <bool F16>
static int gru_launch(const cudaLaunchConfig_t& cfg, const CUtensorMap& tmW, const CUtensorMap (&tmH)[2][2], const CUtensorMap (&tmR)[2],
                      const float* xp, const int32_t* seq_len, void* hbuf, void* rhbuf, float* out, unsigned* counters, int T, int B, int H,
                      int NS, int a_rows, int n_stages, int gc, int MT)
{
    static int configured = -1;
    int dev = 0;
    OCR_CHECK_CUDA(cudaGetDevice(&dev));
    if (configured != dev) {
        OCR_CHECK_CUDA(cudaFuncSetAttribute(gru_persistent_kernel<kGruHS, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
        configured = dev;
    }
    OCR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gru_persistent_kernel<kGruHS, F16>, tmW, tmH[0][0], tmH[0][1], tmH[1][0], tmH[1][1], tmR[0], tmR[1], xp,
                                      seq_len, hbuf, rhbuf, out, counters, T, B, H, NS, a_rows, n_stages, gc, MT));
    count_launch();
    return OCR_OK;
}

// xp [T*B, 6H] (input projection + bias; per direction r | u | candidate), whg [4H, H], whc [2H, H], out [T,B,2H] (zeroed here)
int gru_persistent_run(const float* xp, const float* whg, const float* whc, const int32_t* seq_len, int T, int B, int H,
                       float* out, float* ws, cudaStream_t st)
{
    const int NS = H / kGruHS, MT = (B + kGemmBM - 1) / kGemmBM;
    const bool f16 = gru_f16(H);
    const size_t esz = f16 ? 2 : 4;
    float* whp = ws;
    float* hbuf = whp + (size_t)6 * H * H;
    unsigned char* rhbuf = reinterpret_cast<unsigned char*>(hbuf) + (size_t)4 * B * H * esz;
    unsigned* counters = reinterpret_cast<unsigned*>(hbuf + (size_t)6 * B * H);
    {
        const long long total = (long long)6 * H * H;
        long long gsz = (total + 255) / 256;
        const int grid = (int)(gsz > 148 * 16 ? 148 * 16 : gsz);
        if (f16) permute_gru_wh_kernel<__half><<<grid, 256, 0, st>>>(whg, whc, reinterpret_cast<__half*>(whp), H, kGruHS, NS);
        else permute_gru_wh_kernel<float><<<grid, 256, 0, st>>>(whg, whc, whp, H, kGruHS, NS);
        OCR_CHECK_LAUNCH();
    }
    OCR_CHECK_CUDA(cudaMemsetAsync(hbuf, 0, sizeof(float) * ((size_t)6 * B * H + 64), st));
    OCR_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)T * B * 2 * H, st));
    const int bk = f16 ? kGemmBKh : kGemmBK;
    const int nk = H / bk;
    const int a_rows = B >= kGemmBM ? kGemmBM : (B + 7) / 8 * 8;
    const size_t w_bytes = (size_t)nk * (3 * kGruHS) * 128, a_bytes = (size_t)a_rows * 128;
    const size_t fixed = w_bytes + (a_rows < kGemmBM ? kGemmBM * 128 : 0) + 1024 + 1024;
    const int fit = (int)(((size_t)kMaxDynSmem - fixed) / a_bytes);     // chunk tiles that fit beside the weights
    int gc = nk, n_stages = 1;
    if (fit < nk) {
        gc = 1;
        for (int c = 1; c <= nk; ++c)
            if (nk % c == 0 && 2 * c <= fit) gc = c;
        n_stages = fit / gc;
        if (n_stages > nk / gc) n_stages = nk / gc;
        if (n_stages > kGruMaxStages) n_stages = kGruMaxStages;
    }
    CUtensorMap tmW, tmH[2][2], tmR[2];
    int rc = f16 ? tma_map_2d_h(&tmW, whp, (long long)6 * H, H, H, 3 * kGruHS) : tma_map_2d(&tmW, whp, (long long)6 * H, H, H, 3 * kGruHS);
    if (rc != OCR_OK) return rc;
    for (int p = 0; p < 2; ++p)
        for (int d = 0; d < 2; ++d) {
            const size_t off = ((size_t)p * 2 + d) * B * H;
            rc = f16 ? tma_map_chunks_h(&tmH[p][d], reinterpret_cast<__half*>(hbuf) + off, B, H, H, a_rows, gc)
                     : tma_map_chunks(&tmH[p][d], hbuf + off, B, H, H, a_rows, gc);
            if (rc != OCR_OK) return rc;
        }
    for (int d = 0; d < 2; ++d) {
        const size_t off = (size_t)d * B * H;
        rc = f16 ? tma_map_chunks_h(&tmR[d], reinterpret_cast<__half*>(rhbuf) + off, B, H, H, a_rows, gc)
                 : tma_map_chunks(&tmR[d], reinterpret_cast<float*>(rhbuf) + off, B, H, H, a_rows, gc);
        if (rc != OCR_OK) return rc;
    }
    const size_t smem = fixed + (size_t)n_stages * gc * a_bytes;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * NS * MT);
    cfg.blockDim = dim3(kGruThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident: they wait on each other twice per frame
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return f16 ? gru_launch<true>(cfg, tmW, tmH, tmR, xp, seq_len, hbuf, rhbuf, out, counters, T, B, H, NS, a_rows, n_stages, gc, MT)
               : gru_launch<false>(cfg, tmW, tmH, tmR, xp, seq_len, hbuf, rhbuf, out, counters, T, B, H, NS, a_rows, n_stages, gc, MT);
}

}  // namespace ocr
