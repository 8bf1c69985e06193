// Persistent bidirectional LSTM recurrence for sm_100a: the frame loop of tf.nn.bidirectional_dynamic_rnn over
// tf.contrib.rnn.LSTMCell (/root/reference/src/weinman/model_bu.py:167-199) as ONE kernel launch per layer.
//
// The per-frame product h_{t-1} * W_h^T is tiny (M = batch, N = 4H, K = H) and strictly sequential in t: as a
// stream of per-frame GEMM launches it is bound by launch + ramp-up latency.  Here the recurrent weights never
// leave the chip:
//   * CTA (d, j) owns direction d and hidden units [j*hs, (j+1)*hs): its 4*hs gate rows of W_h (gate-major) are
//     TMA-loaded ONCE into shared memory (128 KB for hs = 16, H = 512, TF32) as 128B-swizzled K-major tiles;
//   * every frame, the previous hidden state of the direction ([B, H], global/L2) streams through a 4-stage TMA
//     ring as the A operand; tcgen05.mma.kind::tf32 (M = 128 batch rows, N = 4*hs, K = 8) accumulates the gate
//     pre-activations of the CTA's units into TMEM;
//   * TMEM lane = batch row, so one epilogue thread holds all four gates of all hs units of its example: it adds
//     the input projection, applies the cell update with the cell state c kept in REGISTERS for the whole sequence,
//     and writes h_t (next frame's operand, double buffered) and the layer output;
//   * the CTAs of a direction meet at a per-frame grid barrier (global atomic counter; cooperative launch
//     guarantees co-residency), both directions run concurrently in the same launch.
// Per-example sequence lengths follow dynamic_rnn: an example is updated only while s < len (its state is carried
// unchanged afterwards, outputs past the length stay zero); the backward direction visits frame len-1-s at step s.
//
// Operand precision (round 2): with F16 the recurrent operands are IEEE binary16 -- h_{t-1} is published in binary16 by the
// epilogue (round to nearest; |h| < 1) and W_h is converted once per weight update.  binary16 carries the same 10 explicit
// mantissa bits the tensor core keeps of a TF32 operand (which it truncates), so the products are as exact as before,
// but a tcgen05.mma.kind::f16 covers K = 16 per instruction instead of 8: half the MMA instructions per frame (the
// frame's critical path is their issue rate, ~70 cycles each from one thread), half the h bytes through L2 -> TMA ->
// shared memory, half the resident weight bytes.  Accumulation, cell state, gate arithmetic and the layer output stay
// float32.  Shapes with H % 64 != 0, or ocr_debug_lstm_operands(0), take the TF32 instantiation.
//
// Measured dead end (round 2, kept out of the tree): replacing the counter barrier + TMA by a flagged-word exchange (8-byte words
// {two binary16 values, frame tag} written with st.relaxed.gpu and polled with 16-byte ld.relaxed.gpu by the epilogue threads,
// which then fill the swizzled operand tile themselves).  Same bits, and slower: a frame at B = 32 took 10900 cycles against
// 7240 -- a round of polls is a ~1000-1500 cycle L2 round trip under load, the 2048 16-byte groups of the h block need two
// batches per thread plus the wait for the slowest slice, and that is no shorter than release (1100) + counter poll (1900) +
// TMA (1400).  (`volatile` accesses compile to system-scope LDG/STG.STRONG.SYS: 30700 cycles per frame.)
#include <cuda_fp16.h>

#include <type_traits>

#include "gemm_tf32.cuh"

namespace ocr {

constexpr int kRnnMaxStages = 32;
constexpr int kRnnThreads = 192;

// one MUFU each (tanh.approx: max relative error 2^-11, the same order as the TF32 rounding of the operands)
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigm(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

// Optional phase timeline for tuning (ocr_debug_lstm_timeline): CTA 0 stamps globaltimer-free clock64() values, 8 per frame:
// [0] producer past the grid barrier, [1] last h tile requested, [2] first h tile landed (MMA lane), [3] last MMA issued,
// [4] accumulator complete (epilogue), [5] TMEM read, [6] cell update + stores done, [7] slice published.
__device__ long long* g_lstm_timeline = nullptr;
// (the pointer is read ONCE per thread at kernel start: read at every mark it was a dependent load on the frame's critical path)
__device__ __forceinline__ void lstm_mark(long long* tl, int s, int slot) {
    if (tl != nullptr) tl[s * 8 + slot] = clock64();
}

// bounded spin on a global counter (acquire)
__device__ __forceinline__ void wait_counter(const unsigned* ctr, unsigned target) {
    for (unsigned it = 0; it < (1u << 27); ++it) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        if (v >= target) return;     // (no nanosleep between polls: 20 ns of sleep cost ~100 cycles per frame at B = 32, ~900 at B = 256)
    }
    __trap();
}

template <int HS, bool TRAIN, bool F16>  // hidden units per CTA; N = 4*HS gate columns.  TRAIN: keep gate activations + cell states.  F16: binary16 operands
__global__ void __launch_bounds__(kRnnThreads, 1)
lstm_persistent_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH00,
                       const __grid_constant__ CUtensorMap tmH01, const __grid_constant__ CUtensorMap tmH10,
                       const __grid_constant__ CUtensorMap tmH11, const float* __restrict__ xp, const int32_t* __restrict__ seq_len,
                       void* __restrict__ hbuf_v /*[2 parity][2 dir][B][H], float32 or binary16*/, float* __restrict__ out /*[T,B,2H]*/,
                       unsigned* __restrict__ counters /*[2]*/, int T, int B, int H, int NS, int a_rows, int n_stages, int gc, int MT,
                       float* gates_out /*[T*B, 8H], may alias xp*/, float* __restrict__ cs_out /*[T,B,2H]*/)
{
    constexpr int N = 4 * HS;
    constexpr int BK = F16 ? kGemmBKh : kGemmBK;   // elements per 128-byte swizzle row
    constexpr unsigned kRowB = 128;
    const int nk = H / BK;
    long long* const tl = blockIdx.x == 0 ? g_lstm_timeline : nullptr;
    extern __shared__ unsigned char rnn_smem_raw[];
    unsigned char* smem = rnn_smem_raw + ((1024u - (g_smem_u32(rnn_smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    const unsigned s_base = g_smem_u32(smem);
    const unsigned w_bytes = (unsigned)N * kRowB;                // one k-chunk of the weight slice
    // one k-chunk of h: only the a_rows (>= B, multiple of 8) real batch rows are fetched; the MMA still reads 128
    // rows, the rest is whatever follows in shared memory and lands in TMEM lanes no thread looks at
    const unsigned a_bytes = (unsigned)a_rows * kRowB;
    const unsigned s_w = s_base;                                 // nk resident weight tiles
    // ring of n_stages groups of gc k-chunk tiles of h; ONE 3-D TMA request per group (issuing sixteen 4 KB requests one by
    // one cost the producer lane ~450 cycles each: measured with ocr_debug_lstm_timeline)
    const unsigned s_a = s_w + (unsigned)nk * w_bytes;
    const unsigned g_bytes = (unsigned)gc * a_bytes;
    const int ng = nk / gc;                                      // groups per frame
    const unsigned s_bar = s_a + (unsigned)n_stages * g_bytes + (a_rows < kGemmBM ? kGemmBM * kRowB : 0);  // + one tile of slack for the 128-row read of a short tile
    const unsigned bar_full = s_bar, bar_empty = s_bar + kRnnMaxStages * 8, bar_w = bar_empty + kRnnMaxStages * 8, bar_acc = bar_w + 8;
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(smem + (s_bar - s_base) + (2 * kRnnMaxStages + 2) * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // direction, hidden slice, batch tile (rows [mt*128, +128) of the batch)
    const int mt = blockIdx.x % MT, j = (blockIdx.x / MT) % NS, d = blockIdx.x / (MT * NS);
    const int m0 = mt * kGemmBM;
    // Two issuing threads when the whole h block is one resident group (small batches): a frame's 4*nk MMAs issue at ~70
    // cycles each from one thread although the tensor pipe needs ~34 (measured: 4560 of 10600 cycles per frame at B = 32).
    // The producer lane, idle once its TMA request is out, issues the odd k-chunks into a SECOND accumulator (TMEM columns
    // N..2N); the epilogue adds the two.
    const bool dual = n_stages == 1 && gc == nk && nk >= 2;
    const unsigned tmem_cols = (unsigned)(2 * N < 32 ? 32 : 2 * N);
    // instruction descriptor: float32 accumulator; operand formats TF32 (2) or binary16 (0); N, M
    constexpr unsigned kIdesc = (1u << 4) | ((F16 ? 0u : 2u) << 7) | ((F16 ? 0u : 2u) << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(kGemmBM >> 4) << 24);
    auto mma = [](unsigned td, unsigned long long da, unsigned long long db, unsigned acc) {
        if constexpr (F16) umma_f16(td, da, db, kIdesc, acc);
        else umma_tf32(td, da, db, kIdesc, acc);
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < n_stages; ++s) { g_mbar_init(bar_full + s * 8, 1); g_mbar_init(bar_empty + s * 8, dual ? 2 : 1); }
        g_mbar_init(bar_w, 1);
        g_mbar_init(bar_acc, dual ? 2 : 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(g_smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_d = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // resident weight slice: rows [(d*NS + j)*N, +N) of the gate-major permuted W_h
            g_mbar_expect_tx(bar_w, (unsigned)nk * w_bytes);
            for (int k = 0; k < nk; ++k) tma_load_2d(s_w + k * w_bytes, &tmW, k * BK, (d * NS + j) * N, bar_w);
            int it = 0;
            for (int s = 0; s < T; ++s) {
                // the first group's stage is armed BEFORE the grid barrier (its release by the previous frame's MMAs and the
                // expect_tx do not depend on the other CTAs): after the barrier only the fence and the request remain
                {
                    const int st0 = it % n_stages;
                    if (it >= n_stages) g_mbar_wait(bar_empty + st0 * 8, ((it / n_stages) - 1) & 1);
                    g_mbar_expect_tx(bar_full + st0 * 8, g_bytes);
                }
                if (s > 0) {
                    wait_counter(counters + d, (unsigned)(NS * MT) * (unsigned)s);   // every slice / batch tile of this direction wrote h_s
                    // generic-proxy writes -> async-proxy (TMA) reads; global state space only (measured: the all-spaces form
                    // `fence.proxy.async` cost 800 cycles per frame; fencing on the writers' side instead gains nothing)
                    asm volatile("fence.proxy.async.global;" ::: "memory");
                }
                const CUtensorMap* tm = (s & 1) ? (d ? &tmH11 : &tmH10) : (d ? &tmH01 : &tmH00);
                lstm_mark(tl, s, 0);
                for (int gi = 0; gi < ng; ++gi, ++it) {
                    const int st = it % n_stages;
                    if (gi > 0) {
                        if (it >= n_stages) g_mbar_wait(bar_empty + st * 8, ((it / n_stages) - 1) & 1);
                        g_mbar_expect_tx(bar_full + st * 8, g_bytes);
                    }
                    tma_load_3d(s_a + st * g_bytes, tm, 0, m0, gi * gc, bar_full + st * 8);
                }
                lstm_mark(tl, s, 1);
                if (dual) {   // second MMA issuer: odd k-chunks -> accumulator 1
                    if (s == 0) g_mbar_wait(bar_w, 0);
                    g_mbar_wait(bar_full, s & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int k = 1; k < nk; k += 2) {
                        const unsigned long long da = umma_desc_k128(s_a + k * a_bytes), db = umma_desc_k128(s_w + k * w_bytes);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)   // 32 bytes of the swizzle row per MMA (8 tf32 / 16 binary16)
                            mma(tmem_d + (unsigned)N, da + (unsigned long long)(kk * 2), db + (unsigned long long)(kk * 2), (k > 1 || kk) ? 1u : 0u);
                    }
                    umma_commit(bar_empty);
                    umma_commit(bar_acc);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            g_mbar_wait(bar_w, 0);
            int it = 0;
            for (int s = 0; s < T; ++s) {
                for (int gi = 0; gi < ng; ++gi, ++it) {
                    const int st = it % n_stages;
                    g_mbar_wait(bar_full + st * 8, (it / n_stages) & 1);
                    if (gi == 0) lstm_mark(tl, s, 2);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int c = 0; c < gc; c += (dual ? 2 : 1)) {   // dual: even chunks here, odd chunks on the producer lane
                        const int k = gi * gc + c;
                        const unsigned long long da = umma_desc_k128(s_a + st * g_bytes + c * a_bytes), db = umma_desc_k128(s_w + k * w_bytes);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            mma(tmem_d, da + (unsigned long long)(kk * 2), db + (unsigned long long)(kk * 2), (k | kk) ? 1u : 0u);
                    }
                    umma_commit(bar_empty + st * 8);
                }
                umma_commit(bar_acc);   // gate pre-activations of frame step s are in TMEM
                lstm_mark(tl, s, 3);
                // the next frame's first MMA overwrites TMEM; it cannot start before the epilogue has read this frame:
                // its operand h_{s+1} only exists after every CTA (this one included) passed the grid barrier
            }
        }
    } else {
        const int q = warp & 3;
        const int r = m0 + q * 32 + lane;      // batch row (TMEM lane q*32 + lane of this batch tile)
        const bool live_row = r < B;
        const int len = live_row ? min(max(seq_len[r], 0), T) : 0;
        float c[HS], h[HS];
#pragma unroll
        for (int u = 0; u < HS; ++u) { c[u] = 0.0f; h[u] = 0.0f; }
        for (int s = 0; s < T; ++s) {
            // the input projection of this frame does not depend on the recurrent product: fetch it while the MMAs run
            const bool upd = live_row && s < len;
            const int t = d ? len - 1 - s : s;
            float xi[HS], xj[HS], xf[HS], xo[HS];
            if (upd) {
                const float* x = xp + ((size_t)t * B + r) * 8 * H + (size_t)d * 4 * H + j * HS;
#pragma unroll
                for (int u = 0; u < HS; u += 4) {   // plain loads: in TRAIN mode the same thread overwrites these slots below
                    *reinterpret_cast<float4*>(xi + u) = *reinterpret_cast<const float4*>(x + u);
                    *reinterpret_cast<float4*>(xj + u) = *reinterpret_cast<const float4*>(x + H + u);
                    *reinterpret_cast<float4*>(xf + u) = *reinterpret_cast<const float4*>(x + 2 * H + u);
                    *reinterpret_cast<float4*>(xo + u) = *reinterpret_cast<const float4*>(x + 3 * H + u);
                }
            }
            g_mbar_wait(bar_acc, s & 1);
            if (threadIdx.x == 128) lstm_mark(tl, s, 4);   // (warp 4 = TMEM lanes 0..31: live at every batch size)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            unsigned g[N];
#pragma unroll
            for (int c0 = 0; c0 < N; c0 += 16) {
                const unsigned taddr = tmem_d + ((unsigned)(q * 32) << 16) + (unsigned)c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(g[c0 + 0]), "=r"(g[c0 + 1]), "=r"(g[c0 + 2]), "=r"(g[c0 + 3]), "=r"(g[c0 + 4]), "=r"(g[c0 + 5]), "=r"(g[c0 + 6]),
                      "=r"(g[c0 + 7]), "=r"(g[c0 + 8]), "=r"(g[c0 + 9]), "=r"(g[c0 + 10]), "=r"(g[c0 + 11]), "=r"(g[c0 + 12]),
                      "=r"(g[c0 + 13]), "=r"(g[c0 + 14]), "=r"(g[c0 + 15])
                    : "r"(taddr) : "memory");
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (dual) {   // add the second accumulator (odd k-chunks)
#pragma unroll
                for (int c0 = 0; c0 < N; c0 += 16) {
                    unsigned g2[16];
                    const unsigned taddr = tmem_d + ((unsigned)(q * 32) << 16) + (unsigned)(N + c0);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=r"(g2[0]), "=r"(g2[1]), "=r"(g2[2]), "=r"(g2[3]), "=r"(g2[4]), "=r"(g2[5]), "=r"(g2[6]), "=r"(g2[7]), "=r"(g2[8]),
                          "=r"(g2[9]), "=r"(g2[10]), "=r"(g2[11]), "=r"(g2[12]), "=r"(g2[13]), "=r"(g2[14]), "=r"(g2[15])
                        : "r"(taddr) : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 16; ++i) g[c0 + i] = __float_as_uint(__uint_as_float(g[c0 + i]) + __uint_as_float(g2[i]));
                }
            }
            if (threadIdx.x == 128) lstm_mark(tl, s, 5);
            if (live_row) {
                const size_t hoff = (((size_t)((s + 1) & 1) * 2 + d) * B + r) * H + j * HS;
                if (upd) {
#pragma unroll
                    for (int u = 0; u < HS; ++u) {
                        const float zi = __uint_as_float(g[u]) + xi[u];
                        const float zj = __uint_as_float(g[HS + u]) + xj[u];
                        const float zf = __uint_as_float(g[2 * HS + u]) + xf[u];
                        const float zo = __uint_as_float(g[3 * HS + u]) + xo[u];
                        const float gi = sigm(zi), gj = tanh_fast(zj), gf = sigm(zf + 1.0f), go = sigm(zo);
                        c[u] = gf * c[u] + gi * gj;
                        h[u] = go * tanh_fast(c[u]);
                        if (TRAIN) { xi[u] = gi; xj[u] = gj; xf[u] = gf; xo[u] = go; }
                    }
                }
                // carried or updated, the state is the next frame's operand
                if constexpr (F16) {
                    __half* hn = reinterpret_cast<__half*>(hbuf_v) + hoff;
#pragma unroll
                    for (int u = 0; u < HS; u += 8) {
                        const __half2 p0 = __floats2half2_rn(h[u], h[u + 1]), p1 = __floats2half2_rn(h[u + 2], h[u + 3]);
                        const __half2 p2 = __floats2half2_rn(h[u + 4], h[u + 5]), p3 = __floats2half2_rn(h[u + 6], h[u + 7]);
                        *reinterpret_cast<uint4*>(hn + u) = make_uint4(*reinterpret_cast<const unsigned*>(&p0), *reinterpret_cast<const unsigned*>(&p1),
                                                                       *reinterpret_cast<const unsigned*>(&p2), *reinterpret_cast<const unsigned*>(&p3));
                    }
                } else {
                    float* hn = reinterpret_cast<float*>(hbuf_v) + hoff;
#pragma unroll
                    for (int u = 0; u < HS; u += 4) *reinterpret_cast<float4*>(hn + u) = make_float4(h[u], h[u + 1], h[u + 2], h[u + 3]);
                }
            }
            if (threadIdx.x == 128) lstm_mark(tl, s, 6);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps: their h stores are ordered before...
            if (warp == 2 && lane == 0) {                    // ...this gpu-scope release (cumulative) that publishes the slice
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counters + d) : "memory");
                lstm_mark(tl, s, 7);
            }
            // everything only later kernels read leaves AFTER the slice is published: the release above waits for the stores
            // issued before it, and the other CTAs of the direction wait for the release
            if (upd) {
                float* o = out + ((size_t)t * B + r) * 2 * H + (size_t)d * H + j * HS;
#pragma unroll
                for (int u = 0; u < HS; u += 4) *reinterpret_cast<float4*>(o + u) = make_float4(h[u], h[u + 1], h[u + 2], h[u + 3]);
                if (TRAIN) {   // what back-propagation through time needs: gate activations (in place of the pre-activations) and c_t
                    float* ga = gates_out + ((size_t)t * B + r) * 8 * H + (size_t)d * 4 * H + j * HS;
                    float* cso = cs_out + ((size_t)t * B + r) * 2 * H + (size_t)d * H + j * HS;
#pragma unroll
                    for (int u = 0; u < HS; u += 4) {
                        *reinterpret_cast<float4*>(ga + u) = *reinterpret_cast<float4*>(xi + u);
                        *reinterpret_cast<float4*>(ga + H + u) = *reinterpret_cast<float4*>(xj + u);
                        *reinterpret_cast<float4*>(ga + 2 * H + u) = *reinterpret_cast<float4*>(xf + u);
                        *reinterpret_cast<float4*>(ga + 3 * H + u) = *reinterpret_cast<float4*>(xo + u);
                        *reinterpret_cast<float4*>(cso + u) = make_float4(c[u], c[u + 1], c[u + 2], c[u + 3]);
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
    }
}

// gate-major permutation of the recurrent weights: row ((d*NS + j)*4 + g)*HS + u  <-  wh[d*4H + g*H + j*HS + u]
template <typename TO>
__global__ void permute_wh_kernel(const float* __restrict__ wh, TO* __restrict__ whp, int H, int HS, int NS)
{
    const long long total = (long long)8 * H * H;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(idx % H);
        long long row = idx / H;
        const int u = (int)(row % HS); row /= HS;
        const int g = (int)(row % 4); row /= 4;
        const int j = (int)(row % NS);
        const int d = (int)(row / NS);
        const float v = wh[((size_t)d * 4 * H + (size_t)g * H + j * HS + u) * H + k];
        if constexpr (std::is_same<TO, __half>::value) whp[idx] = __float2half_rn(v);
        else whp[idx] = v;
    }
}

}  // namespace ocr

using namespace ocr;

constexpr int kHS = 16;

namespace ocr {

static int g_lstm_f16 = 1;   // binary16 recurrent operands where the shape allows (H % 64 == 0); 0: TF32 everywhere

int lstm_set_operands(int f16) {
    g_lstm_f16 = f16 ? 1 : 0;
    return OCR_OK;
}
static bool lstm_f16(int H) { return g_lstm_f16 && (H % kGemmBKh) == 0; }

int lstm_set_timeline(long long* buf) {
    OCR_CHECK_CUDA(cudaMemcpyToSymbol(g_lstm_timeline, &buf, sizeof(buf)));
    return OCR_OK;
}

bool lstm_persistent_supported(int T, int B, int H) {
    if ((H % kGemmBK) != 0 || (H % kHS) != 0 || B < 1) return false;
    const int NS = H / kHS, MT = (B + kGemmBM - 1) / kGemmBM;
    if (2 * NS * MT > 148) return false;   // one CTA per SM, all co-resident
    const int bk = lstm_f16(H) ? kGemmBKh : kGemmBK;
    if (H / bk > kRnnMaxStages) return false;
    const size_t w = (size_t)H / bk * (4 * kHS) * 128;
    const size_t a = (size_t)(B >= kGemmBM ? kGemmBM : (B + 7) / 8 * 8) * 128;
    // weights + at least 2 stages of h + the slack tile + barriers + alignment
    return w + 2 * a + kGemmBM * 128 + 1024 + 1024 <= (size_t)kMaxDynSmem && T >= 1;
}

size_t lstm_persistent_workspace_floats(int B, int H) {
    // permuted W_h [8H, H] + h double buffer [2][2][B][H] + counters (64 floats); the binary16 forms use half of each
    return (size_t)8 * H * H + (size_t)4 * B * H + 64;
}

// xp [T*B, 8H] (input projection + bias), wh [8H, H] (fw i,j,f,o | bw), out [T,B,2H] (pre-zeroed by this call)
// whp: [8H, H] gate-major rows, float32 or (lstm_f16(H)) binary16 -- the form lstm_persistent_run expects for this H
int lstm_permute_wh(const float* wh, int H, float* whp, cudaStream_t st)
{
    const long long total = (long long)8 * H * H;
    long long gsz = (total + 255) / 256;
    const int grid = (int)(gsz > 148 * 16 ? 148 * 16 : gsz);
    if (lstm_f16(H)) permute_wh_kernel<__half><<<grid, 256, 0, st>>>(wh, reinterpret_cast<__half*>(whp), H, kHS, H / kHS);
    else permute_wh_kernel<float><<<grid, 256, 0, st>>>(wh, whp, H, kHS, H / kHS);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

template <bool TRAIN, bool F16>
static int lstm_launch(const cudaLaunchConfig_t& cfg, const CUtensorMap& tmW, const CUtensorMap (&tmH)[2][2], const float* xp,
                       const int32_t* seq_len, void* hbuf, float* out, unsigned* counters, int T, int B, int H, int NS, int a_rows,
                       int n_stages, int gc, int MT, float* gates_out, float* cs_out)
{
    static int configured = -1;
    int dev = 0;
    OCR_CHECK_CUDA(cudaGetDevice(&dev));
    if (configured != dev) {
        OCR_CHECK_CUDA(cudaFuncSetAttribute(lstm_persistent_kernel<kHS, TRAIN, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
        configured = dev;
    }
    OCR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, lstm_persistent_kernel<kHS, TRAIN, F16>, tmW, tmH[0][0], tmH[0][1], tmH[1][0], tmH[1][1], xp, seq_len,
                                      hbuf, out, counters, T, B, H, NS, a_rows, n_stages, gc, MT, gates_out, cs_out));
    count_launch();
    return OCR_OK;
}

// wh_perm: gate-major permuted recurrent weights from lstm_permute_wh, or NULL to permute wh into the workspace now
int lstm_persistent_run(const float* xp, const float* wh, const float* wh_perm, const int32_t* seq_len, int T, int B, int H,
                        float* out, float* ws, cudaStream_t st, float* gates_out, float* cs_out)
{
    const int NS = H / kHS, MT = (B + kGemmBM - 1) / kGemmBM;
    const bool train = gates_out != nullptr;
    const bool f16 = lstm_f16(H);
    float* whp_ws = ws;
    float* hbuf = whp_ws + (size_t)8 * H * H;
    unsigned* counters = reinterpret_cast<unsigned*>(hbuf + (size_t)4 * B * H);
    const float* whp = wh_perm;
    if (whp == nullptr) {
        int rc0 = lstm_permute_wh(wh, H, whp_ws, st);
        if (rc0 != OCR_OK) return rc0;
        whp = whp_ws;
    }
    OCR_CHECK_CUDA(cudaMemsetAsync(hbuf, 0, sizeof(float) * ((size_t)4 * B * H + 64), st));
    OCR_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)T * B * 2 * H, st));
    const int bk = f16 ? kGemmBKh : kGemmBK;
    const int nk = H / bk;
    const int a_rows = B >= kGemmBM ? kGemmBM : (B + 7) / 8 * 8;
    const size_t w_bytes = (size_t)nk * (4 * kHS) * 128, a_bytes = (size_t)a_rows * 128;
    const size_t fixed = w_bytes + (a_rows < kGemmBM ? kGemmBM * 128 : 0) + 1024 + 1024;
    // h streams in groups of gc k-chunk tiles, one TMA request each: the whole row block as one group when it fits,
    // else two ring stages of the largest group (a divisor of nk) that fits twice
    const int fit = (int)(((size_t)kMaxDynSmem - fixed) / a_bytes);     // chunk tiles that fit beside the weights
    int gc = nk, n_stages = 1;
    if (fit < nk) {
        gc = 1;
        for (int c = 1; c <= nk; ++c)
            if (nk % c == 0 && 2 * c <= fit) gc = c;
        n_stages = fit / gc;
        if (n_stages > nk / gc) n_stages = nk / gc;
        if (n_stages > kRnnMaxStages) n_stages = kRnnMaxStages;
    }
    CUtensorMap tmW, tmH[2][2];
    int rc = f16 ? tma_map_2d_h(&tmW, whp, (long long)8 * H, H, H, 4 * kHS) : tma_map_2d(&tmW, whp, (long long)8 * H, H, H, 4 * kHS);
    if (rc != OCR_OK) return rc;
    for (int p = 0; p < 2; ++p)
        for (int d = 0; d < 2; ++d) {
            rc = f16 ? tma_map_chunks_h(&tmH[p][d], reinterpret_cast<__half*>(hbuf) + ((size_t)p * 2 + d) * B * H, B, H, H, a_rows, gc)
                     : tma_map_chunks(&tmH[p][d], hbuf + ((size_t)p * 2 + d) * B * H, B, H, H, a_rows, gc);
            if (rc != OCR_OK) return rc;
        }
    const size_t smem = fixed + (size_t)n_stages * gc * a_bytes;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * NS * MT);
    cfg.blockDim = dim3(kRnnThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident: they wait on each other every frame
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    float* g0 = train ? gates_out : nullptr;
    float* c0 = train ? cs_out : nullptr;
    if (train) return f16 ? lstm_launch<true, true>(cfg, tmW, tmH, xp, seq_len, hbuf, out, counters, T, B, H, NS, a_rows, n_stages, gc, MT, g0, c0)
                          : lstm_launch<true, false>(cfg, tmW, tmH, xp, seq_len, hbuf, out, counters, T, B, H, NS, a_rows, n_stages, gc, MT, g0, c0);
    return f16 ? lstm_launch<false, true>(cfg, tmW, tmH, xp, seq_len, hbuf, out, counters, T, B, H, NS, a_rows, n_stages, gc, MT, g0, c0)
               : lstm_launch<false, false>(cfg, tmW, tmH, xp, seq_len, hbuf, out, counters, T, B, H, NS, a_rows, n_stages, gc, MT, g0, c0);
}

}  // namespace ocr
