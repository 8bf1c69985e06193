// Persistent back-propagation through time of a bidirectional LSTM layer (sm_100a): all T frames of both directions in
// ONE cooperative launch.  Gradient of tf.nn.bidirectional_dynamic_rnn over tf.contrib.rnn.LSTMCell
// (/root/reference/src/weinman/model_bu.py:167-199) as TensorFlow's while-loop gradient computes it.
//
// The per-frame product  d h_{prev}[b, n] = sum_g dG[b, g] * W_h[n, g]  (g over the 4H gate columns) is tiny and strictly
// sequential; launched per frame (train_ops.cu) it streams W_h from L2 every frame and pays two launches per frame.
// Here the weights never leave the chip and the contraction is split over K:
//   * CTA (d, j, mt) owns direction d, hidden units [16j, 16j+16) and batch rows [128mt, +128).  Its K-slice of the
//     contraction is the 64 gate columns of ITS units; the matching slice of W_h ([H rows n] x [64 k], gate-major copy made
//     by lstm_bptt_permute) is TMA-loaded ONCE into shared memory as the B operand (128 KB at H = 512);
//   * each frame, the CTA's 128 epilogue threads (one per batch row) turn d h, d c of their units into the gate
//     gradients dG (cell backward, c-state gradient carried in registers), write them to the caller's buffer AND straight
//     into a swizzled K-major shared-memory tile: the A operand never touches global memory;
//   * tcgen05.mma.kind::tf32 (M = 128 rows, N = H in halves of <= 256, K = 64) puts the CTA's PARTIAL d h_{prev}[128, H] in
//     TMEM; thread r scatters row r, 16 columns at a time, to the slice owners: part[dst slice][src slice][row][16];
//   * the owner of a slice adds the NS partials of its 16 units in a fixed order (deterministic) at the start of the next
//     frame; the CTAs of a (direction, batch tile) meet at a per-frame grid barrier (release/acquire on a global counter;
//     cooperative launch guarantees co-residency).  Partials are double buffered by frame parity.
//   * small batches (round 2): TMEM lane = accumulator row, so with B <= 32 live rows only a quarter of the 128 epilogue
//     threads had work.  The A tile now holds R = 4 (B <= 32) or 2 (B <= 64) COPIES of every batch row -- the MMA costs the
//     same (M = 128 whatever the batch) and every copy of a row receives the same partial d h_{prev}.  Copy rho of a row
//     owns 16/R of the CTA's units for the partial sums and the cell backward (it writes its gate gradients into all R
//     copies of the row in the A tile) and the destination slices [rho*NS/R, +NS/R) for the scatter: a quarter (half) of
//     the loads, arithmetic, stores and TMEM reads per thread, nothing computed twice, sums in the same order (same bits).
//     The exchange blocks are laid out [4 column groups][128 rows][4 floats]: a warp's 16-byte accesses are contiguous.
//   * Measured dead end (kept out): fetching the NS partial blocks with one cp.async.bulk per lane of the idle warp 0 into shared
//     memory ([row][16 values] blocks, the live rows of a source one contiguous kilobyte) instead of per-thread global loads: the
//     blocks land 3750 cycles after the counter is seen -- the same as 32 loads per thread (3500; 16 loads: 2000), whether strong,
//     weak or sector-exact -- and the two-store rows of that layout slow the scatter (1050 -> 2000): frame 12080 -> 14030 cycles.
//     What the gather waits for is the memory system's turn-around on lines other SMs have just written, not the load instructions.
//   * XB (default): the partials cross L2 as bfloat16 ([2 column groups][128 rows][8 values]; with four row copies
//     [4 groups][128 rows][4 values], so that a copy's 8-byte loads use whole sectors: the gather is bound by the sectors an SM
//     pulls through L2, ~35 B/clk).  Scatter and gather are bound by
//     the bytes an SM writes and reads per frame (64 KB each way at 32 rows: ~2000 + ~3500 of 13400 cycles); bfloat16 keeps
//     float32's exponent range (gradients) and halves them.  The sums themselves stay float32.
// Sequence lengths follow dynamic_rnn: an example is touched only while s < len (processing order s = T-1 .. 0), the
// backward direction visits frame len-1-s at step s; rows past their length contribute zero gate gradients.
#include <cuda_bf16.h>

#include <type_traits>

#include "gemm_tf32.cuh"

namespace ocr {

__device__ __forceinline__ unsigned bp_pack2(float lo, float hi) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const unsigned*>(&p);
}
// WEAK loads of the partials (no L1 allocation).  They are ordered after the other CTAs' stores by the acquire of the polling thread
// (ld.acquire.gpu invalidates this SM's L1) and the CTA barrier that follows it.  Strong loads (ld.global.cg / ld.relaxed.gpu) do not
// pipeline: measured ~110 cycles per load and thread, 3600 cycles for the 32 partials of a unit against one round trip for weak loads.
template <typename T> __device__ __forceinline__ T bp_ld_weak(const void* p);
template <> __device__ __forceinline__ uint2 bp_ld_weak<uint2>(const void* p) {
    uint2 v;
    asm volatile("ld.global.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
template <> __device__ __forceinline__ uint4 bp_ld_weak<uint4>(const void* p) {
    uint4 v;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float bp_lo(unsigned v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bp_hi(unsigned v) { return __uint_as_float(v & 0xffff0000u); }

constexpr int kBpThreads = 192;
constexpr int kBpHS = 16;                 // hidden units per CTA
constexpr int kBpK = 4 * kBpHS;           // K-slice: 64 gate columns

// phase timeline of CTA 0 (shared with the forward kernel's hook, ocr_debug_lstm_timeline): 8 stamps per frame
__device__ long long* g_bptt_timeline = nullptr;
// (the pointer is read once per thread at kernel start, not at every mark)
__device__ __forceinline__ void bp_mark(long long* tl, int f, int slot) {
    if (tl != nullptr) tl[f * 8 + slot] = clock64();
}

__device__ __forceinline__ void bp_wait_counter(const unsigned* ctr, unsigned target) {
    for (unsigned it = 0; it < (1u << 27); ++it) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        if (v >= target) return;     // (relaxed polls + one fence.acq_rel.gpu at the end measured slower: the fence is a ~1000-cycle membar)
    }
    __trap();
}

template <int R, bool XB>   // R: copies of a batch row in the A tile: 1, 2 or 4 (one batch tile when R > 1); XB: partials exchanged as bfloat16
__global__ void __launch_bounds__(kBpThreads, 1)
lstm_bptt_kernel(const __grid_constant__ CUtensorMap tmW, float* act /*[T*B, 8H]: activations in, d pre-activations out*/,
                 const float* __restrict__ cs /*[T,B,2H]*/, const float* __restrict__ dout /*[T,B,2H]*/, const int32_t* __restrict__ seq_len,
                 float* part /*[2][2][MT][NS][NS][128][16]*/, unsigned* __restrict__ counters /*[2][MT]*/, int T, int B, int H, int NS, int MT)
{
    long long* const tl = blockIdx.x == 0 ? g_bptt_timeline : nullptr;
    const int NH = H > 256 ? H / 2 : H;                 // columns per MMA (N), halves = H / NH
    const int halves = H / NH;
    extern __shared__ unsigned char bp_smem_raw[];
    unsigned char* smem = bp_smem_raw + ((1024u - (g_smem_u32(bp_smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    const unsigned s_base = g_smem_u32(smem);
    const unsigned w_chunk = (unsigned)H * 128u;        // one 32-wide k-chunk of the weight slice: [H rows][128 B]
    const unsigned s_w = s_base;                        // 2 chunks
    const unsigned s_a = s_w + 2 * w_chunk;             // 2 chunks of [128 rows][128 B]
    const unsigned s_bar = s_a + 2 * 128 * 128;
    const unsigned bar_w = s_bar, bar_a = s_bar + 8, bar_acc = s_bar + 16;
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(smem + (s_bar - s_base) + 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = blockIdx.x % MT, j = (blockIdx.x / MT) % NS, d = blockIdx.x / (MT * NS);
    const int m0 = mt * kGemmBM;
    const unsigned tmem_cols = H <= 32 ? 32u : (H <= 64 ? 64u : (H <= 128 ? 128u : (H <= 256 ? 256u : 512u)));

    if (threadIdx.x == 0) {
        g_mbar_init(bar_w, 1);
        g_mbar_init(bar_a, 1);
        g_mbar_init(bar_acc, 1);
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
            // resident weight slice: rows [(d*NS + j)*H, +H) of the permuted copy [2*NS*H, 64], two 32-wide k-chunks
            g_mbar_expect_tx(bar_w, 2 * w_chunk);
            const int rb = H < 256 ? H : 256;           // rows per TMA box
            for (int c = 0; c < 2; ++c)
                for (int r0 = 0; r0 < H; r0 += rb)
                    tma_load_2d(s_w + c * w_chunk + (unsigned)r0 * 128u, &tmW, c * kGemmBK, (d * NS + j) * H + r0, bar_w);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(NH >> 3) << 17) | ((unsigned)(kGemmBM >> 4) << 24);
            g_mbar_wait(bar_w, 0);
            for (int f = 0; f + 1 < T; ++f) {            // the last step's product has no consumer
                g_mbar_wait(bar_a, f & 1);               // the epilogue threads wrote this frame's dG tile
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const unsigned long long da = umma_desc_k128(s_a + c * (128 * 128));
#pragma unroll
                    for (int kk = 0; kk < kGemmBK / 8; ++kk)
                        for (int h = 0; h < halves; ++h) {
                            const unsigned long long db = umma_desc_k128(s_w + c * w_chunk + (unsigned)(h * NH) * 128u);
                            umma_tf32(tmem_d + (unsigned)(h * NH), da + (unsigned long long)(kk * 2), db + (unsigned long long)(kk * 2), idesc,
                                      (c | kk) ? 1u : 0u);
                        }
                }
                umma_commit(bar_acc);
                bp_mark(tl, f, 3);
            }
        }
    } else {
        constexpr int U = kBpHS / R;                     // units per thread
        constexpr int per = 4 / R;                       // warps per copy
        const int q = warp & 3;
        const int rl = q * 32 + lane;                    // row of the A tile / accumulator = TMEM lane
        // R == 1: batch rows are dealt round-robin to the four lane quarters, so that with a partly filled tile the live rows
        // spread over all four epilogue warps (measured at B = 32, H = 512: 19300 -> 17050 cycles per frame).
        // R > 1: copy rho of row `slot`: units [rho*U, +U) of the CTA's 16, destination slices [rho*NS/R, +NS/R).
        const int rho = q / per;
        const int slot = R > 1 ? (q % per) * 32 + lane : rl;     // row index inside the exchange blocks
        const int r = R > 1 ? slot : m0 + lane * 4 + q;
        const int u0 = rho * U;
        const int nsr = NS / R;
        const bool in_batch = r < B;
        const int len = in_batch ? min(max(seq_len[r], 0), T) : 0;
        const size_t part_tile = (size_t)128 * kBpHS;    // floats of one [4][128][4] block
        const size_t part_pd = (size_t)MT * NS * NS * part_tile;          // per (parity, direction)
        float dc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) dc[u] = 0.f;
        for (int f = 0; f < T; ++f) {
            const int s = T - 1 - f;
            const bool live = s < len;
            const int t = d ? len - 1 - s : s;
            // ---- d h of my units: the layer's output gradient + what step s+1 sent back through W_h
            float dh[U];
#pragma unroll
            for (int u = 0; u < U; ++u) dh[u] = 0.f;
            if (live) {   // this frame's records do not depend on the other CTAs: pull them towards the SM while waiting at the barrier
                const size_t o = ((size_t)t * B + r) * 2 * H + (size_t)d * H + j * kBpHS + u0;
                const float* a = act + ((size_t)t * B + r) * 8 * H + (size_t)d * 4 * H + j * kBpHS + u0;
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) asm volatile("prefetch.global.L1 [%0];" ::"l"(a + gq * H));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(cs + o));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(dout + o));
                if (s > 0) { const int tp = d ? t + 1 : t - 1; asm volatile("prefetch.global.L1 [%0];" ::"l"(cs + ((size_t)tp * B + r) * 2 * H + (size_t)d * H + j * kBpHS + u0)); }
            }
            if (f > 0) {
                if (threadIdx.x == 64) bp_wait_counter(counters + d * MT + mt, (unsigned)NS * (unsigned)f);
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (threadIdx.x == 64) bp_mark(tl, f, 0);
                if (live && s + 1 < len) {
                    // block [4 column groups][128 rows][4 floats]: my column groups rho*U/4 .., row `slot`
                    const float4* src = reinterpret_cast<const float4*>(part + ((size_t)((f - 1) & 1) * 2 + d) * part_pd + (((size_t)mt * NS + j) * NS) * part_tile) +
                                        (size_t)(u0 / 4) * 128 + slot;
                    // 16 loads in flight per thread (each batch is one L2 round trip), added in slice order: deterministic sums
                    if constexpr (XB) {
                        // bfloat16 blocks [2 groups][128 rows][8 values]: my U units = 32 / 16 / 8 bytes of row `slot`
                        const unsigned char* srcb = reinterpret_cast<const unsigned char*>(part) + (((size_t)((f - 1) & 1) * 2 + d) * part_pd + (((size_t)mt * NS + j) * NS) * part_tile) * 2 +
                               (U == 4 ? ((size_t)(u0 >> 2) * 128 + slot) * 8                                   // [4 groups][128 rows][4 values]: my group, whole
                                       : ((size_t)(u0 >> 3) * 128 + slot) * 16 + (size_t)(u0 & 7) * 2);           // [2 groups][128 rows][8 values]
                        constexpr int NV = U == 16 ? 2 : 1;                  // 16-byte (8-byte for U = 4) loads per source slice
                        // loads in flight per thread: a batch comes back in ~2000 cycles whatever its size (measured with the
                        // timeline), so the batches are as large as the registers allow -- all 32 slices at once for U = 4
                        constexpr int JB = U == 4 ? 32 : 16 / NV;
                        using LT = typename std::conditional<U == 4, uint2, uint4>::type;
                        for (int js0 = 0; js0 < NS; js0 += JB) {
                            LT a[JB][NV];
#pragma unroll
                            for (int i = 0; i < JB; ++i)
#pragma unroll
                                for (int v = 0; v < NV; ++v)
                                    if (js0 + i < NS) a[i][v] = bp_ld_weak<LT>(srcb + (size_t)(js0 + i) * part_tile * 2 + (size_t)v * 128 * 16);
#pragma unroll
                            for (int i = 0; i < JB; ++i)
#pragma unroll
                                for (int v = 0; v < NV; ++v)
                                    if (js0 + i < NS) {
                                        dh[8 * v] += bp_lo(a[i][v].x); dh[8 * v + 1] += bp_hi(a[i][v].x); dh[8 * v + 2] += bp_lo(a[i][v].y); dh[8 * v + 3] += bp_hi(a[i][v].y);
                                        if constexpr (U > 4) {
                                            dh[8 * v + 4] += bp_lo(a[i][v].z); dh[8 * v + 5] += bp_hi(a[i][v].z); dh[8 * v + 6] += bp_lo(a[i][v].w); dh[8 * v + 7] += bp_hi(a[i][v].w);
                                        }
                                    }
                            if (js0 == 0 && threadIdx.x == 64 && dh[0] != 12345.678f) bp_mark(tl, f, 7);   // (timeline: first batch of partials summed)
                        }
                    } else {
                    constexpr int JB = 16 / (U / 4);
                    for (int js0 = 0; js0 < NS; js0 += JB) {
                        float4 a[JB][U / 4];
#pragma unroll
                        for (int i = 0; i < JB; ++i)
#pragma unroll
                            for (int v = 0; v < U / 4; ++v)
                                if (js0 + i < NS) a[i][v] = __ldcg(src + (size_t)(js0 + i) * (part_tile / 4) + v * 128);   // written by other SMs this launch: L2, never a stale L1 line
#pragma unroll
                        for (int i = 0; i < JB; ++i)
#pragma unroll
                            for (int v = 0; v < U / 4; ++v)
                                if (js0 + i < NS) { dh[4 * v] += a[i][v].x; dh[4 * v + 1] += a[i][v].y; dh[4 * v + 2] += a[i][v].z; dh[4 * v + 3] += a[i][v].w; }
                    }
                    }
                }
            }
            if (threadIdx.x == 64) bp_mark(tl, f, 1);
            // ---- cell backward for my units; gate gradients to global memory and into the A tile (k = gate*16 + unit)
            float dg[4 * U];
            if (live) {
                const size_t o = ((size_t)t * B + r) * 2 * H + (size_t)d * H + j * kBpHS + u0;
                float* a = act + ((size_t)t * B + r) * 8 * H + (size_t)d * 4 * H + j * kBpHS + u0;
                const float* cprev_p = nullptr;
                if (s > 0) { const int tp = d ? t + 1 : t - 1; cprev_p = cs + ((size_t)tp * B + r) * 2 * H + (size_t)d * H + j * kBpHS + u0; }
#pragma unroll
                for (int u = 0; u < U; u += 4) {
                    const float4 gi = *reinterpret_cast<const float4*>(a + u), gj = *reinterpret_cast<const float4*>(a + H + u);
                    const float4 gf = *reinterpret_cast<const float4*>(a + 2 * H + u), go = *reinterpret_cast<const float4*>(a + 3 * H + u);
                    const float4 cn = __ldg(reinterpret_cast<const float4*>(cs + o + u)), dO = __ldg(reinterpret_cast<const float4*>(dout + o + u));
                    const float4 cp = cprev_p ? __ldg(reinterpret_cast<const float4*>(cprev_p + u)) : make_float4(0.f, 0.f, 0.f, 0.f);
#define OCR_BP1(f_, k_)                                                           \
                    {                                                             \
                        const float tc = tanhf(cn.f_);                            \
                        const float dht = dO.f_ + dh[u + k_];                     \
                        const float d_o = dht * tc * go.f_ * (1.f - go.f_);       \
                        const float dct = dc[u + k_] + dht * go.f_ * (1.f - tc * tc); \
                        dg[u + k_] = dct * gj.f_ * gi.f_ * (1.f - gi.f_);         \
                        dg[U + u + k_] = dct * gi.f_ * (1.f - gj.f_ * gj.f_);     \
                        dg[2 * U + u + k_] = dct * cp.f_ * gf.f_ * (1.f - gf.f_); \
                        dg[3 * U + u + k_] = d_o;                                 \
                        dc[u + k_] = dct * gf.f_;                                 \
                    }
                    OCR_BP1(x, 0) OCR_BP1(y, 1) OCR_BP1(z, 2) OCR_BP1(w, 3)
#undef OCR_BP1
                }
#pragma unroll
                for (int g = 0; g < 4; ++g)
#pragma unroll
                    for (int u = 0; u < U; u += 4)
                        *reinterpret_cast<float4*>(a + g * H + u) = make_float4(dg[g * U + u], dg[g * U + u + 1], dg[g * U + u + 2], dg[g * U + u + 3]);
            } else {
#pragma unroll
                for (int k = 0; k < 4 * U; ++k) dg[k] = 0.f;
            }
            // A tile, every copy of my row (rows c*(128/R) + slot): k = gate*16 + unit -> chunk k / 32, 16-byte group
            // (k % 32) / 4 XOR (row & 7)  (128-byte swizzle, tile 1024-aligned)
            {
#pragma unroll
                for (int c = 0; c < R; ++c) {
                    const int row = R > 1 ? c * (128 / R) + slot : rl;
                    unsigned char* arow = smem + (s_a - s_base) + (size_t)row * 128;
#pragma unroll
                    for (int g = 0; g < 4; ++g)
#pragma unroll
                        for (int u = 0; u < U; u += 4) {
                            const int k = g * kBpHS + u0 + u;     // multiple of 4
                            const int grp = (k & 31) >> 2;
                            *reinterpret_cast<float4*>(arow + (k >> 5) * (128 * 128) + ((grp ^ (row & 7)) << 4)) =
                                make_float4(dg[g * U + u], dg[g * U + u + 1], dg[g * U + u + 2], dg[g * U + u + 3]);
                        }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core (async proxy) reads
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (threadIdx.x == 64) {
                unsigned long long st_;
                asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(st_) : "r"(bar_a) : "memory");
                bp_mark(tl, f, 2);
            }
            if (f == T - 1) break;                        // the last step's product has no consumer (no MMA is issued for it)
            // ---- partial d h_{prev}[row, all H units] over my K-slice: scatter 16 columns to each slice owner (my share of them)
            g_mbar_wait(bar_acc, f & 1);
            if (threadIdx.x == 64) bp_mark(tl, f, 4);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {
                float4* dst = reinterpret_cast<float4*>(part + ((size_t)(f & 1) * 2 + d) * part_pd + ((size_t)mt * NS * NS + j) * part_tile) + slot;
                for (int jd0 = rho * nsr; jd0 < (rho + 1) * nsr; jd0 += 2) {   // 32 columns (two destination slices) per TMEM load; NS / R is even
                    unsigned v[32];
                    const unsigned taddr = tmem_d + ((unsigned)(q * 32) << 16) + (unsigned)(jd0 * kBpHS);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(taddr) : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (live && XB) {
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            unsigned char* blk = reinterpret_cast<unsigned char*>(part) +
                                                 (((size_t)(f & 1) * 2 + d) * part_pd + ((size_t)mt * NS * NS + j) * part_tile + (size_t)(jd0 + i) * NS * part_tile) * 2;
                            if constexpr (R == 4) {      // [4 groups][128 rows][4 values]
                                uint2* d2 = reinterpret_cast<uint2*>(blk) + slot;
#pragma unroll
                                for (int w = 0; w < 4; ++w) {
                                    const int c = 16 * i + 4 * w;
                                    __stcg(d2 + w * 128, make_uint2(bp_pack2(__uint_as_float(v[c]), __uint_as_float(v[c + 1])), bp_pack2(__uint_as_float(v[c + 2]), __uint_as_float(v[c + 3]))));
                                }
                            } else {                     // [2 groups][128 rows][8 values]
                                uint4* d4 = reinterpret_cast<uint4*>(blk) + slot;
#pragma unroll
                                for (int w = 0; w < 2; ++w) {
                                    const int c = 16 * i + 8 * w;
                                    __stcg(d4 + w * 128, make_uint4(bp_pack2(__uint_as_float(v[c]), __uint_as_float(v[c + 1])), bp_pack2(__uint_as_float(v[c + 2]), __uint_as_float(v[c + 3])),
                                                                    bp_pack2(__uint_as_float(v[c + 4]), __uint_as_float(v[c + 5])), bp_pack2(__uint_as_float(v[c + 6]), __uint_as_float(v[c + 7]))));
                                }
                            }
                        }
                    }
                    if (live && !XB) {                     // dead rows have zero gate gradients and nobody reads their partials
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            float4* d4 = dst + (size_t)(jd0 + i) * NS * (part_tile / 4);
#pragma unroll
                            for (int w = 0; w < 4; ++w)
                                __stcg(d4 + w * 128, make_float4(__uint_as_float(v[16 * i + 4 * w]), __uint_as_float(v[16 * i + 4 * w + 1]),
                                                                 __uint_as_float(v[16 * i + 4 * w + 2]), __uint_as_float(v[16 * i + 4 * w + 3])));
                        }
                    }
                }
            }
            if (threadIdx.x == 64) bp_mark(tl, f, 5);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");   // all partial stores of this CTA are ordered before ...
            if (threadIdx.x == 64) {                          // ... this gpu-scope release that publishes the slice
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counters + d * MT + mt) : "memory");
                bp_mark(tl, f, 6);
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

// whp[(d*NS + j)*H + n][g*16 + u] = wh_rows[d*H + n][g*H + 16j + u]
__global__ void bptt_permute_kernel(const float* __restrict__ wh_rows, float* __restrict__ whp, int H, int NS)
{
    const long long total = (long long)2 * NS * H * kBpK;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(idx % kBpK);
        long long row = idx / kBpK;
        const int n = (int)(row % H); row /= H;
        const int j = (int)(row % NS);
        const int d = (int)(row / NS);
        const int g = k / kBpHS, u = k % kBpHS;
        whp[idx] = wh_rows[((size_t)d * H + n) * 4 * H + (size_t)g * H + j * kBpHS + u];
    }
}

static int g_bptt_copies = 1;   // row copies in the A tile at small batches (0: one copy, the round-1 form)
static int g_bptt_xb = 1;       // partials exchanged as bfloat16 (0: float32)
int lstm_bptt_set_copies(int on) {      // bit 0: row copies, bit 1: float32 exchange
    g_bptt_copies = (on & 1) ? 1 : 0;
    g_bptt_xb = (on & 2) ? 0 : 1;
    return OCR_OK;
}

int lstm_bptt_set_timeline(long long* buf) {
    OCR_CHECK_CUDA(cudaMemcpyToSymbol(g_bptt_timeline, &buf, sizeof(buf)));
    return OCR_OK;
}

bool lstm_bptt_supported(int T, int B, int H) {
    if (H < 32 || (H % 32) != 0 || H > 512 || T < 1 || B < 1) return false;
    if (H > 256 && ((H / 2) % 16) != 0) return false;
    const int NS = H / kBpHS, MT = (B + kGemmBM - 1) / kGemmBM;
    if (2 * NS * MT > 148) return false;                          // one CTA per SM, all co-resident
    return (size_t)2 * H * 128 + 2 * 128 * 128 + 1024 + 1024 <= (size_t)kMaxDynSmem;
}

// floats: permuted weights [2*NS*H, 64] + partial exchange [2][2][MT][NS][NS][128][16] + counters
size_t lstm_bptt_workspace_floats(int B, int H) {
    const size_t NS = H / kBpHS, MT = (B + kGemmBM - 1) / kGemmBM;
    return (size_t)2 * NS * H * kBpK + (size_t)4 * MT * NS * NS * 128 * kBpHS + 64;
}

int lstm_bptt_run(const float* dout, int T, int B, int H, const int32_t* seq_len, float* act, const float* cstate, const float* wh_rows,
                  float* ws, cudaStream_t st)
{
    const int NS = H / kBpHS, MT = (B + kGemmBM - 1) / kGemmBM;
    float* whp = ws;
    float* part = whp + (size_t)2 * NS * H * kBpK;
    unsigned* counters = reinterpret_cast<unsigned*>(part + (size_t)4 * MT * NS * NS * 128 * kBpHS);
    {
        const long long total = (long long)2 * NS * H * kBpK;
        long long g = (total + 255) / 256;
        bptt_permute_kernel<<<(int)(g > 148 * 8 ? 148 * 8 : g), 256, 0, st>>>(wh_rows, whp, H, NS);
        OCR_CHECK_LAUNCH();
    }
    OCR_CHECK_CUDA(cudaMemsetAsync(counters, 0, 64 * sizeof(float), st));
    CUtensorMap tmW;
    int rc = tma_map_2d(&tmW, whp, (long long)2 * NS * H, kBpK, kBpK, H < 256 ? H : 256);
    if (rc != OCR_OK) return rc;
    const size_t smem = (size_t)2 * H * 128 + 2 * 128 * 128 + 64 + 1024;
    static int configured = -1;
    int dev = 0;
    OCR_CHECK_CUDA(cudaGetDevice(&dev));
    if (configured != dev) {
        OCR_CHECK_CUDA(cudaFuncSetAttribute(lstm_bptt_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
        OCR_CHECK_CUDA(cudaFuncSetAttribute(lstm_bptt_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
        OCR_CHECK_CUDA(cudaFuncSetAttribute(lstm_bptt_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
        OCR_CHECK_CUDA(cudaFuncSetAttribute(lstm_bptt_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
        OCR_CHECK_CUDA(cudaFuncSetAttribute(lstm_bptt_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
        OCR_CHECK_CUDA(cudaFuncSetAttribute(lstm_bptt_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
        configured = dev;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * NS * MT);
    cfg.blockDim = dim3(kBpThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // copies of a batch row in the A tile (small batches): the largest R in {4, 2, 1} with B <= 128 / R and NS / R even
    int R = 1;
    if (g_bptt_copies && MT == 1) {
        if (B <= 32 && (NS % 8) == 0) R = 4;
        else if (B <= 64 && (NS % 4) == 0) R = 2;
    }
#define OCR_BP_GO(R_, X_) OCR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, lstm_bptt_kernel<R_, X_>, tmW, act, cstate, dout, seq_len, part, counters, T, B, H, NS, MT))
    if (g_bptt_xb) {
        if (R == 4) OCR_BP_GO(4, true); else if (R == 2) OCR_BP_GO(2, true); else OCR_BP_GO(1, true);
    } else {
        if (R == 4) OCR_BP_GO(4, false); else if (R == 2) OCR_BP_GO(2, false); else OCR_BP_GO(1, false);
    }
#undef OCR_BP_GO
    count_launch();
    return OCR_OK;
}

}  // namespace ocr
