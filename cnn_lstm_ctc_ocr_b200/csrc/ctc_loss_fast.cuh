// Bandwidth-regime CTC loss + gradient kernel for sm_100a (the default path of ocr_ctc_loss).
//
// Replaces tf.nn.ctc_loss as called by the reference's ctc_loss_layer
// (/root/reference/src/weinman/model.py:224-229); op semantics per SURVEY.md App. A.4.
//
// HBM plan: every logit is read from DRAM once and every gradient element written once
// (algorithmic bytes 2*T*C*4 per sequence); nothing else touches DRAM.
//   * a CTA owns G consecutive sequences; for each frame t their logits are ONE contiguous
//     G*C*4-byte chunk of the time-major [T,B,C] tensor, fetched by the TMA engine
//     (cp.async.bulk -> shared memory, completion on an mbarrier) into a [T][G][C] staging
//     block; the finished gradient leaves the same block through cp.async.bulk stores, so
//     the LSU never issues a global access on the aligned path;
//   * two warps per sequence.  The softmax is computed "lane per frame": each lane pulls one logit
//     row out of shared memory into registers (LDS.128 from the sequence's first 16-byte aligned class on; the
//     row pitch is 4*odd words, so each quarter-warp phase covers all 32 banks), reduces max and sum(exp) there
//     and writes y * grad_scale back -- one read and one write of the staged block, which is already
//     the gradient of every class the label does not contain;
//   * the lattice runs in the LINEAR domain on y, register resident: lane i holds the state pair
//     (blank before label i, label i), one shuffle per frame carries the neighbour state,
//     the chain per frame is shuffle + 3 dependent FP ops.  Dynamic range is handled with an
//     exact power-of-two rescale every second frame (warp max by one redux.sync, applied two
//     frames late, exponent summed as an integer; struct Rescale) -- no exp/log inside the chain
//     and no rounding from the scaling;
//   * the alpha warp walks forward while the beta warp walks backward; they meet in the middle:
//     each stores its half of the lattice, then over the other half multiplies its live values
//     into what the partner stored, so shared memory holds ONE lattice of products
//     alpha_t(u)*beta_t(u) (brought back to O(1) by an exact power of two, see prod_exponent);
//   * posterior = product / (row sum of products), subtracted "lane per frame" from the label and
//     blank columns of the staged row (duplicates handled by plain sequential order);
//   * every row of products must sum to the p(z|x) the alpha chain ended with; a sequence where it
//     does not has lost probability mass to float32 underflow (states more than 2^-126 below the
//     warp-wide maximum) and is handed to the exact log-domain kernel (status kCtcRedo).
// Numerics: float32; loss rel. error ~1e-6, gradient abs. error ~1e-6 against the float64
// recursion (an order of magnitude tighter than the float32 log-domain recursion TF itself runs).
#pragma once
#include <cuda.h>
#include <type_traits>

#include "common.cuh"

namespace ocr {

constexpr int kFastMaxG = 8;
constexpr int kCtcRedo = 100;  // internal status: recompute this sequence with the exact log-domain kernel

// The exact log-domain routine of ctc_loss.cu (one sequence, 128 threads).  kInFast: called from the tail of the fast
// kernel by its first 128 threads (named barrier instead of __syncthreads) on the CTA's own, by then idle, shared memory.
template <bool kLatticeInSmem, bool kInFast>
__device__ void ctc_general_one(unsigned char* smem, const int b, const float* __restrict__ logits, int T, int B, int C,
                                const int32_t* __restrict__ labels, const int32_t* __restrict__ label_offsets,
                                const int32_t* __restrict__ seq_len, int Lmax, float* __restrict__ loss, float* __restrict__ grad,
                                int32_t* __restrict__ status, float grad_scale, float* __restrict__ workspace);

// Optional phase timeline for tuning (ocr_debug_ctc_timeline): per warp, clock64() at up to 12 phase boundaries.  The
// buffer pointer is a kernel PARAMETER: as a __device__ global it was a load every warp waited on at its first mark
// (6 % of the kernel's stall samples, and it delayed the TMA requests of every CTA).
constexpr int kCtcTimelineSlots = 12;
__device__ __forceinline__ void ctc_mark(long long* tl, int slot) {
    if (tl != nullptr && (threadIdx.x & 31) == 0)
        tl[((size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kCtcTimelineSlots + slot] = clock64();
}

struct FastLayout {
    int RS;        // floats between consecutive frames in the staging block (4 * odd)
    int LS;        // floats per lattice row: [trash][state 0 .. 2*Lmax][pad][trash pair][exponent]  (odd)
    int EX;        // index of the exponent slot inside a lattice row
    int HI;        // index of the high trash pair
    int stage, lat, lab, info, zero, mbar, total;  // byte offsets
    int lat_seq;   // floats per sequence in the lattice block
};

__host__ __device__ inline FastLayout fast_layout(int T, int C, int Lmax, int G) {
    FastLayout f;
    int rs4 = (G * C + 3) >> 2;
    if (!(rs4 & 1)) rs4 += 1;
    f.RS = rs4 * 4;
    // positions: 0 low trash, 1+u state u (u <= 2*Lmax), HI,HI+1 high trash, EX exponent
    f.LS = 2 * Lmax + 5;  // odd: pass 3 reads the lattice lane-per-frame (stride LS words) without bank conflicts
    f.HI = 2 * Lmax + 2;
    f.EX = 2 * Lmax + 4;
    f.lat_seq = T * f.LS;
    int o = (f.RS * 4 + 127) & ~127;  // one pad row (the beta chain prefetches e_{t-1} unconditionally), staging block 128-byte aligned
    f.stage = o; o += T * f.RS * 4;
    f.lat = o;   o += (G * f.lat_seq + f.LS) * 4;  // + one row: the alpha warp's look-ahead read past the last frame
    f.lab = o;   o += G * (Lmax + 1) * 4;
    f.info = o;  o += G * 8 * 4;  // per sequence: [2] no-valid flag, [3] log2(pe), [4] Ea, [5] Pt(alpha), [6] Pt(beta)
    o = (o + 15) & ~15;
    f.zero = o;  o += 16;
    f.mbar = o;  o += 16;
    f.total = o;
    return f;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// waits that may last thousands of cycles: a nanosleep between polls keeps the waiting threads out of the issue slots
// (ncu: the spinning form cost 7.5 % of all executed instructions, and so did try_wait with a suspend-time hint)
__device__ __forceinline__ void mbar_wait_sleep(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "nanosleep.u32 %2;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity), "r"(128u) : "memory");
}
__device__ __forceinline__ void bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, unsigned src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
// tensor-map forms (one request per kTmRows frames instead of one per frame)
constexpr int kTmRows = 16;
__device__ __forceinline__ void tm_load_2d(unsigned dst, const CUtensorMap* tm, int c0, int c1, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tm_store_2d(const CUtensorMap* tm, int c0, int c1, unsigned src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(src), "r"(c0), "r"(c1) : "memory");
}
// L2 prefetch of a tensor-map box / of a contiguous run: no shared-memory destination, no completion to wait for
__device__ __forceinline__ void tm_prefetch_2d(const CUtensorMap* tm, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tm), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_prefetch(const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// shared-memory accesses of the lattice chains by 32-bit shared address (volatile: program order is the schedule)
__device__ __forceinline__ float lds(unsigned a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds4(unsigned a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int ldsi(unsigned a) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
// loads of data that no thread writes during the phase that reads it (lattice products and labels in pass 3): not
// volatile, no memory clobber, so the compiler may hoist them over the read-modify-writes of the staged rows
__device__ __forceinline__ float lds_pure(unsigned a) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int ldsi_pure(unsigned a) {
    int v;
    asm("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v)); }
__device__ __forceinline__ void sts4(unsigned a, float v) { asm volatile("st.shared.f32 [%0+4], %1;" ::"r"(a), "f"(v)); }
__device__ __forceinline__ void stsi(unsigned a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v)); }

// Exact power-of-two rescaling of a lattice chain.  The stored values are v_t = value_t * 2^-E_t.  The warp-wide
// maximum of frame t-1 (one redux.sync, issued a whole frame before it is read, so it never sits on the
// dependency chain) and the factor already applied to frame t fix the factor of frame t+1:
//   c_{t+1} = x_{t-1} - c_t   (x = exponent of the maximum)   =>   x_{t+1} = growth_t + growth_{t+1}  (deadbeat).
// Factors are exact powers of two and E is an integer, so the scaling adds no rounding error.
struct Rescale {
    float sc = 1.0f;             // factor for the frame being computed (2^-c)
    int c = 0;                   // its exponent
    int E = 0;                   // sum of the exponents applied so far: value = stored * 2^E
    unsigned mb = 0x3F800000u;   // bits of the previous frame's maximum
    __device__ __forceinline__ void next(float mloc) {
        E += c;
        const int x = (int)(mb >> 23) - 127;
        c = max(-126, min(126, x - c));
        sc = __uint_as_float((unsigned)(127 - c) << 23);
        mb = __reduce_max_sync(kFullMask, __float_as_uint(mloc));
    }
    // Every-second-frame form (kCtcRescaleAlt): only the "A" frames carry a factor, the frames between them run
    // unscaled.  An A frame takes the exponent of the maximum left by the previous A frame (its redux.sync was issued
    // two frames earlier) as its factor: nothing was applied in between, so c_t = x_{t-2} and the maximum after frame t
    // is again growth_{t-1} + growth_t -- the same two-frame bound as the every-frame form, with the maximum, the
    // redux and the exponent arithmetic on half the frames.
    __device__ __forceinline__ void begin_a() {
        const int x = (int)(mb >> 23) - 127;
        c = max(-126, min(126, x));
        sc = __uint_as_float((unsigned)(127 - c) << 23);
    }
    __device__ __forceinline__ void end_a(float mloc) {
        E += c;
        mb = __reduce_max_sync(kFullMask, __float_as_uint(mloc));
    }
};
#ifndef OCR_CTC_RESCALE_ALT
#define OCR_CTC_RESCALE_ALT 1
#endif
constexpr bool kCtcRescaleAlt = OCR_CTC_RESCALE_ALT != 0;
using FrameA = std::true_type;   // chain frame that carries the rescale factor
using FrameB = std::false_type;  // chain frame without one (every second frame when kCtcRescaleAlt)

// The stored lattice halves are scaled: alpha_t = v_t * 2^Ea_t, beta_t = w_t * 2^Eb_t with max(v), max(w) ~ 1,
// but the two maxima usually sit on different states, so v*w at the states that matter can be far below the
// float32 range.  sum_u alpha_t(u) beta_t(u) = p(z|x) for EVERY t, hence sum_u v_t(u) w_t(u) = p * 2^-(Ea_t+Eb_t):
// one estimate Pt ~ log2 p (taken where the two chains meet) gives the exact power of two that brings every
// row of products back to O(1).
template <int NP>
__device__ __forceinline__ int prod_exponent(const float (&a0)[NP], const float (&a1)[NP], const float (&b0)[NP],
                                             const float (&b1)[NP]) {
    int k = -100000;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        const int e0a = (int)((__float_as_uint(a0[j]) >> 23) & 0xff), e0b = (int)((__float_as_uint(b0[j]) >> 23) & 0xff);
        const int e1a = (int)((__float_as_uint(a1[j]) >> 23) & 0xff), e1b = (int)((__float_as_uint(b1[j]) >> 23) & 0xff);
        if (e0a != 0 && e0b != 0 && e0b != 255) k = max(k, e0a + e0b - 254);
        if (e1a != 0 && e1b != 0 && e1b != 255) k = max(k, e1a + e1b - 254);
    }
    k = __reduce_max_sync(kFullMask, k);
    return k <= -100000 ? 0 : k;
}
// product (a * w) * 2^k without leaving the float32 range on the way: the factor is always applied as two halves
// (a * 2^(k/2)) * (w * 2^(k - k/2)), exact for |k| <= 240 -- no branch on the size of k inside the chain loops.
// Beyond that the row sum misses p(z|x) and the sequence goes to the exact kernel.
struct Boost {
    float f1, f2;
    __device__ __forceinline__ explicit Boost(int k) {
        k = max(-240, min(240, k));
        const int k1 = k >> 1;
        f1 = __uint_as_float((unsigned)(127 + k1) << 23);
        f2 = __uint_as_float((unsigned)(127 + k - k1) << 23);
    }
};
// "lane per frame" sweeps over one staged row of C floats, skewed by r columns (r < 4) against bank conflicts
template <typename F>
__device__ __forceinline__ void row_sweep(float* row, int C, int r, F&& f) {
    const int tail = C < 3 ? C : 3;
    float* p = row + r;
    const int n = C - tail;
    int j = 0;
    for (; j + 8 <= n; j += 8) {
#pragma unroll
        for (int q = 0; q < 8; ++q) f(p[j + q], q);
    }
    for (; j < n; ++j) f(p[j], 0);
    for (int q = 0; q < tail; ++q) {
        int k = r + n + q;
        if (k >= C) k -= C;
        f(row[k], 0);
    }
}

// f(k) for k in [0, n) with k a compile-time constant after unrolling (register-array indexing); n <= CR
template <int CR, typename F>
__device__ __forceinline__ void for_row(int n, F&& f) {
#pragma unroll
    for (int kb = 0; kb < CR; kb += 8) {
        if (kb + 8 <= n) {
#pragma unroll
            for (int q = 0; q < 8; ++q) f(kb + q);
        } else if (kb < n) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (kb + q < n) f(kb + q);
        }
    }
}

// fb(kb) for every whole step of eight classes [kb, kb+8) below n, fp(k) for the classes of the last, partial step;
// kb and k are compile-time constants after unrolling
template <int CR, typename FB, typename FP>
__device__ __forceinline__ void for_blocks(int n, FB&& fb, FP&& fp) {
#pragma unroll
    for (int kb = 0; kb < CR; kb += 8) {
        if (kb + 8 <= n) {
            fb(kb);
        } else if (kb < n) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (kb + q < n) fp(kb + q);
        }
    }
}

// two float32 per 64-bit register pair: sm_100 issues FFMA2 / FADD2 / FMUL2 on them (IEEE results per half)
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void up2(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// TL: the phase timeline of ocr_debug_ctc_timeline is compiled in (fifteen marks cost ~3 % of the kernel's instructions even when off)
template <int NP, int CR, bool TL = false>
__global__ void __launch_bounds__(CR > 64 ? 64 * 4 : 64 * kFastMaxG)
ctc_loss_fast_kernel(const float* __restrict__ logits, int T, int B, int C, const int32_t* __restrict__ labels,
                     const int32_t* __restrict__ label_offsets, const int32_t* __restrict__ seq_len, int Lmax, int G,
                     int use_bulk, float* __restrict__ loss, float* __restrict__ grad, int32_t* __restrict__ status,
                     float grad_scale, const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmOut, int pf_stride, int inline_redo,
                     long long* const tl, const __grid_constant__ FastLayout lay)
{
    // programmatic dependent launch (launch_pdl): this grid may have been scheduled while its predecessor was still running;
    // nothing here touches global memory before the predecessor's results are visible
    if (inline_redo & 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // tuning knob (ocr_debug_ctc_pdl(2))
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const bool spec_flag = (inline_redo & 4) != 0;   // request every box of the group before its lengths are known
    inline_redo &= 1;
    extern __shared__ __align__(128) unsigned char smem_f[];
    // tensor-map TMA wants 128-byte aligned shared-memory boxes: align by hand (the launch adds 128 bytes)
    // (pointer arithmetic, not an integer round trip: the compiler keeps the shared address space and emits LDS/STS, not generic LD/ST)
    unsigned char* smem = smem_f + ((128u - (smem_u32(smem_f) & 127u)) & 127u);
    // (the layout comes as a kernel parameter: computing it took ~100 instructions per warp, 1.5 % of the kernel)
    float* stage = reinterpret_cast<float*>(smem + lay.stage);
    float* s_info = reinterpret_cast<float*>(smem + lay.info);
    float* s_zero = reinterpret_cast<float*>(smem + lay.zero);
    const unsigned bar = smem_u32(smem + lay.mbar);
    const int RS = lay.RS, LS = lay.LS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = warp >> 1, role = warp & 1;  // role 0: alpha (forward), 1: beta (backward)
    if constexpr (TL) ctc_mark(tl, 0);
    const int b0 = blockIdx.x * G;
    const int nb = min(G, B - b0);
    const bool bulk = use_bulk && nb == G;
    const int blank = C - 1;
    const int b = b0 + s;
    const bool have_seq = s < nb;

    // The group that will take a finished CTA's place on this SM (pf_stride CTAs further down the grid: CTAs start in
    // index order) should find its logits, lengths and labels in L2.  Its label range is read now and used after this
    // CTA's own loads have landed.
    const long long pf_b0 = ((long long)blockIdx.x + pf_stride) * G;
    const bool pf = pf_stride > 0 && bulk && warp == 1 && pf_b0 + G <= B;
    int pf_lab0 = 0, pf_lab1 = 0;
    if (pf && lane == 31) {
        pf_lab0 = __ldg(label_offsets + pf_b0);
        pf_lab1 = __ldg(label_offsets + pf_b0 + G);
    }

    // ---- group frame count, TMA loads (warp 0: one bulk copy per frame, spread over the lanes)
    __shared__ int s_tmax;
    __shared__ int s_redo[kFastMaxG];   // sequences of this CTA whose lattice left the float32 range (redone in the tail)
    __shared__ int s_anyredo;
    if (tid < kFastMaxG) s_redo[tid] = 0;
    if (tid == 0) s_anyredo = 0;   // ordered before the first write by the CTA barrier below
    // Speculative requests (tensor-map path): all T/16 boxes go out at once instead of after the lengths have come back
    // from global memory (a dependent load of ~800 cycles at the head of every CTA).  The rows past the group's longest
    // sequence are wasted L2 -> shared-memory traffic only: the predecessor's L2 prefetch fetched them anyway.
    const bool spec = spec_flag && bulk && (use_bulk == 2 || use_bulk == 5);
    if (warp == 0) {
        if (spec && lane == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect_tx(bar, (unsigned)(G * C * 4) * (unsigned)T);
            const unsigned dst = smem_u32(stage);
            for (int q = 0; q * kTmRows < T; ++q) tm_load_2d(dst + (unsigned)(q * kTmRows) * RS * 4, &tmIn, b0 * C, q * kTmRows, bar);
        }
        int tm = 0;
        for (int i = lane; i < nb; i += 32) tm = max(tm, min(max(seq_len[b0 + i], 0), T));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tm = max(tm, __shfl_xor_sync(kFullMask, tm, o));
        if (lane == 0) {
            s_tmax = tm;
            s_zero[0] = 0.0f;
        }
        if (bulk && !spec) {
            const unsigned row_bytes = (unsigned)(G * C * 4);
            if (lane == 0) {
                mbar_init(bar, 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                // tensor-map requests deliver whole boxes of kTmRows frames: expect the rows of the last box past tm too
                const bool tmap = use_bulk == 2 || use_bulk == 5;
                if (tm > 0) mbar_expect_tx(bar, row_bytes * (unsigned)(tmap ? (tm + kTmRows - 1) / kTmRows * kTmRows : tm));
            }
            __syncwarp();
            const float* src = logits + (size_t)b0 * C;
            const unsigned dst = smem_u32(stage);
            if (use_bulk == 2 || use_bulk == 5) {
                // use_bulk 2: the logits as a 2-D tensor [T, B*C]; one request per 16 frames (box G*C floats x 16 rows lands
                // in the dense [T][G*C] staging layout).
                if (lane == 0 && tm > 0) {
                    const int nreq = (tm + kTmRows - 1) / kTmRows;
                    for (int q = 0; q < nreq; ++q) tm_load_2d(dst + (unsigned)(q * kTmRows) * RS * 4, &tmIn, b0 * C, q * kTmRows, bar);
                }
            } else {
                for (int t = lane; t < tm; t += 32)
                    bulk_load(dst + (unsigned)t * RS * 4, src + (size_t)t * B * C, row_bytes, bar);
            }
        }
    }

    // ---- labels, feasibility (both warps of a pair compute the same answer)
    int* s_lab = reinterpret_cast<int*>(smem + lay.lab) + s * (Lmax + 1);
    int off = 0, L = 0, Tb = 0, bad = 0;
    if (have_seq) {
        off = label_offsets[b];
        L = label_offsets[b + 1] - off;
        Tb = seq_len[b];
        if (Tb < 0 || Tb > T || L < 0 || L > Lmax) { bad = 3; Tb = 0; L = 0; }
    }
    int need_cnt = 0, badlab = 0;
    for (int i = lane; i < L; i += 32) {
        const int l = labels[off + i];
        if (l < 0 || l >= blank) badlab = 1;
        if (i > 0 && l == labels[off + i - 1]) need_cnt++;
        if (role == 0) s_lab[i] = (l < 0 || l >= blank) ? 0 : l;
    }
    need_cnt = warp_sum_int(need_cnt);
    badlab = __any_sync(kFullMask, badlab);
    if (!bad && badlab) bad = 3;
    if (!bad && Tb > 0 && L + need_cnt > Tb) bad = 2;
    __syncthreads();  // mbarrier init + labels visible

    const int tmax = s_tmax;
    if (!bulk) {
        // LSU path (rows the TMA engine cannot take: unaligned tensors, the ragged last group): eight loads in flight
        // per thread before the first store (one load at a time paid a DRAM round trip per frame: 50 vs 13 us at cfg2)
        const int n = nb * C, total = tmax * n, nt = (int)blockDim.x;
        for (int base = tid; base < total; base += 8 * nt) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * nt;
                if (idx < total) {
                    const int t = idx / n, j = idx - t * n;
                    v[u] = ld_stream(logits + ((size_t)t * B + b0) * C + j);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * nt;
                if (idx < total) {
                    const int t = idx / n, j = idx - t * n;
                    stage[t * RS + j] = v[u];
                }
            }
        }
        __syncthreads();
    } else if (tmax > 0 || spec) {
        mbar_wait_sleep(bar, 0);
    }
    if constexpr (TL) ctc_mark(tl, 1);
    // L2 prefetch for the successor group: its load phase becomes an L2 hit instead of a DRAM round trip under load, and
    // DRAM sees requests while this CTA computes.  Issued after this CTA's own loads landed, so never queued ahead of them.
    if (pf) {
        if (use_bulk == 2 || use_bulk == 5) {
            if (lane < (T + kTmRows - 1) / kTmRows) tm_prefetch_2d(&tmIn, (int)pf_b0 * C, lane * kTmRows);
        } else {
            for (int t = lane; t < T; t += 32) bulk_prefetch(logits + ((size_t)t * B + pf_b0) * C, (unsigned)(G * C * 4));
        }
        if (lane == 31) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(seq_len + pf_b0));
            for (int o = pf_lab0; o < pf_lab1; o += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(labels + o));
        }
    }

    float* st_s = stage + s * C;  // this sequence's column block
    float* lat = reinterpret_cast<float*>(smem + lay.lat) + s * lay.lat_seq;
    float* info = s_info + s * 8;
    int* infoi = reinterpret_cast<int*>(info);
    const bool run = have_seq && !bad && Tb > 0;
    const int mid = (Tb + 1) >> 1;
    const int skew = (C >= 8) ? (lane >> 3) : 0;  // three-sweep path (CR == 0), scalar accesses: lanes l, l+8, l+16, l+24 share a bank
    bool novalid = false;

    if (run) {
        // ================= pass 1: softmax, lane per frame; the staged row becomes y * grad_scale =================
        const int r_lo = role == 0 ? 0 : mid, r_hi = role == 0 ? mid : Tb;
        for (int t0 = r_lo; t0 < r_hi; t0 += 32) {
            const int t = t0 + lane;
            if (t < r_hi) {
                float* row = st_s + t * RS;
                if constexpr (CR > 0) {
                    // the whole row lives in registers between the one read and the one write
                    const int tail = C < 3 ? C : 3, n = C - tail;
                    // the sweep starts at the first 16-byte aligned class of this sequence's column block (row pitch and
                    // block base are multiples of 4 words: the offset depends on the sequence only) and moves four classes
                    // per LDS.128 / STS.128; eight consecutive rows at a pitch of 4*odd words cover all 32 banks, so the
                    // quarter-warp phases of a 128-bit access are conflict-free without a per-lane skew
                    const int c0 = (C >= 8) ? ((4 - ((s * C) & 3)) & 3) : 0;
                    float* p = row + c0;
                    int kt[3];
#pragma unroll
                    for (int q = 0; q < 3; ++q) { int k = c0 + n + q; kt[q] = (k >= C) ? k - C : k; }
                    // Three passes over the register-resident row, eight classes per step; the arithmetic of whole steps is
                    // packed two floats per instruction (FFMA2 / FADD2 / FMUL2: same IEEE results, half the issue slots).
                    float x[CR], xt[3];
                    float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                    for_blocks<CR>(n,
                        [&](int kb) {
                            const float4 v0 = *reinterpret_cast<const float4*>(p + kb), v1 = *reinterpret_cast<const float4*>(p + kb + 4);
                            x[kb] = v0.x; x[kb + 1] = v0.y; x[kb + 2] = v0.z; x[kb + 3] = v0.w;
                            x[kb + 4] = v1.x; x[kb + 5] = v1.y; x[kb + 6] = v1.z; x[kb + 7] = v1.w;
#pragma unroll
                            for (int q = 0; q < 4; ++q) m[q] = fmaxf(m[q], fmaxf(x[kb + q], x[kb + 4 + q]));   // FMNMX3
                        },
                        [&](int k) { x[k] = p[k]; m[k & 3] = fmaxf(m[k & 3], x[k]); });
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        xt[q] = (q < tail) ? row[kt[q]] : -INFINITY;
                        m[q] = fmaxf(m[q], xt[q]);
                    }
                    const float l2e = 1.4426950408889634f;
                    const float ml = fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])) * l2e;
                    const unsigned long long l2e2 = pk2(l2e, l2e), nml2 = pk2(-ml, -ml);
                    unsigned long long z01 = pk2(0.0f, 0.0f), z23 = z01;   // (z0, z1), (z2, z3): class k adds into z[k & 3]
                    float zs[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                    for_blocks<CR>(n,
                        [&](int kb) {
#pragma unroll
                            for (int q = 0; q < 8; q += 2) {
                                float a0, a1;
                                up2(ffma2(pk2(x[kb + q], x[kb + q + 1]), l2e2, nml2), a0, a1);
                                x[kb + q] = fast_ex2(a0);
                                x[kb + q + 1] = fast_ex2(a1);
                                if (q & 2) z23 = fadd2(z23, pk2(x[kb + q], x[kb + q + 1]));
                                else z01 = fadd2(z01, pk2(x[kb + q], x[kb + q + 1]));
                            }
                        },
                        [&](int k) {
                            x[k] = fast_ex2(fmaf(x[k], l2e, -ml));
                            zs[k & 3] += x[k];
                        });
                    float z[4];
                    up2(z01, z[0], z[1]);
                    up2(z23, z[2], z[3]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) z[q] += zs[q];
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        xt[q] = (q < tail) ? fast_ex2(fmaf(xt[q], l2e, -ml)) : 0.0f;
                        z[q] += xt[q];
                    }
                    const float gz = grad_scale / ((z[0] + z[1]) + (z[2] + z[3]));
                    const unsigned long long gz2 = pk2(gz, gz);
                    for_blocks<CR>(n,
                        [&](int kb) {
                            float o[8];
#pragma unroll
                            for (int q = 0; q < 8; q += 2) up2(fmul2(pk2(x[kb + q], x[kb + q + 1]), gz2), o[q], o[q + 1]);
                            *reinterpret_cast<float4*>(p + kb) = make_float4(o[0], o[1], o[2], o[3]);
                            *reinterpret_cast<float4*>(p + kb + 4) = make_float4(o[4], o[5], o[6], o[7]);
                        },
                        [&](int k) { p[k] = x[k] * gz; });
#pragma unroll
                    for (int q = 0; q < 3; ++q)
                        if (q < tail) row[kt[q]] = xt[q] * gz;
                } else {
                    // rows wider than the register budget: three sweeps over shared memory
                    float m0 = -INFINITY, m1 = -INFINITY;
                    row_sweep(row, C, skew, [&](float& v, int q) { if (q & 1) m1 = fmaxf(m1, v); else m0 = fmaxf(m0, v); });
                    const float ml = fmaxf(m0, m1) * 1.4426950408889634f;
                    float z0 = 0.0f, z1 = 0.0f;
                    row_sweep(row, C, skew, [&](float& v, int q) {
                        const float e = fast_ex2(fmaf(v, 1.4426950408889634f, -ml));
                        v = e;
                        if (q & 1) z1 += e; else z0 += e;
                    });
                    const float gz = grad_scale / (z0 + z1);
                    row_sweep(row, C, skew, [&](float& v, int) { v *= gz; });
                }
            }
        }
        if constexpr (TL) ctc_mark(tl, 2);
        pair_barrier(1 + s);  // both halves of the staged block now hold y * grad_scale
        if constexpr (TL) ctc_mark(tl, 3);

        // ================= pass 2: lattice chains =================
        // lattice row t: position 1+u holds state u (blank i -> 1+2i, label i -> 2+2i); positions 0 and HI,HI+1
        // absorb the stores of lanes that own no state; position EX holds the scale exponent of the stored half.
        // All chain accesses go through 32-bit shared addresses advanced by a constant per frame.
        const unsigned a_st = smem_u32(st_s), a_lat = smem_u32(lat), a_zero = smem_u32(s_zero);
        const float kinv = 1.0f / grad_scale;
        const unsigned RSB = (unsigned)RS * 4, LSB = (unsigned)LS * 4;
        const int HI = lay.HI, EX = lay.EX;
        if (role == 0) {
            // pair i = lane*NP + j : (ab = alpha(blank before label i), al = alpha(label i))
            float ab[NP], al[NP], skp[NP];
            unsigned pe[NP], pes[NP], pl[NP];  // e_t(label i) address / stride (0: constant zero), lattice slot
            const float first = lane == 0 ? 0.0f : 1.0f;  // lane 0 has no left neighbour
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const int i = lane * NP + j;
                const bool hasl = i < L;
                const int li = hasl ? s_lab[i] : 0;
                pe[j] = hasl ? a_st + 4u * li : a_zero;
                pes[j] = hasl ? RSB : 0u;
                skp[j] = (i >= 1 && hasl && s_lab[i - 1] != li) ? ((j == 0) ? first : 1.0f) : 0.0f;
                pl[j] = a_lat + 4u * (i <= L ? 1 + 2 * i : HI);
                ab[j] = 0.0f; al[j] = 0.0f;
            }
            unsigned pb = a_st + 4u * blank;
            unsigned pex = a_lat + 4u * EX;
            Rescale rs;
            float eb_n = lds(pb), el_n[NP];
#pragma unroll
            for (int j = 0; j < NP; ++j) el_n[j] = lds(pe[j]);
            // one forward step: alpha_t from alpha_{t-1}; stored value v_t = alpha_t * 2^-E
            auto step = [&](auto frame) {
                constexpr bool kA = !kCtcRescaleAlt || decltype(frame)::value;
                if constexpr (kCtcRescaleAlt && kA) rs.begin_a();
                const float sck = kA ? rs.sc * kinv : kinv;  // staged rows hold y * grad_scale
                const float ebs = eb_n * sck;
                float els[NP];
#pragma unroll
                for (int j = 0; j < NP; ++j) els[j] = el_n[j] * sck;
                pb += RSB;  // prefetch e_{t+1} (row Tb is never consumed; the read stays inside the CTA's shared memory)
                eb_n = lds(pb);
#pragma unroll
                for (int j = 0; j < NP; ++j) { pe[j] += pes[j]; el_n[j] = lds(pe[j]); }
                const float up = __shfl_up_sync(kFullMask, al[NP - 1], 1);
                float nbv[NP], nlv[NP];
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const float pv = (j == 0) ? up : al[j - 1];
                    const float t1 = al[j] + ab[j];
                    nbv[j] = ebs * ((j == 0) ? fmaf(pv, first, ab[j]) : (ab[j] + pv));
                    nlv[j] = els[j] * fmaf(pv, skp[j], t1);
                }
                float mloc = 0.0f;
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    ab[j] = nbv[j];
                    al[j] = nlv[j];
                    if constexpr (kA) mloc = fmaxf(mloc, fmaxf(ab[j], al[j]));
                }
                if constexpr (!kCtcRescaleAlt) rs.next(mloc);
                else if constexpr (kA) rs.end_a(mloc);
            };
            auto store = [&]() {
#pragma unroll
                for (int j = 0; j < NP; ++j) { sts(pl[j], ab[j]); sts4(pl[j], al[j]); pl[j] += LSB; }
                stsi(pex, rs.E);
                pex += LSB;
            };
            // t = 0: alpha_0(blank 0) = e_0(blank), alpha_0(label 0) = e_0(label 0)
            {
                const float eb0 = eb_n;
                pb += RSB;
                eb_n = lds(pb);
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const int i = lane * NP + j;
                    const float el0 = el_n[j];
                    pe[j] += pes[j];
                    el_n[j] = lds(pe[j]);
                    ab[j] = (i == 0) ? eb0 * kinv : 0.0f;
                    al[j] = (i == 0 && L > 0) ? el0 * kinv : 0.0f;
                }
                float mloc = 0.0f;
#pragma unroll
                for (int j = 0; j < NP; ++j) mloc = fmaxf(mloc, fmaxf(ab[j], al[j]));
                if constexpr (kCtcRescaleAlt) rs.end_a(mloc); else rs.next(mloc);
            }
            int t = 1;
            store();  // mid >= 1
            // (no alive-masks: a state that cannot reach the end any more has beta = 0, so its product is zero whatever its
            // alpha; it only takes part in the rescale maximum, and the exactness guard covers that.  The masked loop forms
            // doubled the code of the chains: 9 % of the kernel's stall samples were instruction-cache misses.)
            // frames in (B, A) pairs; a single frame left over runs as B (one three-frame gap between factors)
#pragma unroll 1
            for (; t + 1 < mid; t += 2) { step(FrameB()); store(); step(FrameA()); store(); }
            if (t < mid) { step(FrameB()); store(); ++t; }
            if constexpr (TL) ctc_mark(tl, 4);
            pair_barrier(1 + s);  // partner has stored beta_t (and its exponents) for t >= mid
            if constexpr (TL) ctc_mark(tl, 5);
            int Pt = 0;
            // the partner's beta_t for the frame about to be consumed is fetched one frame ahead
            float wb_n[NP], wl_n[NP];
#pragma unroll
            for (int j = 0; j < NP; ++j) { wb_n[j] = lds(pl[j]); wl_n[j] = lds4(pl[j]); }
            int ex_n = ldsi(pex);
            // `first`: the frame where the chains meet fixes Pt (peeled out of the loops: no flag inside them)
            auto consume = [&](auto first) {
                float wb[NP], wl[NP];
#pragma unroll
                for (int j = 0; j < NP; ++j) { wb[j] = wb_n[j]; wl[j] = wl_n[j]; }
                const int Es = rs.E + ex_n;
#pragma unroll
                for (int j = 0; j < NP; ++j) { wb_n[j] = lds(pl[j] + LSB); wl_n[j] = lds4(pl[j] + LSB); }
                ex_n = ldsi(pex + LSB);  // one row past the sequence on the last frame: inside the CTA's shared memory
                if constexpr (decltype(first)::value) Pt = Es + prod_exponent<NP>(ab, al, wb, wl);
                const Boost bo(Es - Pt);
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    sts(pl[j], (ab[j] * bo.f1) * (wb[j] * bo.f2));
                    sts4(pl[j], (al[j] * bo.f1) * (wl[j] * bo.f2));
                    pl[j] += LSB;
                }
                pex += LSB;
            };
            if (t < Tb) {
                step(FrameB());
                consume(std::true_type());
                ++t;
            }
#pragma unroll 1
            for (; t + 1 < Tb; t += 2) {
                step(FrameA()); consume(std::false_type());
                step(FrameB()); consume(std::false_type());
            }
            if (t < Tb) { step(FrameA()); consume(std::false_type()); }
            // p(z|x) in e-units: alpha(2L) + alpha(2L-1) at the last frame
            float up = __shfl_up_sync(kFullMask, al[NP - 1], 1);
            if (lane == 0) up = 0.0f;
            float pev = 0.0f;
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const int i = lane * NP + j;
                const float pv = (j == 0) ? up : al[j - 1];
                if (i == L) pev = ab[j] + pv;
            }
            pev = __shfl_sync(kFullMask, pev, L / NP);
            if constexpr (TL) ctc_mark(tl, 6);
            pair_barrier(1 + s);  // both chains done: products complete
            novalid = !(pev > 0.0f);
            if (lane == 0) {
                const float lp = novalid ? -INFINITY : (logf(pev) + (float)rs.E * 0.6931471805599453f);
                loss[b] = -lp;
                status[b] = novalid ? kCtcRedo : 0;  // an all-zero lattice may be underflow: the exact kernel decides
                if (novalid) { s_redo[s] = 1; s_anyredo = 1; }
                info[2] = novalid ? 1.0f : 0.0f;
                info[3] = novalid ? 0.0f : log2f(pev);
                infoi[4] = rs.E;
                infoi[5] = Pt;
            }
        } else {
            // pair i = lane*NP + j : (bl = beta(label i-1), bb = beta(blank after label i-1))
            float bb[NP], bl[NP], skp[NP], c1[NP];
            unsigned pe[NP], pes[NP], pl[NP];
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const int i = lane * NP + j;
                const bool hasl = i >= 1 && i <= L;
                const int li = hasl ? s_lab[i - 1] : 0;
                pe[j] = hasl ? a_st + (unsigned)(Tb - 1) * RSB + 4u * li : a_zero;
                pes[j] = hasl ? RSB : 0u;
                skp[j] = (hasl && i < L && s_lab[i] != li) ? 1.0f : 0.0f;
                c1[j] = hasl ? 1.0f : 0.0f;
                pl[j] = a_lat + (unsigned)(Tb - 1) * LSB + 4u * (i <= L ? 2 * i : HI);
                bb[j] = 0.0f; bl[j] = 0.0f;
            }
            unsigned pb = a_st + (unsigned)(Tb - 1) * RSB + 4u * blank;
            unsigned pex = a_lat + (unsigned)(Tb - 1) * LSB + 4u * EX;
            Rescale rs;
            float eb_n = lds(pb), el_n[NP];  // e_{Tb-1}: consumed by the step that produces beta_{Tb-2}
#pragma unroll
            for (int j = 0; j < NP; ++j) el_n[j] = lds(pe[j]);
            // one backward step: beta_t from beta_{t+1} and y_{t+1}; stored value w_t = beta_t * 2^-E
            auto step = [&](auto frame) {
                constexpr bool kA = !kCtcRescaleAlt || decltype(frame)::value;
                if constexpr (kCtcRescaleAlt && kA) rs.begin_a();
                const float sck = kA ? rs.sc * kinv : kinv;  // staged rows hold y * grad_scale
                const float ebs = eb_n * sck;
                float els[NP];
#pragma unroll
                for (int j = 0; j < NP; ++j) els[j] = el_n[j] * sck;
                pb -= RSB;  // prefetch e_t for the next step (t = 0 reads the pad row; never consumed)
                eb_n = lds(pb);
#pragma unroll
                for (int j = 0; j < NP; ++j) { pe[j] -= pes[j]; el_n[j] = lds(pe[j]); }
                float wb[NP], wl[NP];
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    wb[j] = bb[j] * ebs;
                    wl[j] = bl[j] * els[j];
                }
                float dn = __shfl_down_sync(kFullMask, wl[0], 1);
                if (lane == 31) dn = 0.0f;
                float mloc = 0.0f;
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const float nx = (j == NP - 1) ? dn : wl[j + 1];
                    const float nbb = wb[j] + nx;
                    const float nbl = fmaf(skp[j], nx, fmaf(c1[j], wb[j], wl[j]));
                    bb[j] = nbb;
                    bl[j] = nbl;
                    if constexpr (kA) mloc = fmaxf(mloc, fmaxf(bb[j], bl[j]));
                }
                if constexpr (!kCtcRescaleAlt) rs.next(mloc);
                else if constexpr (kA) rs.end_a(mloc);
            };
            auto store = [&]() {
#pragma unroll
                for (int j = 0; j < NP; ++j) { sts(pl[j], bl[j]); sts4(pl[j], bb[j]); pl[j] -= LSB; }
                stsi(pex, rs.E);
                pex -= LSB;
            };
            // t = Tb-1: beta(last blank) = beta(last label) = 1
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const int i = lane * NP + j;
                bb[j] = (i == L) ? 1.0f : 0.0f;
                bl[j] = (i == L && L >= 1) ? 1.0f : 0.0f;
            }
            if constexpr (kCtcRescaleAlt) rs.end_a(1.0f); else rs.next(1.0f);
            int Pt = 0;
            float vl_n[NP], vb_n[NP];
            int ex_n = 0;
            auto consume_prefetch = [&]() {  // the partner's alpha_t of the first frame to be consumed
#pragma unroll
                for (int j = 0; j < NP; ++j) { vl_n[j] = lds(pl[j]); vb_n[j] = lds4(pl[j]); }
                ex_n = ldsi(pex);
            };
            auto consume = [&](auto first) {
                float vl[NP], vb[NP];
#pragma unroll
                for (int j = 0; j < NP; ++j) { vl[j] = vl_n[j]; vb[j] = vb_n[j]; }
                const int Es = rs.E + ex_n;
#pragma unroll
                for (int j = 0; j < NP; ++j) { vl_n[j] = lds(pl[j] - LSB); vb_n[j] = lds4(pl[j] - LSB); }
                ex_n = ldsi(pex - LSB);  // one row before the sequence on the last frame: inside the CTA's shared memory
                if constexpr (decltype(first)::value) Pt = Es + prod_exponent<NP>(bb, bl, vb, vl);
                const Boost bo(Es - Pt);
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    sts(pl[j], (bl[j] * bo.f1) * (vl[j] * bo.f2));
                    sts4(pl[j], (bb[j] * bo.f1) * (vb[j] * bo.f2));
                    pl[j] -= LSB;
                }
                pex -= LSB;
            };
            int t = Tb - 1;
            if (t >= mid) {
                store();
                --t;
#pragma unroll 1
                for (; t - 1 >= mid; t -= 2) { step(FrameB()); store(); step(FrameA()); store(); }
                if (t >= mid) { step(FrameB()); store(); --t; }
                if constexpr (TL) ctc_mark(tl, 4);
                pair_barrier(1 + s);  // partner has stored alpha_t (and its exponents) for t < mid
                if constexpr (TL) ctc_mark(tl, 5);
                consume_prefetch();
                // t = mid - 1 >= 0: the frame where the chains meet fixes Pt
                step(FrameB());
                consume(std::true_type());
                --t;
            } else {
                pair_barrier(1 + s);  // Tb == 1: the only frame belongs to the partner's half
                consume_prefetch();
                consume(std::true_type());
                --t;
            }
#pragma unroll 1
            for (; t - 1 >= 0; t -= 2) {
                step(FrameA()); consume(std::false_type());
                step(FrameB()); consume(std::false_type());
            }
            if (t >= 0) { step(FrameA()); consume(std::false_type()); }
            if (lane == 0) infoi[6] = Pt;
            if constexpr (TL) ctc_mark(tl, 6);
            pair_barrier(1 + s);
        }
        pair_barrier(1 + s);  // flags / log2 p written by the alpha warp
        if constexpr (TL) ctc_mark(tl, 7);
        novalid = info[2] != 0.0f;
    } else if (have_seq && role == 0 && lane == 0) {
        // TF: zero-length sequence -> loss 0, grad 0.  Infeasible / invalid -> flagged, zero outputs.
        loss[b] = 0.0f;
        status[b] = bad;
    }

    // ================= pass 3: posterior scatter + gradient rows, lane per frame =================
    if (have_seq) {
        const int Tr = run ? Tb : 0, midr = run ? mid : 0;
        const int r_lo = role == 0 ? 0 : midr, r_hi = role == 0 ? midr : Tr;
        // Rows below mid carry Pt of the beta warp (it formed those products), rows above Pt of the alpha warp.
        const float l2pe = info[3];
        const float dexp = (float)(infoi[role == 0 ? 6 : 5] - infoi[4]);
        // rounding noise of the two chains grows with T (~1e-7 per frame); anything above it is lost mass
        const float thr = 1.0e-5f + 4.0e-7f * (float)T;
        bool lost = false;
        for (int t0 = r_lo; t0 < r_hi; t0 += 32) {
            const int t = t0 + lane;
            if (t < r_hi) {
                float* row = st_s + t * RS;
                if (!novalid) {
                    const unsigned rp = smem_u32(lat + (size_t)t * LS + 1);  // rp + 4u: product at state u
                    const unsigned lb = smem_u32(s_lab);
                    float S = 0.0f, Bs = 0.0f;
#pragma unroll 4
                    for (int i = 0; i < L; ++i) {
                        Bs += lds_pure(rp + 8u * i);
                        S += lds_pure(rp + 8u * i + 4u);
                    }
                    Bs += lds_pure(rp + 8u * L);
                    S += Bs;
                    lost = lost || !(fabsf(__log2f(S) - l2pe + dexp) < thr);   // (MUFU.LG2: 2^-22 absolute, far inside thr)
                    if (grad) {
                        const float r = (S > 0.0f) ? __fdividef(grad_scale, S) : 0.0f;
                        row[blank] -= Bs * r;
                        // in label order: a class that occurs twice is updated twice, one after the other
#pragma unroll 4
                        for (int i = 0; i < L; ++i) row[ldsi_pure(lb + 4u * i)] -= lds_pure(rp + 8u * i + 4u) * r;
                    }
                }
            }
        }
        if (run && !novalid && __any_sync(kFullMask, lost) && lane == 0) { status[b] = kCtcRedo; s_redo[s] = 1; s_anyredo = 1; }
    }
    if constexpr (TL) ctc_mark(tl, 8);
    // Tail of a CTA that flagged a sequence (rare): once the CTA's own gradient block has left shared memory, its first
    // 128 threads recompute the flagged sequences with the exact log-domain routine, in place of a second kernel launch
    // that would have to visit every sequence's flag (inline_redo: the routine's layout fits this CTA's allocation).
    // Called by the whole CTA after a barrier that follows pass 3 and the CTA's own stores.
    bool do_redo = false;   // CTA-uniform; set after a CTA barrier that follows pass 3 and the CTA's stores
    if (!grad) {
        if (inline_redo) {
            __syncthreads();
            do_redo = true;
        }
    } else {
        if (have_seq) {
            // frames past the sequence end: zero gradient (both warps, lane per class: conflict-free)
            const int Tr = run ? Tb : 0;
            float* p = st_s + (size_t)(Tr + role) * RS + lane;
            if (CR > 0) {
#pragma unroll 4
                for (int t = Tr + role; t < T; t += 2, p += 2 * RS) {
#pragma unroll
                    for (int k = 0; k < CR; k += 32)
                        if (k + lane < C) p[k] = 0.0f;
                }
            } else {
                for (int t = Tr + role; t < T; t += 2, p += 2 * RS)
                    for (int k = 0; k + lane < C; k += 32) p[k] = 0.0f;
            }
        }
        if constexpr (TL) ctc_mark(tl, 9);
        if (bulk) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if constexpr (TL) ctc_mark(tl, 10);
            const bool redo = inline_redo && s_anyredo != 0;   // CTA-uniform
            if (warp == 0) {
                const unsigned row_bytes = (unsigned)(G * C * 4);
                float* dst = grad + (size_t)b0 * C;
                const unsigned src = smem_u32(stage);
                if (use_bulk == 2 || use_bulk == 6) {
                    if (lane == 0)
                        for (int q = 0; q * kTmRows < T; ++q) tm_store_2d(&tmOut, b0 * C, q * kTmRows, src + (unsigned)(q * kTmRows) * RS * 4);
                } else {
                    for (int t = lane; t < T; t += 32) bulk_store(dst + (size_t)t * B * C, src + (unsigned)t * RS * 4, row_bytes);
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (redo) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the block has reached global memory: the redo overwrites part of it
                else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            if constexpr (TL) ctc_mark(tl, 11);
            if (redo) {
                __syncthreads();
                do_redo = true;
            }
        } else {
            __syncthreads();
            const int n = nb * C;
            for (int t = 0; t < T; ++t) {
                float* dst = grad + ((size_t)t * B + b0) * C;
                for (int j = tid; j < n; j += blockDim.x) st_stream(dst + j, stage[t * RS + j]);
            }
            if (inline_redo) {
                __syncthreads();
                do_redo = true;
            }
        }
    }
    // (one call site: the routine is inlined once)
    if (do_redo && tid < 128)
        for (int i = 0; i < nb; ++i)
            if (s_redo[i])
                ctc_general_one<true, true>(smem, b0 + i, logits, T, B, C, labels, label_offsets, seq_len, Lmax, loss, grad, status,
                                            grad_scale, nullptr);
}

}  // namespace ocr
