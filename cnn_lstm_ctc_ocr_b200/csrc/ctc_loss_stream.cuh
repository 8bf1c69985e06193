// Streaming CTC loss + gradient kernel for sm_100a: the default path of ocr_ctc_loss for the shapes the reference
// trains on (labels up to 31, up to 67 classes, TMA-eligible tensors).  Replaces tf.nn.ctc_loss as called by
// ctc_loss_layer (/root/reference/src/weinman/model.py:224-229); op semantics per SURVEY.md App. A.4.
//
// Same arithmetic as ctc_loss_fast_kernel (linear-domain lattice in registers, exact power-of-two rescale, alpha and
// beta chains meeting in the middle, exactness guard + exact redo).  What changes is how long a sequence occupies
// shared memory.  The fast kernel holds the whole [T][G*C] block of y from load to store (26 KB per sequence, 2 CTAs per
// SM, 16 warps: ncu occupancy_limit_shared_mem).  Here:
//   front   the logits STREAM through a small ring of 16-frame TMA boxes: a producer warp issues the box loads and
//           stores; the compute warps turn each box into y * grad_scale in place (lane per row, the row in two register
//           halves joined by their maxima), copy the L+1 values the lattice will read -- y(blank), y(label_i) -- into a
//           compact e block ([T][Lmax+1] per sequence), and the box leaves for the gradient tensor at once.  That store
//           already IS the gradient of every class the label does not contain, and of every frame past the sequence end.
//   chains  as in the fast kernel (alpha warp and beta warp per sequence, register-resident, meeting in the middle), reading
//           the e block (conflict-free: lane i reads slot 1+i); the lattice of products overlays the ring, idle by then.
//           (Measured dead end, kept out: four sequences per chain warp, eight lanes each with two state pairs per lane and
//           the two warps in lock step over named barriers -- 110-150 instructions per step on ONE warp run at ~5 cycles
//           per instruction, 350-630 cycles per step against ~100 for 36 instructions: the chain phase of a CTA went from
//           8 k to 40 k cycles and the SM has no other work to put into the idle issue slots.)
//   fix-up  lane per frame: posterior = product / row sum, subtracted from the e row (a class that occurs twice is
//           updated twice, in label order, on the slot of its first occurrence), then the L+1 touched columns are
//           written over the y already stored: 4-byte stores into lines that are still dirty in L2, so DRAM sees
//           every gradient byte once.
// Shared memory per sequence: e block 4.3 KB + lattice 9.5 KB (cfg2) -> 4 CTAs of 4 sequences (36 warps) per SM.
#pragma once
#include "ctc_loss_fast.cuh"

namespace ocr {

constexpr int kStreamMaxBuf = 8;

struct StreamLayout {
    int RS;        // floats per frame row of a ring box (= G*C: the dense TMA box)
    int CB;        // bytes per ring box (kTmRows frames)
    int EP;        // floats per e row: [blank][label 0 .. Lmax-1]  (odd: lane-per-frame accesses are conflict-free)
    int LS;        // floats per lattice row: [trash][state 0 .. 2*Lmax][pad][trash pair][exponent]  (odd), as FastLayout
    int EX, HI;    // index of the exponent slot / of the high trash pair
    int lat_seq, e_seq;   // floats per sequence in the lattice / e block
    int ring, eb, lab, first, info, zero, total;   // byte offsets
};

__host__ __device__ inline StreamLayout stream_layout(int T, int C, int Lmax, int G, int nbuf) {
    StreamLayout f;
    f.RS = G * C;
    f.CB = kTmRows * f.RS * 4;
    f.EP = (Lmax + 1) | 1;
    f.LS = 2 * Lmax + 5;
    f.HI = 2 * Lmax + 2;
    f.EX = 2 * Lmax + 4;
    f.lat_seq = T * f.LS;
    f.e_seq = T * f.EP;
    int o = 0;
    f.ring = o;   // the ring of boxes during the front phase, the lattice afterwards
    const int lat_bytes = G * f.lat_seq * 4, ring_bytes = nbuf * f.CB;
    o += lat_bytes > ring_bytes ? lat_bytes : ring_bytes;
    o = (o + 15) & ~15;
    f.eb = o;    o += (G * f.e_seq + 2 * f.EP) * 4;   // one pad row on either side: the chains prefetch one frame past their end
    f.lab = o;   o += G * (Lmax + 1) * 4;
    f.first = o; o += G * (Lmax + 1) * 4;
    f.info = o;  o += G * 8 * 4;
    o = (o + 15) & ~15;
    f.zero = o;  o += 16;
    f.total = o;
    return f;
}

__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ float4 lds128(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(unsigned a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}

template <int G> struct StreamOcc { static constexpr int kMinBlocks = G >= 8 ? 2 : (G == 4 ? 4 : 7); };

// G sequences per CTA (2, 4, 8); 2 warps per sequence.  The last warp sits the front phase out: its lane 0 is the producer
// (TMA box loads and stores), its other lanes prefetch the successor group into L2.
template <int G>
__global__ void __launch_bounds__(64 * G, StreamOcc<G>::kMinBlocks)
ctc_loss_stream_kernel(const float* __restrict__ logits, int T, int B, int C, const int32_t* __restrict__ labels,
                       const int32_t* __restrict__ label_offsets, const int32_t* __restrict__ seq_len, int Lmax, int nbuf,
                       float* __restrict__ loss, float* __restrict__ grad, int32_t* __restrict__ status, float grad_scale,
                       const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmOut, int pf_stride,
                       int inline_redo, long long* const tl, const __grid_constant__ StreamLayout lay)
{
    constexpr int NW = 2 * G;           // compute warps
    constexpr int WPB = G / 2;          // warps per box in the front phase (lane per row: 16 frames x 2 sequences per warp)
    constexpr int CBX = NW / WPB;       // boxes worked on at once (4)
    static_assert(G >= 2, "group size");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    extern __shared__ __align__(128) unsigned char smem_f[];
    unsigned char* smem = smem_f + ((128u - (smem_u32(smem_f) & 127u)) & 127u);
    float* ring = reinterpret_cast<float*>(smem + lay.ring);
    float* s_info = reinterpret_cast<float*>(smem + lay.info);
    float* s_zero = reinterpret_cast<float*>(smem + lay.zero);
    const int RS = lay.RS, LS = lay.LS, EP = lay.EP;
    __shared__ __align__(8) unsigned long long s_full[kStreamMaxBuf], s_done[kStreamMaxBuf], s_ringfree, s_stored;
    __shared__ int s_tmax;
    __shared__ int s_redo[kFastMaxG];
    __shared__ int s_Tz[kFastMaxG], s_L[kFastMaxG];   // per sequence: frames that carry a gradient (0: none), label length

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = warp == NW - 1;   // during the front phase only
    const int s = warp >> 1, role = warp & 1;  // warp 2s = alpha (forward), 2s + 1 = beta (backward) of sequence s
    ctc_mark(tl, 0);
    const int b0 = blockIdx.x * G;
    const int blank = C - 1;
    const int b = b0 + s;
    const int NC = (T + kTmRows - 1) / kTmRows;
    const unsigned row_bytes = (unsigned)(G * C * 4);
    const unsigned a_ring = smem_u32(ring);

    // ---- producer: barriers, the first boxes (requested before anything is known about the group)
    {
        if (producer && lane == 0) {
            for (int i = 0; i < nbuf; ++i) {
                mbar_init(smem_u32(&s_full[i]), 1);
                mbar_init(smem_u32(&s_done[i]), WPB);   // the warps that share a box
            }
            mbar_init(smem_u32(&s_ringfree), 1);
            mbar_init(smem_u32(&s_stored), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            for (int q = 0; q < nbuf && q < NC; ++q) {
                // (a box that reaches past T is zero-filled in shared memory and still counts in full)
                mbar_expect_tx(smem_u32(&s_full[q]), row_bytes * (unsigned)kTmRows);
                tm_load_2d(a_ring + (unsigned)q * lay.CB, &tmIn, b0 * C, q * kTmRows, smem_u32(&s_full[q]));
            }
        }
        if (tid < kFastMaxG) s_redo[tid] = 0;
        if (warp == 0) {
            int tm = 0;
            for (int i = lane; i < G; i += 32) tm = max(tm, min(max(seq_len[b0 + i], 0), T));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tm = max(tm, __shfl_xor_sync(kFullMask, tm, o));
            if (lane == 0) {
                s_tmax = tm;
                s_zero[0] = 0.0f;
            }
        }
    }

    // ---- labels, feasibility (both warps of a pair compute the same answer)
    int* s_lab = reinterpret_cast<int*>(smem + lay.lab) + s * (Lmax + 1);
    int* s_first = reinterpret_cast<int*>(smem + lay.first) + s * (Lmax + 1);
    int L = 0, Tb = 0, bad = 0;
    {
        const int off = label_offsets[b];
        L = label_offsets[b + 1] - off;
        Tb = seq_len[b];
        if (Tb < 0 || Tb > T || L < 0 || L > Lmax) { bad = 3; Tb = 0; L = 0; }
        int need_cnt = 0, badlab = 0, mine = 0;
        for (int i = lane; i < L; i += 32) {   // L <= 31: one trip
            const int l = labels[off + i];
            if (l < 0 || l >= blank) badlab = 1;
            if (i > 0 && l == labels[off + i - 1]) need_cnt++;
            mine = (l < 0 || l >= blank) ? 0 : l;
            if (role == 0) s_lab[i] = mine;
        }
        need_cnt = warp_sum_int(need_cnt);
        badlab = __any_sync(kFullMask, badlab);
        if (!bad && badlab) bad = 3;
        if (!bad && Tb > 0 && L + need_cnt > Tb) bad = 2;
        // slot of a class's first occurrence in the label: the fix-up accumulates a repeated class there
        const unsigned same = __match_any_sync(kFullMask, lane < L ? mine : -1 - lane);
        if (role == 0 && lane < L) s_first[lane] = __ffs(same) - 1;
        if (role == 0 && lane == 0) {
            s_Tz[s] = (!bad && Tb > 0) ? Tb : 0;
            s_L[s] = L;
        }
    }
    __syncthreads();  // barriers initialised, labels and s_tmax visible
    const int tmax = s_tmax;

    if (producer) {
        // L2 prefetch for the group that will take this CTA's place on the SM (pf_stride CTAs further down the grid), issued
        // once this CTA's own first box has landed so that it is never queued ahead of it
        const long long pf_b0 = ((long long)blockIdx.x + pf_stride) * G;
        if (pf_stride > 0 && pf_b0 + G <= B && lane > 0) {
            mbar_wait_sleep(smem_u32(&s_full[0]), 0);
            if (lane - 1 < NC) tm_prefetch_2d(&tmIn, (int)pf_b0 * C, (lane - 1) * kTmRows);
            if (lane == 31) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(seq_len + pf_b0));
                const int l0 = __ldg(label_offsets + pf_b0), l1 = __ldg(label_offsets + pf_b0 + G);
                for (int o = l0; o < l1; o += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(labels + o));
            }
        }
        if (lane == 0) {
            for (int q = 0; q < NC; ++q) {
                const int buf = q % nbuf;
                mbar_wait_sleep(smem_u32(&s_done[buf]), (unsigned)(q / nbuf) & 1u);
                if (grad != nullptr) {   // (loss only: the box is simply recycled)
                    tm_store_2d(&tmOut, b0 * C, q * kTmRows, a_ring + (unsigned)buf * lay.CB);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                const int nq = q + nbuf;
                if (nq < NC) {
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the box has left shared memory
                    if (nq * kTmRows < tmax) {
                        mbar_expect_tx(smem_u32(&s_full[buf]), row_bytes * (unsigned)kTmRows);
                        tm_load_2d(a_ring + (unsigned)buf * lay.CB, &tmIn, b0 * C, nq * kTmRows, smem_u32(&s_full[buf]));
                    } else {
                        mbar_arrive(smem_u32(&s_full[buf]));   // frames past every sequence of the group: nothing to fetch, the compute warps write zeros
                    }
                }
            }
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            mbar_arrive(smem_u32(&s_ringfree));   // the lattice may overlay the ring
        }
        __syncwarp();
    }

    float* const eb_all = reinterpret_cast<float*>(smem + lay.eb) + EP;   // row 0 of sequence 0
    const unsigned a_zero = smem_u32(s_zero);
    const unsigned EPB = (unsigned)EP * 4, LSB = (unsigned)LS * 4;
    const float kinv = 1.0f / grad_scale;   // the e block holds y * grad_scale

    if (!producer) {
        // ================= front: softmax of each box in place, lane per row; e block =================
        // A warp takes 16 frames of two sequences of a box (a quarter-warp = 8 consecutive frames of one sequence: with
        // RS = 63 float4 = -1 (mod 8) the eight rows of a 128-bit access phase start in eight different 16-byte bank
        // groups).  The row passes through registers in two halves of at most 8 float4, each reduced against its own
        // maximum; the halves are joined by exp2(m_half - m) when the row is scaled.  No shuffles: a row past its
        // sequence's end simply takes the other branch.
        const int wb = warp % WPB, fp = warp / WPB;
        const int fs = 2 * wb + (lane >> 4), f = lane & 15;
        const int Tz = s_Tz[fs], Lf = s_L[fs];   // frames from Tz on: zero gradient
        const int* f_lab = reinterpret_cast<const int*>(smem + lay.lab) + fs * (Lmax + 1);
        float* eb_f = eb_all + fs * lay.e_seq;
        int c0 = (4 - ((fs * C) & 3)) & 3;   // first 16-byte aligned class of this sequence's column block
        if (c0 > C) c0 = C;
        const int nb = (C - c0) >> 2;        // aligned float4 of the row
        const int nA = (nb + 1) >> 1, nB = nb - nA;
        const int nl = C - 4 * nb;           // head + tail scalars (<= 6): with the second half, through shared memory
        const int tail0 = 4 * nb;            // scalar u sits at class u (u < c0) or tail0 + u
        const float l2e = 1.4426950408889634f;
        const unsigned a_row0 = a_ring + 4u * (unsigned)(f * RS + fs * C);
        // Slots are revisited in phase order only if a warp keeps to one slot: with fewer slots than box groups the surplus
        // groups sit the front phase out (a fresh mbarrier reports the phase of parity 1 as complete).
        const int stride = min(CBX - 1, nbuf);   // (the last group's last warp is the producer)
        for (int q = fp; q < NC && fp < stride; q += stride) {
            const int slot = q % nbuf;
            mbar_wait_sleep(smem_u32(&s_full[slot]), (unsigned)(q / nbuf) & 1u);
            if (q == 0) ctc_mark(tl, 11);
            const int t = q * kTmRows + f;
            const unsigned a_row = a_row0 + (unsigned)slot * lay.CB;
            const unsigned a_pA = a_row + 4u * c0, a_pB = a_pA + 16u * nA;
            const bool loaded = q < nbuf || q * kTmRows < tmax;   // CTA-uniform
            const bool live = loaded && t < Tz;
            if (live) {
                float4 v[8];
                // ---- first half: exp(x - mA), unnormalised, back to the box
                float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (k < nA) {
                        v[k] = lds128(a_pA + 16u * k);
                        m0 = fmaxf(m0, fmaxf(v[k].x, v[k].y));
                        m1 = fmaxf(m1, fmaxf(v[k].z, v[k].w));
                    }
                const float mlA = fmaxf(m0, m1) * l2e;
                const unsigned long long l2e2 = pk2(l2e, l2e);
                unsigned long long z2 = pk2(0.0f, 0.0f), nm2 = pk2(-mlA, -mlA);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (k < nA) {
                        float a0, a1, a2, a3;
                        up2(ffma2(pk2(v[k].x, v[k].y), l2e2, nm2), a0, a1);
                        up2(ffma2(pk2(v[k].z, v[k].w), l2e2, nm2), a2, a3);
                        v[k].x = fast_ex2(a0); v[k].y = fast_ex2(a1); v[k].z = fast_ex2(a2); v[k].w = fast_ex2(a3);
                        z2 = fadd2(z2, fadd2(pk2(v[k].x, v[k].y), pk2(v[k].z, v[k].w)));
                        sts128(a_pA + 16u * k, v[k]);
                    }
                float sA, sB, zt;
                up2(z2, sA, zt);
                sA += zt;
                // ---- second half and the scalars
                m0 = -INFINITY; m1 = -INFINITY;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (k < nB) {
                        v[k] = lds128(a_pB + 16u * k);
                        m0 = fmaxf(m0, fmaxf(v[k].x, v[k].y));
                        m1 = fmaxf(m1, fmaxf(v[k].z, v[k].w));
                    }
                for (int u = 0; u < nl; ++u) m0 = fmaxf(m0, lds(a_row + 4u * (u < c0 ? u : tail0 + u)));
                const float mlB = fmaxf(m0, m1) * l2e;
                z2 = pk2(0.0f, 0.0f); nm2 = pk2(-mlB, -mlB);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (k < nB) {
                        float a0, a1, a2, a3;
                        up2(ffma2(pk2(v[k].x, v[k].y), l2e2, nm2), a0, a1);
                        up2(ffma2(pk2(v[k].z, v[k].w), l2e2, nm2), a2, a3);
                        v[k].x = fast_ex2(a0); v[k].y = fast_ex2(a1); v[k].z = fast_ex2(a2); v[k].w = fast_ex2(a3);
                        z2 = fadd2(z2, fadd2(pk2(v[k].x, v[k].y), pk2(v[k].z, v[k].w)));
                    }
                up2(z2, sB, zt);
                sB += zt;
                for (int u = 0; u < nl; ++u) {
                    const unsigned a = a_row + 4u * (u < c0 ? u : tail0 + u);
                    const float e = fast_ex2(fmaf(lds(a), l2e, -mlB));
                    sts(a, e);
                    sB += e;
                }
                // ---- join the halves: y = e_half * 2^(ml_half - ml) * grad_scale / Z
                const float ml = fmaxf(mlA, mlB);
                const float fA = fast_ex2(mlA - ml), fB = fast_ex2(mlB - ml);
                const float gz = __fdividef(grad_scale, fmaf(sA, fA, sB * fB));
                const float gA = fA * gz, gB = fB * gz;
                const unsigned long long gB2 = pk2(gB, gB), gA2 = pk2(gA, gA);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (k < nB) {   // still in registers
                        float4 o;
                        up2(fmul2(pk2(v[k].x, v[k].y), gB2), o.x, o.y);
                        up2(fmul2(pk2(v[k].z, v[k].w), gB2), o.z, o.w);
                        sts128(a_pB + 16u * k, o);
                    }
                for (int u = 0; u < nl; ++u) {
                    const unsigned a = a_row + 4u * (u < c0 ? u : tail0 + u);
                    sts(a, lds(a) * gB);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (k < nA) {
                        const float4 e = lds128(a_pA + 16u * k);
                        float4 o;
                        up2(fmul2(pk2(e.x, e.y), gA2), o.x, o.y);
                        up2(fmul2(pk2(e.z, e.w), gA2), o.z, o.w);
                        sts128(a_pA + 16u * k, o);
                    }
                // ---- the values the lattice reads
                const unsigned a_er = smem_u32(eb_f + (size_t)t * EP);
                sts(a_er, lds(a_row + 4u * blank));
                for (int i = 0; i < Lf; ++i) sts(a_er + 4u * (1 + i), lds(a_row + 4u * f_lab[i]));
            } else {
                const float4 zero4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                for (int k = 0; k < nb; ++k) sts128(a_pA + 16u * k, zero4);
                for (int u = 0; u < nl; ++u) sts(a_row + 4u * (u < c0 ? u : tail0 + u), 0.0f);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes of the box before the TMA store reads it
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_done[slot]));
        }
        ctc_mark(tl, 1);
    }
    named_sync(10, 64 * G);   // the e block is complete (ids 1..8: pair barriers, 9: exact redo)
    ctc_mark(tl, 2);

    // ================= chains: warp 2s = alpha, warp 2s + 1 = beta of sequence s (as ctc_loss_fast_kernel, on the e block) =================
    const bool run = !bad && Tb > 0;
    float* eb_s = eb_all + s * lay.e_seq;
    float* lat = ring + s * lay.lat_seq;
    float* info = s_info + s * 8;
    int* infoi = reinterpret_cast<int*>(info);
    const int mid = (Tb + 1) >> 1;
    bool novalid = false;
    if (run) {
        constexpr int NP = 1;
        mbar_wait_sleep(smem_u32(&s_ringfree), 0);   // the last box has left the ring: the lattice may use it
        ctc_mark(tl, 3);
        const unsigned a_eb = smem_u32(eb_s), a_lat = smem_u32(lat);


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
                pe[j] = hasl ? a_eb + 4u * (1 + i) : a_zero;   // e row: [blank][label 0 .. L-1]
                pes[j] = hasl ? EPB : 0u;
                skp[j] = (i >= 1 && hasl && s_lab[i - 1] != li) ? ((j == 0) ? first : 1.0f) : 0.0f;
                pl[j] = a_lat + 4u * (i <= L ? 1 + 2 * i : HI);
                ab[j] = 0.0f; al[j] = 0.0f;
            }
            unsigned pb = a_eb;
            unsigned pex = a_lat + 4u * EX;
            Rescale rs;
            float eb_n = lds(pb), el_n[NP];
#pragma unroll
            for (int j = 0; j < NP; ++j) el_n[j] = lds(pe[j]);
            // one forward step: alpha_t from alpha_{t-1}; stored value v_t = alpha_t * 2^-E
            auto step = [&](int t, auto masked) {
                const float sck = rs.sc * kinv;  // staged rows hold y * grad_scale
                const float ebs = eb_n * sck;
                float els[NP];
#pragma unroll
                for (int j = 0; j < NP; ++j) els[j] = el_n[j] * sck;
                pb += EPB;  // prefetch e_{t+1} (row Tb is never consumed; the read stays inside the CTA's shared memory)
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
                    mloc = fmaxf(mloc, fmaxf(ab[j], al[j]));
                }
                rs.next(mloc);
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
                pb += EPB;
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
                rs.next(mloc);
            }
            int t = 1;
            store();  // mid >= 1
            // (no alive-masks here: a state that cannot reach the end any more has beta = 0, so its product is zero whatever
            // its alpha; it only takes part in the rescale maximum, and the exactness guard covers that)
#pragma unroll 2
            for (; t < mid; ++t) { step(t, std::false_type()); store(); }
            ctc_mark(tl, 4);
            pair_barrier(1 + s);  // partner has stored beta_t (and its exponents) for t >= mid
            ctc_mark(tl, 5);
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
                step(t, std::false_type());
                consume(std::true_type());
                ++t;
            }
#pragma unroll 2
            for (; t < Tb; ++t) { step(t, std::false_type()); consume(std::false_type()); }
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
            ctc_mark(tl, 6);
            pair_barrier(1 + s);  // both chains done: products complete
            novalid = !(pev > 0.0f);
            if (lane == 0) {
                const float lp = novalid ? -INFINITY : (logf(pev) + (float)rs.E * 0.6931471805599453f);
                loss[b] = -lp;
                status[b] = novalid ? kCtcRedo : 0;  // an all-zero lattice may be underflow: the exact kernel decides
                if (novalid) s_redo[s] = 1;
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
                pe[j] = hasl ? a_eb + (unsigned)(Tb - 1) * EPB + 4u * i : a_zero;   // label i-1 sits in slot i
                pes[j] = hasl ? EPB : 0u;
                skp[j] = (hasl && i < L && s_lab[i] != li) ? 1.0f : 0.0f;
                c1[j] = hasl ? 1.0f : 0.0f;
                pl[j] = a_lat + (unsigned)(Tb - 1) * LSB + 4u * (i <= L ? 2 * i : HI);
                bb[j] = 0.0f; bl[j] = 0.0f;
            }
            unsigned pb = a_eb + (unsigned)(Tb - 1) * EPB;
            unsigned pex = a_lat + (unsigned)(Tb - 1) * LSB + 4u * EX;
            Rescale rs;
            float eb_n = lds(pb), el_n[NP];  // e_{Tb-1}: consumed by the step that produces beta_{Tb-2}
#pragma unroll
            for (int j = 0; j < NP; ++j) el_n[j] = lds(pe[j]);
            // one backward step: beta_t from beta_{t+1} and y_{t+1}; stored value w_t = beta_t * 2^-E
            auto step = [&](int t, auto masked) {
                const float sck = rs.sc * kinv;  // staged rows hold y * grad_scale
                const float ebs = eb_n * sck;
                float els[NP];
#pragma unroll
                for (int j = 0; j < NP; ++j) els[j] = el_n[j] * sck;
                pb -= EPB;  // prefetch e_t for the next step (t = 0 reads the pad row; never consumed)
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
                    mloc = fmaxf(mloc, fmaxf(bb[j], bl[j]));
                }
                rs.next(mloc);
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
            rs.next(1.0f);
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
#pragma unroll 2
                for (; t >= mid; --t) { step(t, std::false_type()); store(); }
                ctc_mark(tl, 4);
                pair_barrier(1 + s);  // partner has stored alpha_t (and its exponents) for t < mid
                ctc_mark(tl, 5);
                consume_prefetch();
                // t = mid - 1 >= 0: the frame where the chains meet fixes Pt
                step(t, std::false_type());
                consume(std::true_type());
                --t;
            } else {
                pair_barrier(1 + s);  // Tb == 1: the only frame belongs to the partner's half
                consume_prefetch();
                consume(std::true_type());
                --t;
            }
#pragma unroll 2
            for (; t >= 0; --t) { step(t, std::false_type()); consume(std::false_type()); }
            if (lane == 0) infoi[6] = Pt;
            ctc_mark(tl, 6);
            pair_barrier(1 + s);
        }
        pair_barrier(1 + s);  // flags / log2 p written by the alpha warp
        ctc_mark(tl, 7);
        novalid = info[2] != 0.0f;
    } else if (role == 0 && lane == 0) {
        // TF: zero-length sequence -> loss 0, grad 0.  Infeasible / invalid -> flagged, zero outputs.
        loss[b] = 0.0f;
        status[b] = bad;
    }

    if (producer && lane == 0) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
        mbar_arrive(smem_u32(&s_stored));     // every box has reached the gradient tensor: the fix-up may write over it
    }
    __syncwarp();

    // ================= fix-up: posterior, lane per frame; the L+1 touched columns overwrite the stored y =================
    if (run && !novalid) {   // (no valid path: the stored y is the gradient, the tail decides about underflow)
        mbar_wait_sleep(smem_u32(&s_stored), 0);
        ctc_mark(tl, 8);
        const int r_lo = role == 0 ? 0 : mid, r_hi = role == 0 ? mid : Tb;
        // Rows below mid carry Pt of the beta warp (it formed those products), rows above Pt of the alpha warp.
        const float l2pe = info[3];
        const float dexp = (float)(infoi[role == 0 ? 6 : 5] - infoi[4]);
        // rounding noise of the two chains grows with T (~1e-7 per frame); anything above it is lost mass
        const float thr = 1.0e-5f + 4.0e-7f * (float)T;
        bool lost = false;
        const unsigned lb = smem_u32(s_lab), fb = smem_u32(s_first);
        for (int t0 = r_lo; t0 < r_hi; t0 += 32) {
            const int t = t0 + lane;
            if (t < r_hi) {
                const unsigned rp = smem_u32(lat + (size_t)t * LS + 1);  // rp + 4u: product at state u
                float* er = eb_s + (size_t)t * EP;
                float S = 0.0f, Bs = 0.0f;
#pragma unroll 4
                for (int i = 0; i < L; ++i) {
                    Bs += lds_pure(rp + 8u * i);
                    S += lds_pure(rp + 8u * i + 4u);
                }
                Bs += lds_pure(rp + 8u * L);
                S += Bs;
                lost = lost || !(fabsf(log2f(S) - l2pe + dexp) < thr);
                if (grad == nullptr) continue;
                const float r = (S > 0.0f) ? grad_scale / S : 0.0f;
                float* g = grad + ((size_t)t * B + b) * C;
                st_stream(g + blank, er[0] - Bs * r);
                // in label order: a class that occurs twice is updated twice, one after the other, on its first slot
#pragma unroll 4
                for (int i = 0; i < L; ++i) er[1 + ldsi_pure(fb + 4u * i)] -= lds_pure(rp + 8u * i + 4u) * r;
#pragma unroll 4
                for (int i = 0; i < L; ++i)
                    if (ldsi_pure(fb + 4u * i) == i) st_stream(g + ldsi_pure(lb + 4u * i), er[1 + i]);
            }
        }
        if (__any_sync(kFullMask, lost) && lane == 0) { status[b] = kCtcRedo; s_redo[s] = 1; }
    }
    ctc_mark(tl, 9);
    // Tail of a CTA that flagged a sequence (rare): its first 128 threads recompute the flagged sequences with the exact
    // log-domain routine on the CTA's own shared memory; their plain stores follow every store above in program order
    // (the producer lane waited for its boxes to reach global memory before it came to this barrier).
    __syncthreads();
    ctc_mark(tl, 10);
    if ((inline_redo & 1) && tid < 128)
        for (int i = 0; i < G; ++i)
            if (s_redo[i])
                ctc_general_one<true, true>(smem, b0 + i, logits, T, B, C, labels, label_offsets, seq_len, Lmax, loss, grad, status,
                                            grad_scale, nullptr);
}

}  // namespace ocr
