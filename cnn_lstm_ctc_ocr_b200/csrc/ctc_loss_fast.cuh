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
//   * two warps per sequence.  Softmax statistics are computed "lane per frame" (each lane
//     walks one logit row in shared memory, skewed so banks never collide), leaving
//     e = exp(x - max) in place and Z = sum(e);
//   * the lattice runs in the LINEAR domain on the un-normalised e (the recursion is linear, the
//     row factors Z cancel in the posterior), register resident: lane i holds the state pair
//     (blank before label i, label i), one shuffle per frame carries the neighbour state,
//     the chain per frame is shuffle + 3 dependent FP ops.  Dynamic range is handled with an
//     exact power-of-two rescale every frame (warp max by one redux.sync, applied one frame
//     late, exponent summed as an integer) -- no exp/log inside the chain and no rounding from
//     the scaling;
//   * the alpha warp walks forward while the beta warp walks backward; they meet in the middle:
//     each stores its half of the lattice, then over the other half multiplies its live values
//     into what the partner stored, so shared memory holds ONE lattice of products
//     alpha_t(u)*beta_t(u);
//   * posterior = product / (row sum of products) (each row normalised by its own total, which
//     equals p(z|x) up to scaling), scattered "lane per frame" into the staged row, duplicates
//     handled by plain sequential order, then grad = (e - Z*occ) * (grad_scale / Z).
// Numerics: float32; loss rel. error ~1e-6, gradient abs. error ~1e-6 against the float64
// recursion (an order of magnitude tighter than the float32 log-domain recursion TF itself runs).
#pragma once
#include "common.cuh"

namespace ocr {

constexpr int kFastMaxG = 8;
constexpr int kCtcRedo = 100;  // internal status: recompute this sequence with the exact log-domain kernel

struct FastLayout {
    int RS;        // floats between consecutive frames in the staging block
    int Lp;        // lattice row length (Lmax + 1)
    int stage, lat, zz, exr, lab, slz, mbar, total;  // byte offsets
    int lat_seq;   // floats per sequence in the lattice block
};

__host__ __device__ inline FastLayout fast_layout(int T, int C, int Lmax, int G) {
    FastLayout f;
    f.RS = (G * C + 3) & ~3;
    f.Lp = Lmax + 1;
    f.lat_seq = T * 2 * f.Lp;
    int o = 0;
    f.stage = o; o += T * f.RS * 4;
    f.lat = o;   o += G * f.lat_seq * 4;
    f.zz = o;    o += G * T * 4;
    f.exr = o;   o += G * T * 4;
    f.lab = o;   o += G * f.Lp * 4;
    f.slz = o;   o += G * 8 * 4;  // per sequence: sum(log Z) x2, no-valid flag, log2(pe), Ea, Pt(alpha), Pt(beta), spare
    o = (o + 15) & ~15;
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
__device__ __forceinline__ void bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, unsigned src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// exact power-of-two rescale: returns the factor 2^-ex for the warp-wide maximum m (>= 0) and ex
__device__ __forceinline__ float pow2_rescale(float mloc, int& ex) {
    const unsigned mb = __reduce_max_sync(kFullMask, __float_as_uint(mloc));
    const int e = (int)(mb >> 23);
    if (e == 0 || e == 255) { ex = 0; return 1.0f; }  // zero / denormal / inf: leave alone
    ex = e - 127;
    return __uint_as_float((unsigned)(254 - e) << 23);
}

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
        const int e0a = (int)(__float_as_uint(a0[j]) >> 23), e0b = (int)(__float_as_uint(b0[j]) >> 23);
        const int e1a = (int)(__float_as_uint(a1[j]) >> 23), e1b = (int)(__float_as_uint(b1[j]) >> 23);
        if (e0a != 0 && e0b != 0) k = max(k, e0a + e0b - 254);
        if (e1a != 0 && e1b != 0) k = max(k, e1a + e1b - 254);
    }
    k = __reduce_max_sync(kFullMask, k);
    return k <= -100000 ? 0 : k;
}
// 2^k split into two representable factors (|k| clamped to 240)
__device__ __forceinline__ void boost_factors(int k, float& f1, float& f2) {
    k = max(-240, min(240, k));
    const int k1 = k >> 1, k2 = k - k1;
    f1 = __uint_as_float((unsigned)(127 + k1) << 23);
    f2 = __uint_as_float((unsigned)(127 + k2) << 23);
}

template <int NP>
__global__ void __launch_bounds__(64 * kFastMaxG)
ctc_loss_fast_kernel(const float* __restrict__ logits, int T, int B, int C, const int32_t* __restrict__ labels,
                     const int32_t* __restrict__ label_offsets, const int32_t* __restrict__ seq_len, int Lmax, int G,
                     int use_bulk, float* __restrict__ loss, float* __restrict__ grad, int32_t* __restrict__ status,
                     float grad_scale)
{
    extern __shared__ __align__(128) unsigned char smem_f[];
    unsigned char* smem = smem_f;
    const FastLayout lay = fast_layout(T, C, Lmax, G);
    float* stage = reinterpret_cast<float*>(smem + lay.stage);
    float* s_zz = reinterpret_cast<float*>(smem + lay.zz);
    float* s_slz = reinterpret_cast<float*>(smem + lay.slz);
    const unsigned bar = smem_u32(smem + lay.mbar);
    const int RS = lay.RS, Lp = lay.Lp;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = warp >> 1, role = warp & 1;  // role 0: alpha (forward), 1: beta (backward)
    const int b0 = blockIdx.x * G;
    const int nb = min(G, B - b0);
    const bool bulk = use_bulk && nb == G;
    const int blank = C - 1;
    const int b = b0 + s;
    const bool have_seq = s < nb;

    // ---- group frame count, loads
    __shared__ int s_tmax;
    if (tid == 0) {
        int tm = 0;
        for (int i = 0; i < nb; ++i) tm = max(tm, min(max(seq_len[b0 + i], 0), T));
        s_tmax = tm;
        if (bulk) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            if (tm > 0) {
                const unsigned row_bytes = (unsigned)(G * C * 4);
                mbar_expect_tx(bar, row_bytes * (unsigned)tm);
                const float* src = logits + (size_t)b0 * C;
                const unsigned dst = smem_u32(stage);
                for (int t = 0; t < tm; ++t)
                    bulk_load(dst + (unsigned)t * RS * 4, src + (size_t)t * B * C, row_bytes, bar);
            }
        }
    }

    // ---- labels, feasibility (both warps of a pair compute the same answer)
    int* s_lab = reinterpret_cast<int*>(smem + lay.lab) + s * Lp;
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
        const int n = nb * C;
        for (int t = 0; t < tmax; ++t) {
            const float* src = logits + ((size_t)t * B + b0) * C;
            for (int j = tid; j < n; j += blockDim.x) stage[t * RS + j] = ld_stream(src + j);
        }
        __syncthreads();
    } else if (tmax > 0) {
        mbar_wait(bar, 0);
    }

    float* st_s = stage + s * C;  // this sequence's column block
    float* lat = reinterpret_cast<float*>(smem + lay.lat) + s * lay.lat_seq;
    float* zz = s_zz + s * T;
    const bool run = have_seq && !bad && Tb > 0;
    const int mid = (Tb + 1) >> 1;
    const int rot = (RS & 1) ? 0 : 1;  // skew so that "lane per frame" accesses hit 32 distinct banks
    bool novalid = false;

    if (run) {
        // ================= pass 1: softmax statistics, lane per frame =================
        const int r_lo = role == 0 ? 0 : mid, r_hi = role == 0 ? mid : Tb;
        float slz = 0.0f;
        for (int t0 = r_lo; t0 < r_hi; t0 += 32) {
            const int t = t0 + lane;
            if (t < r_hi) {
                float* row = st_s + t * RS;
                int k0 = rot ? (lane % C) : 0;
                float m = -INFINITY;
                int k = k0;
#pragma unroll 4
                for (int j = 0; j < C; ++j) {
                    m = fmaxf(m, row[k]);
                    k = (k + 1 == C) ? 0 : k + 1;
                }
                const float ml = m * 1.4426950408889634f;
                float z = 0.0f;
                k = k0;
#pragma unroll 4
                for (int j = 0; j < C; ++j) {
                    const float e = fast_ex2(fmaf(row[k], 1.4426950408889634f, -ml));
                    row[k] = e;
                    z += e;
                    k = (k + 1 == C) ? 0 : k + 1;
                }
                zz[t] = z;
                slz += logf(z);
            }
        }
        slz = warp_sum(slz);
        if (lane == 0) s_slz[s * 8 + role] = slz;
        pair_barrier(1 + s);  // both halves of the staged block now hold e = exp(x - max)

        // ================= pass 2: lattice chains =================
        float* latB = lat;        // [t][i]   blank state 2i      (i = 0..L)
        float* latL = lat + Lp;   // [t][i]   label state 2i+1    (i = 0..L-1); row stride 2*Lp
        const int LS = 2 * Lp;
        int* exr = reinterpret_cast<int*>(smem + lay.exr) + s * T;  // per-frame scale exponent of the stored half
        if (role == 0) {
            // pair i = lane*NP + j : (ab = alpha(blank before label i), al = alpha(label i))
            float ab[NP], al[NP], skp[NP];
            int lb[NP], tdl[NP];
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const int i = lane * NP + j;
                const int li = (i < L) ? s_lab[i] : blank;
                lb[j] = li;
                skp[j] = (i >= 1 && i < L && s_lab[i - 1] != li) ? 1.0f : 0.0f;
                tdl[j] = (i < L) ? Tb - (L - i) : (i == L ? Tb : -1);  // label alive iff t <= tdl (i<L); blank alive iff t < tdl
                ab[j] = 0.0f; al[j] = 0.0f;
            }
            float sc = 1.0f;
            int Ea = 0, exn = 0;
            float eb_n = st_s[blank], el_n[NP];
#pragma unroll
            for (int j = 0; j < NP; ++j) el_n[j] = st_s[lb[j]];
            // one forward step: alpha_t from alpha_{t-1}; stored value v_t = alpha_t * 2^-Ea
            auto step = [&](int t) {
                const float eb = eb_n;
                float el[NP];
#pragma unroll
                for (int j = 0; j < NP; ++j) el[j] = el_n[j];
                if (t + 1 < Tb) {
                    const float* nrow = st_s + (t + 1) * RS;
                    eb_n = nrow[blank];
#pragma unroll
                    for (int j = 0; j < NP; ++j) el_n[j] = nrow[lb[j]];
                }
                float nbv[NP], nlv[NP];
                if (t == 0) {
#pragma unroll
                    for (int j = 0; j < NP; ++j) {
                        const int i = lane * NP + j;
                        nbv[j] = (i == 0) ? eb : 0.0f;
                        nlv[j] = (i == 0) ? el[j] : 0.0f;
                    }
                } else {
                    float up = __shfl_up_sync(kFullMask, al[NP - 1], 1);
                    if (lane == 0) up = 0.0f;
#pragma unroll
                    for (int j = 0; j < NP; ++j) {
                        const float pl = (j == 0) ? up : al[j - 1];
                        nbv[j] = eb * (ab[j] + pl);
                        nlv[j] = el[j] * (al[j] + ab[j] + skp[j] * pl);
                    }
                }
                float mloc = 0.0f;
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const int i = lane * NP + j;
                    ab[j] = (t < tdl[j]) ? nbv[j] * sc : 0.0f;
                    al[j] = (t <= tdl[j] && i < L) ? nlv[j] * sc : 0.0f;
                    mloc = fmaxf(mloc, fmaxf(ab[j], al[j]));
                }
                Ea += exn;
                sc = pow2_rescale(mloc, exn);
            };
            for (int t = 0; t < mid; ++t) {
                step(t);
                float* rb = latB + t * LS;
                float* rl = latL + t * LS;
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const int i = lane * NP + j;
                    if (i <= L) rb[i] = ab[j];
                    if (i < L) rl[i] = al[j];
                }
                if (lane == 0) exr[t] = Ea;
            }
            pair_barrier(1 + s);  // partner has stored beta_t (and its exponents) for t >= mid
            int Pt = 0;
            for (int t = mid; t < Tb; ++t) {
                step(t);
                float* rb = latB + t * LS;
                float* rl = latL + t * LS;
                float wb[NP], wl[NP];
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const int i = lane * NP + j;
                    wb[j] = (i <= L) ? rb[i] : 0.0f;
                    wl[j] = (i < L) ? rl[i] : 0.0f;
                }
                const int Es = Ea + exr[t];
                if (t == mid) Pt = Es + prod_exponent<NP>(ab, al, wb, wl);
                float f1, f2;
                boost_factors(Es - Pt, f1, f2);
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const int i = lane * NP + j;
                    if (i <= L) rb[i] = (ab[j] * f1) * (wb[j] * f2);
                    if (i < L) rl[i] = (al[j] * f1) * (wl[j] * f2);
                }
            }
            // p(z|x) in e-units: alpha(2L) + alpha(2L-1) at the last frame
            float up = __shfl_up_sync(kFullMask, al[NP - 1], 1);
            if (lane == 0) up = 0.0f;
            float pe = 0.0f;
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const int i = lane * NP + j;
                const float pl = (j == 0) ? up : al[j - 1];
                if (i == L) pe = ab[j] + pl;
            }
            pe = __shfl_sync(kFullMask, pe, L / NP);
            pair_barrier(1 + s);  // both chains done: products complete, partner's sum(log Z) visible
            novalid = !(pe > 0.0f);
            if (lane == 0) {
                const float sumlz = s_slz[s * 8] + s_slz[s * 8 + 1];
                const float lp = novalid ? -INFINITY : (logf(pe) + (float)Ea * 0.6931471805599453f - sumlz);
                loss[b] = -lp;
                status[b] = novalid ? kCtcRedo : 0;  // an all-zero lattice may be underflow: the exact kernel decides
                s_slz[s * 8 + 2] = novalid ? 1.0f : 0.0f;
                s_slz[s * 8 + 3] = novalid ? 0.0f : log2f(pe);
                reinterpret_cast<int*>(s_slz)[s * 8 + 4] = Ea;
                reinterpret_cast<int*>(s_slz)[s * 8 + 5] = Pt;
            }
        } else {
            // pair i = lane*NP + j : (bl = beta(label i-1), bb = beta(blank after label i-1))
            float bb[NP], bl[NP], skp[NP];
            int lb[NP], tbl[NP], tbb[NP];
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const int i = lane * NP + j;
                const int li = (i >= 1 && i <= L) ? s_lab[i - 1] : blank;
                lb[j] = li;
                skp[j] = (i >= 1 && i < L && s_lab[i] != li) ? 1.0f : 0.0f;
                tbl[j] = (i >= 1 && i <= L) ? i - 1 : 0x7fffffff;  // label i-1 alive iff t >= i-1
                tbb[j] = (i <= L) ? i : 0x7fffffff;                // blank i alive iff t >= i
                bb[j] = 0.0f; bl[j] = 0.0f;
            }
            float sc = 1.0f;
            int Eb = 0, exn = 0;
            float eb_n = 0.0f, el_n[NP];
#pragma unroll
            for (int j = 0; j < NP; ++j) el_n[j] = 0.0f;
            // one backward step: beta_t from beta_{t+1} and e_{t+1}; stored value w_t = beta_t * 2^-Eb
            auto step = [&](int t) {
                const float eb = eb_n;
                float el[NP];
#pragma unroll
                for (int j = 0; j < NP; ++j) el[j] = el_n[j];
                {   // e_t, needed by the NEXT step (t-1)
                    const float* nrow = st_s + t * RS;
                    eb_n = nrow[blank];
#pragma unroll
                    for (int j = 0; j < NP; ++j) el_n[j] = nrow[lb[j]];
                }
                float nbb[NP], nbl[NP];
                if (t == Tb - 1) {
#pragma unroll
                    for (int j = 0; j < NP; ++j) {
                        const int i = lane * NP + j;
                        nbb[j] = (i == L) ? 1.0f : 0.0f;
                        nbl[j] = (i == L && L >= 1) ? 1.0f : 0.0f;
                    }
                } else {
                    float wb[NP], wl[NP];
#pragma unroll
                    for (int j = 0; j < NP; ++j) {
                        wb[j] = bb[j] * eb;
                        wl[j] = bl[j] * el[j];
                    }
                    float dn = __shfl_down_sync(kFullMask, wl[0], 1);
                    if (lane == 31) dn = 0.0f;
#pragma unroll
                    for (int j = 0; j < NP; ++j) {
                        const float nx = (j == NP - 1) ? dn : wl[j + 1];
                        nbb[j] = wb[j] + nx;
                        nbl[j] = wl[j] + wb[j] + skp[j] * nx;
                    }
                }
                float mloc = 0.0f;
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    bb[j] = (t >= tbb[j]) ? nbb[j] * sc : 0.0f;
                    bl[j] = (t >= tbl[j]) ? nbl[j] * sc : 0.0f;
                    mloc = fmaxf(mloc, fmaxf(bb[j], bl[j]));
                }
                Eb += exn;
                sc = pow2_rescale(mloc, exn);
            };
            for (int t = Tb - 1; t >= mid; --t) {
                step(t);
                float* rb = latB + t * LS;
                float* rl = latL + t * LS;
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const int i = lane * NP + j;
                    if (i <= L) rb[i] = bb[j];
                    if (i >= 1 && i <= L) rl[i - 1] = bl[j];
                }
                if (lane == 0) exr[t] = Eb;
            }
            pair_barrier(1 + s);  // partner has stored alpha_t (and its exponents) for t < mid
            int Pt = 0;
            for (int t = mid - 1; t >= 0; --t) {
                step(t);
                float* rb = latB + t * LS;
                float* rl = latL + t * LS;
                float vb[NP], vl[NP];
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const int i = lane * NP + j;
                    vb[j] = (i <= L) ? rb[i] : 0.0f;
                    vl[j] = (i >= 1 && i <= L) ? rl[i - 1] : 0.0f;
                }
                const int Es = Eb + exr[t];
                if (t == mid - 1) Pt = Es + prod_exponent<NP>(bb, bl, vb, vl);
                float f1, f2;
                boost_factors(Es - Pt, f1, f2);
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const int i = lane * NP + j;
                    if (i <= L) rb[i] = (bb[j] * f1) * (vb[j] * f2);
                    if (i >= 1 && i <= L) rl[i - 1] = (bl[j] * f1) * (vl[j] * f2);
                }
            }
            if (lane == 0) reinterpret_cast<int*>(s_slz)[s * 8 + 6] = Pt;
            pair_barrier(1 + s);
        }
        pair_barrier(1 + s);  // flags / log2 p written by the alpha warp
        novalid = s_slz[s * 8 + 2] != 0.0f;
    } else if (have_seq && role == 0 && lane == 0) {
        // TF: zero-length sequence -> loss 0, grad 0.  Infeasible / invalid -> flagged, zero outputs.
        loss[b] = 0.0f;
        status[b] = bad;
    }

    // ================= pass 3: posterior scatter + gradient rows, lane per frame =================
    // Every row of products must sum to the same p(z|x) the alpha chain ended with; a row that does not has
    // lost probability mass to float32 underflow somewhere in the chains (states more than 2^-126 below the
    // warp-wide maximum are flushed).  Such sequences are handed to the exact log-domain kernel (kCtcRedo).
    if (have_seq) {
        const int Tr = run ? Tb : 0;
        const int r_lo = role == 0 ? 0 : mid, r_hi = role == 0 ? mid : Tr;
        const float* latB = lat;
        const float* latL = lat + Lp;
        const int LS = 2 * Lp;
        const float l2pe = s_slz[s * 8 + 3];
        const float dexp = (float)(reinterpret_cast<const int*>(s_slz)[s * 8 + (role == 0 ? 6 : 5)] -
                                   reinterpret_cast<const int*>(s_slz)[s * 8 + 4]);
        bool lost = false;
        for (int t0 = r_lo; t0 < r_hi; t0 += 32) {
            const int t = t0 + lane;
            if (t < r_hi) {
                float* row = st_s + t * RS;
                const float z = zz[t];
                if (!novalid) {
                    const float* rb = latB + t * LS;
                    const float* rl = latL + t * LS;
                    float S = 0.0f, Bs = 0.0f;
                    int i = lane % (L + 1);
                    for (int j = 0; j <= L; ++j) {
                        const float pb = rb[i];
                        Bs += pb;
                        if (i < L) S += rl[i];
                        i = (i == L) ? 0 : i + 1;
                    }
                    S += Bs;
                    lost = lost || !(fabsf(log2f(S) - l2pe + dexp) < 9.765625e-4f);
                    if (!grad) continue;
                    const float r = (S > 0.0f) ? z / S : 0.0f;
                    row[blank] -= Bs * r;
                    if (L > 0) {
                        i = lane % L;
                        for (int j = 0; j < L; ++j) {
                            row[s_lab[i]] -= rl[i] * r;
                            i = (i + 1 == L) ? 0 : i + 1;
                        }
                    }
                }
                if (!grad) continue;
                const float gs = grad_scale / z;
                int k = rot ? (lane % C) : 0;
#pragma unroll 4
                for (int j = 0; j < C; ++j) {
                    row[k] *= gs;
                    k = (k + 1 == C) ? 0 : k + 1;
                }
            }
        }
        if (run && !novalid && __any_sync(kFullMask, lost) && lane == 0) status[b] = kCtcRedo;
    }
    if (!grad) return;
    if (have_seq) {
        const int Tr = run ? Tb : 0;
        // frames past the sequence end: zero gradient
        const int zr = T - Tr;
        for (int t0 = Tr + role * ((zr + 1) >> 1), te = role == 0 ? Tr + ((zr + 1) >> 1) : T; t0 < te; t0 += 32) {
            const int t = t0 + lane;
            if (t < te) {
                float* row = st_s + t * RS;
                int k = rot ? (lane % C) : 0;
                for (int j = 0; j < C; ++j) {
                    row[k] = 0.0f;
                    k = (k + 1 == C) ? 0 : k + 1;
                }
            }
        }
    }
    if (bulk) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            const unsigned row_bytes = (unsigned)(G * C * 4);
            float* dst = grad + (size_t)b0 * C;
            const unsigned src = smem_u32(stage);
            for (int t = 0; t < T; ++t) bulk_store(dst + (size_t)t * B * C, src + (unsigned)t * RS * 4, row_bytes);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    } else {
        __syncthreads();
        const int n = nb * C;
        for (int t = 0; t < T; ++t) {
            float* dst = grad + ((size_t)t * B + b0) * C;
            for (int j = tid; j < n; j += blockDim.x) st_stream(dst + j, stage[t * RS + j]);
        }
    }
}

}  // namespace ocr
