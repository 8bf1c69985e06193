// CTC beam-search decoder for sm_100a.  Replaces tf.nn.ctc_beam_search_decoder as called by the
// reference at /root/reference/src/weinman/test.py:84-88 (beam_width=128, top_paths=1,
// merge_repeated=True) and src/weinman/client.py:227-231 (merge_repeated=False).
//
// Semantics follow upstream TensorFlow's CTCBeamSearchDecoder::Step / TopPaths
// INCLUDING its order-dependent side effects: beams are expanded sequentially in descending order of
// their previous score, every candidate child is offered to a bounded best-`beam_width` list in
// (beam, class) order, a candidate must beat the CURRENT worst entry strictly, and a beam that was
// pushed out before its parent is expanded loses its own expansion (TF resets its `oldp`).  These
// effects change the decoded labels on flat distributions, so they are reproduced, not approximated.
// Ties (frequent: the logits come out of a ReLU) are ordered by (score desc, push order asc); TF's
// own tie order is libstdc++-heap dependent (documented deviation).
//
// Mapping to the GPU: the work per sequence is a chain of data-dependent list updates, so ONE WARP
// owns one sequence and the batch supplies the parallelism (B warps in flight, no block barriers):
//   * the best-list R (<=128 entries) lives in registers, 4 sorted 64-bit keys per lane,
//     key = (order-preserving bits of the float score) << 32 | ~push_order;
//     insertion = one redux.sync rank + one shuffle; the worst entry is a broadcast of the last key;
//   * beam state (probabilities, label, parent slot, prefix hash) lives in shared memory, double
//     buffered per frame; prefix identity is a 64-bit hash chain, so a re-created prefix node is
//     recognised without TF's pointer tree; back-pointers for the final path go to a global pool;
//   * log-probabilities use det_math.cuh so scores are bit-identical to the CPU oracle.
// Bound: latency / dependency chain (T frames x up to beam_width expansions), not HBM or tensor.
#include "common.cuh"
#include "det_math.cuh"

namespace ocr {

typedef unsigned long long u64;
constexpr int kSlots = 128;     // maximum beam width
constexpr int kBeamWarps = 4;   // sequences per CTA
constexpr unsigned kNegInfOrd = 0x007FFFFFu;  // ord(-inf)
constexpr u64 kRootHash = 0x9E3779B97F4A7C15ull;

__device__ __forceinline__ unsigned ord(float f) {
    const unsigned u = __float_as_uint(__fadd_rn(f, 0.0f));  // -0 -> +0
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unord(unsigned o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
__device__ __forceinline__ u64 mix_hash(u64 h, int k) {
    u64 z = h + (u64)(k + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ u64 shfl64(u64 v, int src) {
    unsigned lo = __shfl_sync(kFullMask, (unsigned)v, src);
    unsigned hi = __shfl_sync(kFullMask, (unsigned)(v >> 32), src);
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ u64 shfl64_xor(u64 v, int m) {
    unsigned lo = __shfl_xor_sync(kFullMask, (unsigned)v, m);
    unsigned hi = __shfl_xor_sync(kFullMask, (unsigned)(v >> 32), m);
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ u64 shfl64_up1(u64 v) {
    unsigned lo = __shfl_up_sync(kFullMask, (unsigned)v, 1);
    unsigned hi = __shfl_up_sync(kFullMask, (unsigned)(v >> 32), 1);
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ void cmpswap(u64& a, u64& b, bool desc) {
    const bool sw = desc ? (a < b) : (a > b);
    const u64 ta = sw ? b : a, tb = sw ? a : b;
    a = ta; b = tb;
}
// position p = lane*4 + r; sorts all 128 keys descending (bitonic network, registers + shuffles)
__device__ __forceinline__ void bitonic_sort_desc(u64 (&key)[4], int lane) {
#pragma unroll 1
    for (int k = 2; k <= kSlots; k <<= 1) {
        const bool desc = ((lane * 4) & k) == 0;
#pragma unroll 1
        for (int j = k >> 1; j >= 4; j >>= 1) {
            const int lm = j >> 2;
            const bool takemax = (((lane & lm) == 0) == desc);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const u64 o = shfl64_xor(key[r], lm);
                key[r] = takemax ? (key[r] > o ? key[r] : o) : (key[r] < o ? key[r] : o);
            }
        }
        if (k >= 4) {
            cmpswap(key[0], key[2], desc);
            cmpswap(key[1], key[3], desc);
        }
        cmpswap(key[0], key[1], k == 2 ? true : desc);
        cmpswap(key[2], key[3], k == 2 ? false : desc);
    }
}
__device__ __forceinline__ u64 pick(const u64 (&key)[4], int r) {
    return r == 0 ? key[0] : (r == 1 ? key[1] : (r == 2 ? key[2] : key[3]));
}
// insert nk (known to rank ahead of position K-1) into the sorted list; positions >= K stay empty
__device__ __forceinline__ void insert_key(u64 (&key)[4], u64 nk, int lane, int K) {
    const int c = (key[0] > nk) + (key[1] > nk) + (key[2] > nk) + (key[3] > nk);
    const int q = __reduce_add_sync(kFullMask, c);
    const int ql = q >> 2, qr = q & 3;
    const u64 from_prev = shfl64_up1(key[3]);
    if (lane > ql) {
        key[3] = key[2]; key[2] = key[1]; key[1] = key[0]; key[0] = from_prev;
    } else if (lane == ql) {
        const u64 k0 = key[0], k1 = key[1], k2 = key[2];
        key[3] = (qr == 3) ? nk : k2;
        key[2] = (qr == 2) ? nk : (qr < 2 ? k1 : k2);
        key[1] = (qr == 1) ? nk : (qr < 1 ? k0 : k1);
        key[0] = (qr == 0) ? nk : k0;
    }
    if (K < kSlots && lane == (K >> 2)) {
        const int kr = K & 3;
        if (kr == 0) key[0] = 0; else if (kr == 1) key[1] = 0; else if (kr == 2) key[2] = 0; else key[3] = 0;
    }
}

struct BeamLayout {
    int tot, blk, labp, label, hash, phash, pool, ps;  // double buffered: second copy at +half
    int n_tot, n_blk, n_lab, newpos, inR, oreset, in, ex, cm;
    int per_warp, CW, Cpad;
};
__host__ __device__ inline BeamLayout beam_layout(int C) {
    BeamLayout L;
    L.CW = (C + 31) / 32;
    L.Cpad = L.CW * 32;
    int o = 0;
    L.hash = o;  o += 2 * kSlots * 8;
    L.phash = o; o += 2 * kSlots * 8;
    L.tot = o;   o += 2 * kSlots * 4;
    L.blk = o;   o += 2 * kSlots * 4;
    L.labp = o;  o += 2 * kSlots * 4;
    L.label = o; o += 2 * kSlots * 4;
    L.pool = o;  o += 2 * kSlots * 4;
    L.ps = o;    o += 2 * kSlots * 4;
    L.n_tot = o; o += kSlots * 4;
    L.n_blk = o; o += kSlots * 4;
    L.n_lab = o; o += kSlots * 4;
    L.newpos = o; o += kSlots * 4;
    L.in = o;    o += L.Cpad * 4;
    L.ex = o;    o += L.Cpad * 4;
    L.cm = o;    o += kSlots * L.CW * 4;
    L.inR = o;   o += kSlots;
    L.oreset = o; o += kSlots;
    L.per_warp = (o + 15) & ~15;
    return L;
}

__global__ void __launch_bounds__(kBeamWarps * 32)
ctc_beam_kernel(const float* __restrict__ logits, int T, int B, int C, const int32_t* __restrict__ seq_len, int K,
                int top_paths, int merge_repeated, int normalize, int64_t* __restrict__ decoded,
                int32_t* __restrict__ decoded_len, float* __restrict__ log_prob, int2* __restrict__ pool_ws)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * kBeamWarps + warp;
    if (b >= B) return;
    const BeamLayout L = beam_layout(C);
    unsigned char* base = smem_raw + (size_t)warp * L.per_warp;
    u64* s_hash = reinterpret_cast<u64*>(base + L.hash);
    u64* s_phash = reinterpret_cast<u64*>(base + L.phash);
    float* s_tot = reinterpret_cast<float*>(base + L.tot);
    float* s_blk = reinterpret_cast<float*>(base + L.blk);
    float* s_labp = reinterpret_cast<float*>(base + L.labp);
    int* s_label = reinterpret_cast<int*>(base + L.label);
    int* s_pool = reinterpret_cast<int*>(base + L.pool);
    int* s_ps = reinterpret_cast<int*>(base + L.ps);
    float* n_tot = reinterpret_cast<float*>(base + L.n_tot);
    float* n_blk = reinterpret_cast<float*>(base + L.n_blk);
    float* n_lab = reinterpret_cast<float*>(base + L.n_lab);
    int* s_newpos = reinterpret_cast<int*>(base + L.newpos);
    float* s_in = reinterpret_cast<float*>(base + L.in);
    float* s_ex = reinterpret_cast<float*>(base + L.ex);
    unsigned* s_cm = reinterpret_cast<unsigned*>(base + L.cm);
    volatile unsigned char* s_inR = base + L.inR;
    volatile unsigned char* s_oreset = base + L.oreset;
    const int CW = L.CW;
    const int blank = C - 1;
    const float NEG = -CUDART_INF_F;

    const int Tb = min(max(seq_len[b], 0), T);
    int2* pool = pool_ws + (size_t)b * (1 + (size_t)T * K);

    // ---- initial beam: the root (empty prefix) with P_blank = P_total = 1
    int cur = 0;
    for (int i = lane; i < kSlots; i += 32) {
        s_tot[i] = (i == 0) ? 0.0f : NEG;
        s_blk[i] = (i == 0) ? 0.0f : NEG;
        s_labp[i] = NEG;
        s_label[i] = -1;
        s_hash[i] = (i == 0) ? kRootHash : 0;
        s_phash[i] = 0;
        s_pool[i] = 0;
        s_ps[i] = -1;
    }
    for (int i = lane; i < kSlots * CW; i += 32) s_cm[i] = 0;
    if (lane == 0) pool[0] = make_int2(-1, -1);
    int nb = 1;
    __syncwarp();

    u64 key[4];
    for (int t = 0; t < Tb; ++t) {
        const int co = cur * kSlots, no = (cur ^ 1) * kSlots;
        // ---- per-frame scores in[k] = logit - (max [+ log sum exp])   (CTCBeamSearchDecoder::Step head)
        const float* row = logits + ((size_t)t * B + b) * C;
        float m = NEG;
        for (int k = lane; k < C; k += 32) {
            const float v = ld_stream(row + k);
            s_in[k] = v;
            m = fmaxf(m, v);
        }
        m = warp_max(m);
        float off = m;
        if (normalize) {
            for (int k = lane; k < C; k += 32) s_ex[k] = det_expf(__fadd_rn(s_in[k], -m));
            __syncwarp();
            if (lane == 0) {
                float s = 0.0f;
                for (int k = 0; k < C; ++k) s = __fadd_rn(s, s_ex[k]);
                off = __fadd_rn(m, det_logf(s));
            }
            off = __shfl_sync(kFullMask, off, 0);
        }
        for (int k = lane; k < C; k += 32) s_in[k] = __fadd_rn(s_in[k], -off);
        __syncwarp();
        const float in_blank = s_in[blank];

        // ---- loop 1: advance every beam of the previous frame, then sort them into R
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = lane * 4 + r;
            u64 kk = (u64)(~(unsigned)i);  // empty pad (high word 0), distinct
            if (i < nb) {
                const float ot = s_tot[co + i];
                float nl = s_labp[co + i];
                const int lb = s_label[co + i];
                if (lb >= 0) {
                    const int p = s_ps[co + i];
                    if (p >= 0) {
                        const float prev = (lb == s_label[co + p]) ? s_blk[co + p] : s_tot[co + p];
                        nl = det_lse2(nl, prev);
                    }
                    nl = __fadd_rn(nl, s_in[lb]);
                }
                const float nbk = __fadd_rn(ot, in_blank);
                const float nt = det_lse2(nbk, nl);
                n_tot[i] = nt; n_blk[i] = nbk; n_lab[i] = nl;
                s_inR[i] = 1; s_oreset[i] = 0;
                kk = ((u64)ord(nt) << 32) | (u64)(~(unsigned)i);
            }
            key[r] = kk;
        }
        bitonic_sort_desc(key, lane);
        __syncwarp();
        u64 bottom = shfl64(pick(key, (K - 1) & 3), (K - 1) >> 2);

        // ---- loop 2: grow children, beam by beam in descending previous score
        for (int i = 0; i < nb; ++i) {
            if (s_oreset[i]) continue;  // TF: oldp was reset -> is_candidate(oldp) fails
            const float ot = s_tot[co + i];
            unsigned thr = max((unsigned)(bottom >> 32), kNegInfOrd);
            if (!(ord(ot) > thr)) break;  // beams are in descending oldp order and the bar only rises
            const float ob = s_blk[co + i];
            const int lb = s_label[co + i];
            const unsigned seq0 = (unsigned)nb + (unsigned)i * (unsigned)C;
            for (int kb = 0; kb < blank; kb += 32) {
                const int k = kb + lane;
                const bool valid = k < blank;
                const unsigned word = s_cm[i * CW + (kb >> 5)];
                const bool child = valid && ((word >> lane) & 1u);
                float s = NEG;
                if (valid) s = __fadd_rn(s_in[k], (k == lb) ? ob : ot);
                const unsigned os = ord(s);
                const bool pass = valid && !child && os > thr;
                unsigned ev = __ballot_sync(kFullMask, pass || child);
                while (ev) {
                    const int j = __ffs(ev) - 1;
                    ev &= ev - 1;
                    if ((word >> j) & 1u) {
                        // child (i, kb+j) already exists as a beam: active -> skip; pushed out -> TF
                        // re-creates it, the re-creation cannot beat the bar, and its oldp is reset
                        int found = -1;
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const int e = lane * 4 + r;
                            const bool mt = e < nb && s_ps[co + e] == i && s_label[co + e] == kb + j;
                            const unsigned bm = __ballot_sync(kFullMask, mt);
                            if (bm) found = (__ffs(bm) - 1) * 4 + r;
                        }
                        if (found >= 0 && !s_inR[found] && lane == 0) s_oreset[found] = 1;
                        __syncwarp();
                    } else {
                        const unsigned osj = __shfl_sync(kFullMask, os, j);
                        const unsigned thr2 = max((unsigned)(bottom >> 32), kNegInfOrd);
                        if (osj > thr2) {
                            if ((bottom >> 32) != 0) {  // list full: the worst entry leaves the beam
                                const unsigned eseq = ~(unsigned)bottom;
                                if (eseq < (unsigned)nb && lane == 0) s_inR[eseq] = 0;
                            }
                            const u64 nk = ((u64)osj << 32) | (u64)(~(seq0 + (unsigned)(kb + j)));
                            insert_key(key, nk, lane, K);
                            bottom = shfl64(pick(key, (K - 1) & 3), (K - 1) >> 2);
                            __syncwarp();
                        }
                    }
                }
                thr = max((unsigned)(bottom >> 32), kNegInfOrd);
            }
        }

        // ---- rebuild the beam state from R (already sorted: next frame's expansion order)
        for (int i = lane; i < kSlots; i += 32) s_newpos[i] = -1;
        __syncwarp();
        int live_cnt = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int p = lane * 4 + r;
            const u64 kk = key[r];
            if ((kk >> 32) != 0) {
                ++live_cnt;
                const unsigned seq = ~(unsigned)kk;
                if (seq < (unsigned)nb) {
                    const int i = (int)seq;
                    s_tot[no + p] = n_tot[i]; s_blk[no + p] = n_blk[i]; s_labp[no + p] = n_lab[i];
                    s_label[no + p] = s_label[co + i];
                    s_hash[no + p] = s_hash[co + i]; s_phash[no + p] = s_phash[co + i];
                    s_pool[no + p] = s_pool[co + i];
                    s_newpos[i] = p;
                } else {
                    const unsigned c = seq - (unsigned)nb;
                    const int i = (int)(c / (unsigned)C);
                    const int k = (int)(c - (unsigned)i * (unsigned)C);
                    const float s = unord((unsigned)(kk >> 32));
                    s_tot[no + p] = s; s_blk[no + p] = NEG; s_labp[no + p] = s;
                    s_label[no + p] = k;
                    const u64 ph = s_hash[co + i];
                    s_phash[no + p] = ph;
                    s_hash[no + p] = mix_hash(ph, k);
                    const int id = 1 + t * K + p;
                    s_pool[no + p] = id;
                    pool[id] = make_int2(s_pool[co + i], k);
                }
            }
        }
        const int nb_new = warp_sum_int(live_cnt);
        __syncwarp();
        // parent slots
        unsigned orphan_r = 0;  // bit r: my slot r is a surviving beam whose parent was not in the beam
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int p = lane * 4 + r;
            const u64 kk = key[r];
            if ((kk >> 32) != 0) {
                const unsigned seq = ~(unsigned)kk;
                int psn;
                if (seq < (unsigned)nb) {
                    const int po = s_ps[co + (int)seq];
                    psn = (po >= 0) ? s_newpos[po] : -1;
                    if (po < 0 && s_label[co + (int)seq] >= 0) orphan_r |= 1u << r;
                } else {
                    const int i = (int)((seq - (unsigned)nb) / (unsigned)C);
                    psn = s_newpos[i];
                }
                s_ps[no + p] = psn;
            }
        }
        __syncwarp();
        // a surviving beam whose parent prefix was re-created in this frame gets its parent back
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            unsigned om = __ballot_sync(kFullMask, (orphan_r >> r) & 1u);
            while (om) {
                const int src = __ffs(om) - 1;
                om &= om - 1;
                const int po = src * 4 + r;
                const u64 ph = s_phash[no + po];
                int found = -1;
#pragma unroll
                for (int r2 = 0; r2 < 4; ++r2) {
                    const u64 kk = key[r2];
                    const bool isnew = (kk >> 32) != 0 && (~(unsigned)kk) >= (unsigned)nb;
                    const bool mt = isnew && s_hash[no + lane * 4 + r2] == ph;
                    const unsigned bm = __ballot_sync(kFullMask, mt);
                    if (bm) found = (__ffs(bm) - 1) * 4 + r2;
                }
                if (found >= 0 && lane == 0) s_ps[no + po] = found;
            }
        }
        for (int i = lane; i < kSlots * CW; i += 32) s_cm[i] = 0;
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int p = lane * 4 + r;
            if ((key[r] >> 32) != 0) {
                const int psn = s_ps[no + p];
                const int lb = s_label[no + p];
                if (psn >= 0 && lb >= 0) atomicOr(&s_cm[psn * CW + (lb >> 5)], 1u << (lb & 31));
            }
        }
        __syncwarp();
        cur ^= 1;
        nb = nb_new;
    }

    // ---- TopPaths: beams are sorted; walk the back-pointers of the first top_paths
    __threadfence_block();
    __syncwarp();
    const int co = cur * kSlots;
    for (int p = lane; p < top_paths; p += 32) {
        int64_t* out = decoded + ((size_t)b * top_paths + p) * T;
        int n = 0;
        float lp = NEG;  // TF raises when fewer leaves than requested paths exist
        if (p < nb) {
            lp = s_tot[co + p];
            int id = s_pool[co + p];
            int prev = -1;
            while (id > 0) {  // collect from the leaf towards the root, stored from the end of the row
                const int2 nd = pool[id];
                if (!merge_repeated || nd.y != prev) { out[T - 1 - n] = nd.y; ++n; }
                prev = nd.y;
                id = nd.x;
            }
            for (int q = 0; q < n; ++q) out[q] = out[T - n + q];
        }
        for (int q = n; q < T; ++q) out[q] = -1;
        decoded_len[(size_t)b * top_paths + p] = n;
        log_prob[(size_t)b * top_paths + p] = lp;
    }
}


// =================================================================================================================
// CTA per sequence (default path).  The warp-per-sequence kernel above replays TensorFlow's list updates one insertion at
// a time: ~1000 insertions per frame at ~300 cycles of dependent warp collectives each (12.4 ms for 1024 sequences, and the
// same 11.8 ms for 128 of them: pure latency).  Its result, though, can be stated without the replay:
//   * at every moment the bounded list holds exactly the top K (by key = score, then push order) of everything offered so
//     far -- a candidate enters iff it beats the current worst strictly, and the worst only rises;
//   * the only way the ORDER of the offers reaches the result is TF's reset: when beam i is expanded and comes to the
//     class of one of its children c that is already a beam, c loses its own expansion if its carried-over entry a_c has
//     been pushed out by then, i.e. iff  #{entries of the old beam ahead of a_c} + #{candidates offered before that moment
//     with a score above a_c's}  >= K.  (The early exit on beams whose old score is not above the bar needs no replay: all
//     their candidates lie at or below the bar.)
// So a frame is: (1) per-beam candidate counts above a threshold by binary search in the frame's sorted class scores
// (fl(in + old) is monotone in both arguments), (2) the K-th largest score by an 8-way search over those counts, (3) the
// reset test for the few children whose carried-over entry does not survive, by a staircase walk over (sorted classes) x
// (beams in list order), resolved in time order when one fires, (4) winners written by prefix sums in push order, ranked by
// counting, and the list rebuilt exactly as above.  Thread i of the 128 owns beam slot i.
__device__ long long g_beam_prof[16];   // tuning: cycles per phase of CTA 0, summed over its frames (ocr_debug_beam_profile)
constexpr int kCtaThreads = 128;
constexpr int kCtaWarps = kCtaThreads / 32;

struct CtaLayout {
    int hash, phash, tot, blk, labp, label, pool, ps;   // double buffered (second copy at +kSlots elements)
    int n_tot, n_blk, n_lab, newpos, in, ex, sin, scls, tmp, cm, akey, wkey, skey, exs, lbs, exoff, cnt0, evl, red, flags, total;
    int CW, Cpad;
};
__host__ __device__ inline CtaLayout cta_layout(int C) {
    CtaLayout L;
    L.CW = (C + 31) / 32;
    L.Cpad = L.CW * 32;
    int o = 0;
    L.hash = o;  o += 2 * kSlots * 8;
    L.phash = o; o += 2 * kSlots * 8;
    L.akey = o;  o += kSlots * 8;
    L.wkey = o;  o += kSlots * 8;
    L.skey = o;  o += kSlots * 8;
    L.tot = o;   o += 2 * kSlots * 4;
    L.blk = o;   o += 2 * kSlots * 4;
    L.labp = o;  o += 2 * kSlots * 4;
    L.label = o; o += 2 * kSlots * 4;
    L.pool = o;  o += 2 * kSlots * 4;
    L.ps = o;    o += 2 * kSlots * 4;
    L.n_tot = o; o += kSlots * 4;
    L.n_blk = o; o += kSlots * 4;
    L.n_lab = o; o += kSlots * 4;
    L.newpos = o; o += kSlots * 4;
    L.in = o;    o += L.Cpad * 4;
    L.ex = o;    o += L.Cpad * 4;
    L.sin = o;   o += L.Cpad * 4;
    L.scls = o;  o += L.Cpad * 4;
    L.tmp = o;   o += 2 * kSlots * 4;
    L.cm = o;    o += kSlots * L.CW * 4;
    L.exs = o;   o += 2 * kSlots * 4;      // scores (with the beam's total) of the classes a beam does not offer: its children and its own label
    L.lbs = o;   o += kSlots * 4;          // score of the beam's own label offered with its blank-ending probability
    L.exoff = o; o += (kSlots + 1) * 4;
    L.cnt0 = o;  o += kSlots * 4;
    L.evl = o;   o += kSlots * 4;
    L.red = o;   o += 2 * kCtaWarps * 8 * 4;
    L.flags = o; o += kSlots + 16;
    L.total = (o + 15) & ~15;
    return L;
}

__global__ void __launch_bounds__(kCtaThreads)
ctc_beam_cta_kernel(const float* __restrict__ logits, int T, int B, int C, const int32_t* __restrict__ seq_len, int K,
                    int top_paths, int merge_repeated, int normalize, int64_t* __restrict__ decoded,
                    int32_t* __restrict__ decoded_len, float* __restrict__ log_prob, int2* __restrict__ pool_ws)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    const CtaLayout L = cta_layout(C);
    unsigned char* base = smem_raw;
    u64* s_hash = reinterpret_cast<u64*>(base + L.hash);
    u64* s_phash = reinterpret_cast<u64*>(base + L.phash);
    u64* s_akey = reinterpret_cast<u64*>(base + L.akey);
    u64* s_wkey = reinterpret_cast<u64*>(base + L.wkey);
    u64* s_skey = reinterpret_cast<u64*>(base + L.skey);
    float* s_tot = reinterpret_cast<float*>(base + L.tot);
    float* s_blk = reinterpret_cast<float*>(base + L.blk);
    float* s_labp = reinterpret_cast<float*>(base + L.labp);
    int* s_label = reinterpret_cast<int*>(base + L.label);
    int* s_pool = reinterpret_cast<int*>(base + L.pool);
    int* s_ps = reinterpret_cast<int*>(base + L.ps);
    float* n_tot = reinterpret_cast<float*>(base + L.n_tot);
    float* n_blk = reinterpret_cast<float*>(base + L.n_blk);
    float* n_lab = reinterpret_cast<float*>(base + L.n_lab);
    int* s_newpos = reinterpret_cast<int*>(base + L.newpos);
    float* s_in = reinterpret_cast<float*>(base + L.in);
    float* s_ex = reinterpret_cast<float*>(base + L.ex);
    float* s_sin = reinterpret_cast<float*>(base + L.sin);
    int* s_scls = reinterpret_cast<int*>(base + L.scls);
    unsigned* s_tmp = reinterpret_cast<unsigned*>(base + L.tmp);
    unsigned* s_cm = reinterpret_cast<unsigned*>(base + L.cm);
    float* s_exs = reinterpret_cast<float*>(base + L.exs);
    float* s_lbs = reinterpret_cast<float*>(base + L.lbs);
    int* s_exoff = reinterpret_cast<int*>(base + L.exoff);
    int* s_cnt0 = reinterpret_cast<int*>(base + L.cnt0);
    int* s_evl = reinterpret_cast<int*>(base + L.evl);
    int* s_red = reinterpret_cast<int*>(base + L.red);
    volatile unsigned char* s_reset = base + L.flags;
    volatile int* s_misc = reinterpret_cast<volatile int*>(base + L.flags + kSlots);   // [0] any event fired, [1] resets, [2] new beam count, [3] gather counter / result
    const int CW = L.CW;
    const int blank = C - 1;
    const int NCLS = blank;
    const float NEG = -CUDART_INF_F;

    const int Tb = min(max(seq_len[b], 0), T);
    int2* pool = pool_ws + (size_t)b * (1 + (size_t)T * K);
    int redph = 0;

    // ---- block-wide helpers (4 warps; one barrier each: the scratch alternates)
    auto block_sum = [&](int v) -> int {
        v = __reduce_add_sync(kFullMask, v);
        int* r = s_red + (redph & 1) * (kCtaWarps * 8);
        redph++;
        if (lane == 0) r[warp] = v;
        __syncthreads();
        return r[0] + r[1] + r[2] + r[3];
    };
    auto block_max_u = [&](unsigned v) -> unsigned {
        v = __reduce_max_sync(kFullMask, v);
        unsigned* r = reinterpret_cast<unsigned*>(s_red + (redph & 1) * (kCtaWarps * 8));
        redph++;
        if (lane == 0) r[warp] = v;
        __syncthreads();
        return max(max(r[0], r[1]), max(r[2], r[3]));
    };
    // exclusive prefix sum over the 128 threads; *total receives the sum
    auto block_scan = [&](int v, int* total) -> int {
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(kFullMask, inc, o);
            if (lane >= o) inc += u;
        }
        int* r = s_red + (redph & 1) * (kCtaWarps * 8);
        redph++;
        if (lane == 31) r[warp] = inc;
        __syncthreads();
        int off = 0;
#pragma unroll
        for (int w = 0; w < kCtaWarps; ++w) if (w < warp) off += r[w];
        *total = r[0] + r[1] + r[2] + r[3];
        return off + inc - v;
    };

    // ---- initial beam: the root (empty prefix) with P_blank = P_total = 1
    int cur = 0;
    {
        const int i = tid;
        s_tot[i] = (i == 0) ? 0.0f : NEG;
        s_blk[i] = (i == 0) ? 0.0f : NEG;
        s_labp[i] = NEG;
        s_label[i] = -1;
        s_hash[i] = (i == 0) ? kRootHash : 0;
        s_phash[i] = 0;
        s_pool[i] = 0;
        s_ps[i] = -1;
    }
    for (int i = tid; i < kSlots * CW; i += kCtaThreads) s_cm[i] = 0;
    if (tid == 0) pool[0] = make_int2(-1, -1);
    int nb = 1;
    __syncthreads();

    long long tprev = clock64();
    for (int t = 0; t < Tb; ++t) {
        const int co = cur * kSlots, no = (cur ^ 1) * kSlots;
        // ---- per-frame scores in[k] = logit - (max [+ log sum exp])   (CTCBeamSearchDecoder::Step head)
        const float* row = logits + ((size_t)t * B + b) * C;
        float m = NEG;
        for (int k = tid; k < C; k += kCtaThreads) {
            const float v = ld_stream(row + k);
            s_in[k] = v;
            m = fmaxf(m, v);
        }
        m = unord(block_max_u(ord(m)));
        float off = m;
        if (normalize) {
            for (int k = tid; k < C; k += kCtaThreads) s_ex[k] = det_expf(__fadd_rn(s_in[k], -m));
            __syncthreads();
            if (tid == 0) {
                float sacc = 0.0f;
                for (int k = 0; k < C; ++k) sacc = __fadd_rn(sacc, s_ex[k]);   // in class order: the oracle's order
                s_ex[0] = __fadd_rn(m, det_logf(sacc));
            }
            __syncthreads();
            off = s_ex[0];
        }
        __syncthreads();
        for (int k = tid; k < C; k += kCtaThreads) s_in[k] = __fadd_rn(s_in[k], -off);
        __syncthreads();
        const float in_blank = s_in[blank];
        // the frame's label scores in descending order (ties by class: any order of equal values serves the counts)
        for (int k = tid; k < NCLS; k += kCtaThreads) {
            const float v = s_in[k];
            int rank = 0;
#pragma unroll 8
            for (int k2 = 0; k2 < NCLS; ++k2) {
                const float v2 = s_in[k2];
                rank += (v2 > v) || (v2 == v && k2 < k);
            }
            s_sin[rank] = v;
            s_scls[rank] = k;
        }

        if (blockIdx.x == 0 && tid == 0) { const long long now = clock64(); g_beam_prof[0] += now - tprev; tprev = now; }
        // ---- advance every beam of the previous frame (the carried-over entries a_i)
        const int i = tid;
        const bool have = i < nb;
        float ot = NEG, ob = NEG;
        int lb = -1;
        unsigned ont = 0;
        if (have) {
            ot = s_tot[co + i];
            ob = s_blk[co + i];
            lb = s_label[co + i];
            float nl = s_labp[co + i];
            if (lb >= 0) {
                const int p = s_ps[co + i];
                if (p >= 0) {
                    const float prev = (lb == s_label[co + p]) ? s_blk[co + p] : s_tot[co + p];
                    nl = det_lse2(nl, prev);
                }
                nl = __fadd_rn(nl, s_in[lb]);
            }
            const float nbk = __fadd_rn(ot, in_blank);
            const float nt = det_lse2(nbk, nl);
            n_tot[i] = nt; n_blk[i] = nbk; n_lab[i] = nl;
            ont = ord(nt);
            s_akey[i] = ((u64)ont << 32) | (u64)(~(unsigned)i);
        }
        s_reset[i] = 0;
        if (tid == 0) { s_misc[0] = 0; s_misc[1] = 0; }
        // the classes beam i does not offer with its total: its children that are beams already, and its own label
        const unsigned* mycm = s_cm + i * CW;
        bool lb_child = false;
        int nex = 0;
        if (have) {
            for (int w = 0; w < CW; ++w) nex += __popc(mycm[w]);
            lb_child = lb >= 0 && ((mycm[lb >> 5] >> (lb & 31)) & 1u);
            if (lb >= 0 && !lb_child) nex += 1;
        }
        int extot;
        const int exo = block_scan(nex, &extot);   // (barrier: s_sin, s_akey, n_* visible)
        s_exoff[i] = exo;
        if (tid == 0) s_exoff[kSlots] = extot;
        const float lbscore = (have && lb >= 0 && !lb_child) ? __fadd_rn(s_in[lb], ob) : NEG;
        if (have) {
            int x = exo;
            for (int w = 0; w < CW; ++w) {
                unsigned bits = mycm[w];
                while (bits) {
                    const int e = w * 32 + __ffs(bits) - 1;
                    bits &= bits - 1;
                    s_exs[x++] = __fadd_rn(s_in[e], ot);
                }
            }
            if (lb >= 0 && !lb_child) s_exs[x++] = __fadd_rn(s_in[lb], ot);
        }
        s_lbs[i] = lbscore;
        // rank of a_i among the carried-over entries
        int pos = 0;
        if (have) {
            const u64 mykey = s_akey[i];
#pragma unroll 8
            for (int j = 0; j < nb; ++j) pos += s_akey[j] > mykey;
        }
        s_newpos[i] = pos;   // (scratch until the rebuild)
        __syncthreads();

        if (blockIdx.x == 0 && tid == 0) { const long long now = clock64(); g_beam_prof[1] += now - tprev; tprev = now; }
        // candidates of beam `bi` (total bt, own label offered with score bscore_lb, exclusions [excl0, excl1)) with a score
        // above Tq: branch-free binary search over the sorted class scores (fl(in + total) is monotone in `in`)
        int step0 = 1;
        while (step0 * 2 <= NCLS) step0 *= 2;
        auto cnt_beam = [&](float bt, float bscore_lb, int excl0, int excl1, unsigned Tq) -> int {
            int r = 0;
            for (int st = step0; st > 0; st >>= 1) {
                const int nx = r + st;
                if (nx <= NCLS && ord(__fadd_rn(s_sin[nx - 1], bt)) > Tq) r = nx;
            }
            for (int x = excl0; x < excl1; ++x) r -= ord(s_exs[x]) > Tq;
            r += ord(bscore_lb) > Tq;
            return r;
        };
        const int myex1 = exo + nex;
        auto cnt_mine = [&](unsigned Tq) -> int { return cnt_beam(ot, lbscore, exo, myex1, Tq); };
        // the same for seven thresholds at once (seven independent chains per step)
        auto cnt_mine7 = [&](const unsigned (&Tq)[7], int (&out)[7]) {
            int r[7];
#pragma unroll
            for (int q = 0; q < 7; ++q) r[q] = 0;
            for (int st = step0; st > 0; st >>= 1) {
#pragma unroll
                for (int q = 0; q < 7; ++q) {
                    const int nx = r[q] + st;
                    const float v = s_sin[min(nx, NCLS) - 1];
                    if (nx <= NCLS && ord(__fadd_rn(v, ot)) > Tq[q]) r[q] = nx;
                }
            }
            for (int x = exo; x < myex1; ++x) {
                const unsigned oe = ord(s_exs[x]);
#pragma unroll
                for (int q = 0; q < 7; ++q) r[q] -= oe > Tq[q];
            }
            const unsigned ol = ord(lbscore);
#pragma unroll
            for (int q = 0; q < 7; ++q) out[q] = r[q] + (ol > Tq[q]);
        };

        // ---- selection: V = the K-th largest score among the carried-over entries and the candidates of the beams that expand
        unsigned V = kNegInfOrd;
        bool allwin = false;
        int wq = 0, tq = 0, taq = 0;   // this beam's strict winners among its candidates, its candidates tied at V, a_i tied at V
        auto select = [&]() {
            const bool expand = have && !s_reset[i];
            const int c_all = expand ? cnt_mine(kNegInfOrd) : 0;
            const int total = nb + block_sum(c_all);
            if (total <= K) {
                allwin = true; V = kNegInfOrd; wq = c_all; tq = 0; taq = 0;
                return;
            }
            allwin = false;
            unsigned hiT = have ? max(ont, max(ord(__fadd_rn(s_sin[0], ot)), ord(lbscore))) : 0u;
            hiT = block_max_u(hiT);
            unsigned loT = kNegInfOrd;
            int gLo = total, gHi = 0;   // g(loT - 1) and g(hiT), g(T) = #entries with a score above T
            if (nb >= K) {
                // K real items bound the K-th largest from below: the carried-over entries (there are K of them)
                const unsigned lb1 = ~block_max_u(have ? ~ont : 0u);
                if (lb1 > loT) {
                    loT = lb1;
                    gLo = block_sum((have && ont > lb1 - 1u) + (expand ? cnt_mine(lb1 - 1u) : 0));
                }
            }
            if (hiT < loT) hiT = loT;
            while (loT < hiT) {   // smallest T in [loT, hiT] with g(T) < K, eight-way
                if (gLo >= 0 && gLo - gHi <= 2 * kSlots) break;   // few enough entries left in the bracket: rank them
                const unsigned wdt = hiT - loT;
                const unsigned step = max(1u, wdt >> 3);
                unsigned pq[7];
                int cq[7];
#pragma unroll
                for (int q = 0; q < 7; ++q) {
                    const unsigned long long pp = (unsigned long long)loT + (unsigned long long)(q + 1) * step - 1ull;
                    pq[q] = (unsigned)min(pp, (unsigned long long)(hiT - 1u));
                    cq[q] = 0;
                }
                if (expand) cnt_mine7(pq, cq);
#pragma unroll
                for (int q = 0; q < 7; ++q) cq[q] += have && ont > pq[q];
                int* r = s_red + (redph & 1) * (kCtaWarps * 8);
                redph++;
#pragma unroll
                for (int q = 0; q < 7; ++q) {
                    const int v = __reduce_add_sync(kFullMask, cq[q]);
                    if (lane == 0) r[warp * 8 + q] = v;
                }
                __syncthreads();
                unsigned nlo = loT, nhi = hiT;
                int nglo = gLo, nghi = gHi;
                bool found = false;
#pragma unroll
                for (int q = 0; q < 7; ++q) {
                    const int g = r[q] + r[8 + q] + r[16 + q] + r[24 + q];
                    if (!found) {
                        if (g < K) { nhi = pq[q]; nghi = g; found = true; }
                        else { nlo = pq[q] + 1u; nglo = g; }
                    }
                }
                loT = nlo; hiT = nhi; gLo = nglo; gHi = nghi;
            }
            if (loT < hiT) {
                // the entries with a score in [loT, hiT], gathered and ranked: V is the (K - g(hiT))-th largest of them
                if (tid == 0) s_misc[3] = 0;
                __syncthreads();
                if (have) {
                    if (ont >= loT && ont <= hiT) s_tmp[atomicAdd((int*)&s_misc[3], 1)] = ont;
                    if (expand) {
                        int r0 = 0, r1 = 0;   // sorted positions above hiT / at or above loT
                        for (int st = step0; st > 0; st >>= 1) {
                            const int n0 = r0 + st, n1 = r1 + st;
                            if (n0 <= NCLS && ord(__fadd_rn(s_sin[n0 - 1], ot)) > hiT) r0 = n0;
                            if (n1 <= NCLS && ord(__fadd_rn(s_sin[n1 - 1], ot)) >= loT) r1 = n1;
                        }
                        for (int r = r0; r < r1; ++r) {
                            const int cls = s_scls[r];
                            if (cls == lb || ((mycm[cls >> 5] >> (cls & 31)) & 1u)) continue;
                            const unsigned os = ord(__fadd_rn(s_sin[r], ot));
                            if (os > kNegInfOrd) s_tmp[atomicAdd((int*)&s_misc[3], 1)] = os;
                        }
                        const unsigned ol = ord(lbscore);
                        if (lb >= 0 && !lb_child && ol >= loT && ol <= hiT && ol > kNegInfOrd) s_tmp[atomicAdd((int*)&s_misc[3], 1)] = ol;
                    }
                }
                __syncthreads();
                const int n = s_misc[3], mth = K - gHi;
                __syncthreads();
                for (int q = tid; q < n; q += kCtaThreads) {
                    const unsigned v = s_tmp[q];
                    int gt = 0, ge = 0;
#pragma unroll 8
                    for (int u = 0; u < n; ++u) { const unsigned w = s_tmp[u]; gt += w > v; ge += w >= v; }
                    if (gt < mth && mth <= ge) s_misc[3] = (int)v;
                }
                __syncthreads();
                loT = (unsigned)s_misc[3];
                __syncthreads();
            }
            V = loT;
            wq = expand ? cnt_mine(V) : 0;
            tq = (expand && V > kNegInfOrd) ? cnt_mine(V - 1u) - wq : 0;
            taq = have && ont == V;
        };
        select();
        if (blockIdx.x == 0 && tid == 0) { const long long now = clock64(); g_beam_prof[2] += now - tprev; tprev = now; }

        // ---- resets: children whose carried-over entry has left the list by the time their parent comes to their class
        {
            // does a_i survive the selection above (computed with every beam expanding: the most candidates there can be)?
            int tot_ta;
            const int aw = have && (allwin || ont > V);
            const int nwin = block_sum(aw + wq);
            const int ta_before = block_scan(taq, &tot_ta);
            const bool a_in = aw || (taq && ta_before < K - nwin);
            if (blockIdx.x == 0 && tid == 0) { const long long now = clock64(); g_beam_prof[12] += now - tprev; tprev = now; }
            const int par = have ? s_ps[co + i] : -1;
            const bool ev_all = have && par >= 0 && par < i && !a_in && (long long)par * NCLS + lb + pos >= K;   // (at most par*NCLS + lb candidates come before)
            // First only the children that could put a candidate of their own into the list (best candidate at or above V): if
            // none of them loses its expansion, no reset can change the list, whatever happens to the others.  Only when one
            // fires are all of them evaluated and settled in time order.
            const bool relevant = have && (ord(__fadd_rn(s_sin[0], ot)) >= V || ord(lbscore) >= V);
            int nev = 0;
            for (int pass = 0; pass < 2; ++pass) {
                const bool ev = ev_all && (pass == 1 || relevant);
                const int eo = block_scan(ev ? 1 : 0, &nev);
                if (ev) s_evl[eo] = i;
                if (tid == 0) s_misc[0] = 0;
                __syncthreads();
                // G threads share an event: thread g takes every G-th class / exclusion / beam
                int G = 1;
                while (G * 2 * nev <= kCtaThreads && G < 32) G *= 2;
                if (tid < nev) s_cnt0[s_evl[tid]] = 0;
                __syncthreads();
                if (tid < nev * G) {
                    const int c = s_evl[tid / G], g = tid % G;
                    const unsigned S = ord(n_tot[c]);
                    const int p = s_ps[co + c], lab = s_label[co + c];
                    // candidates of the beams ahead of the parent with a score above S
                    int cnt = 0;
                    int pstep = 1;
                    while (pstep * 2 <= p) pstep *= 2;
                    for (int r = g; r < NCLS; r += G) {   // per class: how many of the p beams ahead (totals descending) clear S
                        const float v = s_sin[r];
                        int j = 0;
                        for (int st = pstep; st > 0; st >>= 1) {
                            const int nx = j + st;
                            if (nx <= p && ord(__fadd_rn(v, s_tot[co + nx - 1])) > S) j = nx;
                        }
                        if (j == 0) break;   // classes are in descending order: none of the later ones clears it either
                        cnt += j;
                    }
                    const int xe = s_exoff[p];
#pragma unroll 4
                    for (int x = g; x < xe; x += G) cnt -= ord(s_exs[x]) > S;
#pragma unroll 4
                    for (int q = g; q < p; q += G) cnt += ord(s_lbs[q]) > S;
                    // the parent's own candidates offered before the child's class
                    {
                        const float pt = s_tot[co + p], pbk = s_blk[co + p];
                        const int plb = s_label[co + p];
                        const unsigned* pcm = s_cm + p * CW;
#pragma unroll 4
                        for (int k = g; k < lab; k += G) {
                            const bool skip = (pcm[k >> 5] >> (k & 31)) & 1u;
                            cnt += !skip && ord(__fadd_rn(s_in[k], (k == plb) ? pbk : pt)) > S;
                        }
                    }
                    atomicAdd(&s_cnt0[c], cnt);
                }
                __syncthreads();
                if (tid < nev) {
                    const int c = s_evl[tid];
                    if (s_newpos[c] + s_cnt0[c] >= K) s_misc[0] = 1;
                }
                __syncthreads();
                if (blockIdx.x == 0 && tid == 0) { g_beam_prof[10 + pass] += nev; if (pass == 1) g_beam_prof[8] += 1; }
                if (!s_misc[0]) break;
            }
            if (blockIdx.x == 0 && tid == 0) { const long long now = clock64(); g_beam_prof[13] += now - tprev; tprev = now; }
            __syncthreads();
            if (s_misc[0]) {
                // One fired with every earlier beam expanding.  A reset removes candidates, so the true set is a subset: settle
                // the events that fired in time order (parent position, then class), taking out what the beams reset so far
                // would have offered.  The whole CTA works on one event at a time: thread q answers for beam q.
                const int c_e = tid < nev ? s_evl[tid] : 0;
                const bool fired = tid < nev && s_newpos[c_e] + s_cnt0[c_e] >= K;
                const unsigned key_e = fired ? (unsigned)s_ps[co + c_e] * (unsigned)C + (unsigned)s_label[co + c_e] : 0u;
                int nf;
                const int fo = block_scan(fired ? 1 : 0, &nf);
                if (fired) { s_tmp[2 * fo] = key_e; s_tmp[2 * fo + 1] = (unsigned)c_e; }
                __syncthreads();
                if (tid < nf) {
                    const unsigned mk = s_tmp[2 * tid];
                    int rk = 0;
                    for (int q = 0; q < nf; ++q) rk += s_tmp[2 * q] < mk;   // (parent, class) pairs are distinct
                    s_evl[rk] = (int)s_tmp[2 * tid + 1];
                }
                __syncthreads();
                int nres = 0;
                for (int f = 0; f < nf; ++f) {
                    const int cc = s_evl[f];
                    const int p = s_ps[co + cc];
                    if (s_reset[p]) continue;   // the parent does not expand: it never comes to the child's class
                    const unsigned S = ord(n_tot[cc]);
                    int adj = 0;
                    if (nres > 0) adj = block_sum((have && s_reset[i] && i < p) ? cnt_mine(S) : 0);
                    const bool out = s_newpos[cc] + s_cnt0[cc] - adj >= K;
                    __syncthreads();
                    if (out) {
                        if (tid == 0) s_reset[cc] = 1;
                        ++nres;
                    }
                    __syncthreads();
                }
                if (tid == 0) s_misc[1] = nres;
                __syncthreads();
                if (blockIdx.x == 0 && tid == 0) { const long long now = clock64(); g_beam_prof[14] += now - tprev; tprev = now; }
                if (blockIdx.x == 0 && tid == 0 && s_misc[1] > 0) g_beam_prof[9] += 1;
                if (s_misc[1] > 0) select();   // without the candidates of the beams that lost their expansion
            }
        }

        if (blockIdx.x == 0 && tid == 0) { const long long now = clock64(); g_beam_prof[3] += now - tprev; tprev = now; }
        // ---- winners, in push order: carried-over entries first, then the candidates beam by beam, class by class
        {
            const int aw = have && (allwin || ont > V);
            const int nwin = block_sum(aw + wq);
            int need = allwin ? 0 : K - nwin;        // places left for entries tied at V
            int tot_ta, tot_tc, tot_out;
            const int ta_before = block_scan(taq, &tot_ta);
            const int a_tie_in = taq && ta_before < need;
            need = max(0, need - tot_ta);
            const int tc_before = block_scan(tq, &tot_tc);
            const int quota = min(max(need - tc_before, 0), tq);
            const int nout = aw + a_tie_in + wq + quota;
            int o = block_scan(nout, &tot_out);
            if (have) {
                if (aw || a_tie_in) s_wkey[o++] = s_akey[i];
                if (!s_reset[i] && (wq + quota) > 0) {
                    const unsigned seq0 = (unsigned)nb + (unsigned)i * (unsigned)C;
                    // strict winners: the head of the sorted class list (the order inside the list of winners is irrelevant: it is
                    // ranked by key below), minus the classes this beam does not offer, plus its own label with its own score
                    int left = wq;
                    for (int r = 0; r < NCLS && left > 0; ++r) {
                        const unsigned os = ord(__fadd_rn(s_sin[r], ot));
                        if (!(os > V)) break;
                        const int cls = s_scls[r];
                        if (cls == lb || ((mycm[cls >> 5] >> (cls & 31)) & 1u)) continue;
                        s_wkey[o++] = ((u64)os << 32) | (u64)(~(seq0 + (unsigned)cls));
                        --left;
                    }
                    if (lb >= 0 && !lb_child && ord(lbscore) > V) s_wkey[o++] = ((u64)ord(lbscore) << 32) | (u64)(~(seq0 + (unsigned)lb));
                    // entries tied at V: the first `quota` in class order
                    int taken = 0;
                    for (int k = 0; k < NCLS && taken < quota; ++k) {
                        if ((mycm[k >> 5] >> (k & 31)) & 1u) continue;
                        const unsigned os = ord(__fadd_rn(s_in[k], (k == lb) ? ob : ot));
                        if (os == V) { s_wkey[o++] = ((u64)os << 32) | (u64)(~(seq0 + (unsigned)k)); ++taken; }
                    }
                }
            }
            __syncthreads();
            // rank by counting (keys are distinct), empty places last
            const int nw = tot_out;
            u64 mk = 0;
            int rk = tid;
            if (tid < nw) {
                mk = s_wkey[tid];
                rk = 0;
#pragma unroll 8
                for (int q = 0; q < nw; ++q) rk += s_wkey[q] > mk;
            }
            s_skey[rk] = mk;
            s_newpos[tid] = -1;
            __syncthreads();
        }

        if (blockIdx.x == 0 && tid == 0) { const long long now = clock64(); g_beam_prof[4] += now - tprev; tprev = now; }
        // ---- rebuild the beam state from the sorted list (as in the kernel above; thread p builds slot p)
        {
            const int p = tid;
            const u64 kk = s_skey[p];
            const bool live = (kk >> 32) != 0;
            const unsigned seq = ~(unsigned)kk;
            const bool carried = live && seq < (unsigned)nb;
            int src = -1, kcls = -1;   // the old slot this entry continues / is a child of
            if (live) {
                if (carried) {
                    src = (int)seq;
                    s_tot[no + p] = n_tot[src]; s_blk[no + p] = n_blk[src]; s_labp[no + p] = n_lab[src];
                    s_label[no + p] = s_label[co + src];
                    s_hash[no + p] = s_hash[co + src]; s_phash[no + p] = s_phash[co + src];
                    s_pool[no + p] = s_pool[co + src];
                    s_newpos[src] = p;
                } else {
                    const unsigned cc = seq - (unsigned)nb;
                    src = (int)(cc / (unsigned)C);
                    kcls = (int)(cc - (unsigned)src * (unsigned)C);
                    const float sc = unord((unsigned)(kk >> 32));
                    s_tot[no + p] = sc; s_blk[no + p] = NEG; s_labp[no + p] = sc;
                    s_label[no + p] = kcls;
                    const u64 ph = s_hash[co + src];
                    s_phash[no + p] = ph;
                    s_hash[no + p] = mix_hash(ph, kcls);
                    const int id = 1 + t * K + p;
                    s_pool[no + p] = id;
                    pool[id] = make_int2(s_pool[co + src], kcls);
                }
            }
            const int nb_new = block_sum(live ? 1 : 0);   // (barrier: s_newpos and the new hashes are visible)
            bool orphan = false;   // a surviving beam whose parent was not in the beam
            if (live) {
                int psn;
                if (carried) {
                    const int po = s_ps[co + src];
                    psn = (po >= 0) ? s_newpos[po] : -1;
                    orphan = po < 0 && s_label[co + src] >= 0;
                } else {
                    psn = s_newpos[src];
                }
                // a surviving beam whose parent prefix was re-created in this frame gets its parent back
                if (orphan) {
                    const u64 ph = s_phash[no + p];
                    for (int q = 0; q < kSlots; ++q) {
                        const u64 kq = s_skey[q];
                        if ((kq >> 32) != 0 && (~(unsigned)kq) >= (unsigned)nb && s_hash[no + q] == ph) psn = q;
                    }
                }
                s_ps[no + p] = psn;
            }
            for (int q = tid; q < kSlots * CW; q += kCtaThreads) s_cm[q] = 0;
            __syncthreads();
            if (live) {
                const int psn = s_ps[no + p];
                const int lbn = s_label[no + p];
                if (psn >= 0 && lbn >= 0) atomicOr(&s_cm[psn * CW + (lbn >> 5)], 1u << (lbn & 31));
            }
            if (tid == 0) s_misc[2] = nb_new;
        }
        __syncthreads();
        nb = s_misc[2];
        cur ^= 1;
        if (blockIdx.x == 0 && tid == 0) { const long long now = clock64(); g_beam_prof[5] += now - tprev; tprev = now; }
    }

    // ---- TopPaths: beams are sorted; walk the back-pointers of the first top_paths
    __threadfence_block();
    __syncthreads();
    const int co = cur * kSlots;
    for (int p = tid; p < top_paths; p += kCtaThreads) {
        int64_t* out = decoded + ((size_t)b * top_paths + p) * T;
        int n = 0;
        float lp = NEG;  // TF raises when fewer leaves than requested paths exist
        if (p < nb) {
            lp = s_tot[co + p];
            int id = s_pool[co + p];
            int prev = -1;
            while (id > 0) {  // collect from the leaf towards the root, stored from the end of the row
                const int2 nd = pool[id];
                if (!merge_repeated || nd.y != prev) { out[T - 1 - n] = nd.y; ++n; }
                prev = nd.y;
                id = nd.x;
            }
            for (int q = 0; q < n; ++q) out[q] = out[T - n + q];
        }
        for (int q = n; q < T; ++q) out[q] = -1;
        decoded_len[(size_t)b * top_paths + p] = n;
        log_prob[(size_t)b * top_paths + p] = lp;
    }
}

}  // namespace ocr

using namespace ocr;

static int g_beam_path = 0;   // 0: CTA per sequence (default), 1: warp per sequence (the replay of TensorFlow's list updates)
extern "C" int ocr_debug_beam_path(int path) {
    OCR_CHECK_ARG(path == 0 || path == 1, "ocr_debug_beam_path: path=%d", path);
    g_beam_path = path;
    return OCR_OK;
}

extern "C" int ocr_debug_beam_profile(long long* host16, int reset) {
    long long z[16] = {0};
    if (host16) OCR_CHECK_CUDA(cudaMemcpyFromSymbol(host16, g_beam_prof, sizeof(z)));
    if (reset) OCR_CHECK_CUDA(cudaMemcpyToSymbol(g_beam_prof, z, sizeof(z)));
    return OCR_OK;
}

extern "C" int ocr_ctc_beam_search_workspace_bytes(int T, int B, int C, int beam_width, size_t* bytes)
{
    OCR_CHECK_ARG(bytes != nullptr, "ocr_ctc_beam_search_workspace_bytes: bytes is NULL");
    OCR_CHECK_ARG(T >= 1 && B >= 0 && C >= 2 && beam_width >= 1 && beam_width <= kSlots,
                  "ocr_ctc_beam_search_workspace_bytes: bad shape T=%d B=%d C=%d beam_width=%d (max %d)", T, B, C, beam_width, kSlots);
    *bytes = (size_t)B * (1 + (size_t)T * beam_width) * sizeof(int2);
    return OCR_OK;
}

extern "C" int ocr_ctc_beam_search(const float* logits, int T, int B, int C, const int32_t* seq_len, int beam_width,
                                   int top_paths, int merge_repeated, int normalize, int64_t* decoded,
                                   int32_t* decoded_len, float* log_prob, void* workspace, size_t workspace_bytes,
                                   ocr_stream_t stream)
{
    OCR_CHECK_ARG(T >= 1 && B >= 0 && C >= 2, "ocr_ctc_beam_search: bad shape T=%d B=%d C=%d", T, B, C);
    OCR_CHECK_ARG(beam_width >= 1 && beam_width <= kSlots, "ocr_ctc_beam_search: beam_width=%d outside [1,%d]", beam_width, kSlots);
    OCR_CHECK_ARG(top_paths >= 1 && top_paths <= beam_width, "ocr_ctc_beam_search: top_paths=%d must be in [1,beam_width]", top_paths);
    OCR_CHECK_ARG(C <= 512, "ocr_ctc_beam_search: C=%d > 512 unsupported", C);
    OCR_CHECK_ARG((long long)beam_width * C + beam_width < 0x7fffffffLL && (long long)T * beam_width < 0x7fffff00LL,
                  "ocr_ctc_beam_search: problem too large");
    if (B == 0) return OCR_OK;
    OCR_CHECK_ARG(logits && seq_len && decoded && decoded_len && log_prob, "ocr_ctc_beam_search: NULL argument");
    const size_t need = (size_t)B * (1 + (size_t)T * beam_width) * sizeof(int2);
    if (workspace == nullptr || workspace_bytes < need) {
        set_error("ocr_ctc_beam_search: workspace too small (%zu < %zu)", workspace_bytes, need);
        return OCR_EWORKSPACE;
    }
    if (g_beam_path == 0) {
        const CtaLayout CL = cta_layout(C);
        OCR_CHECK_ARG(CL.total <= kMaxDynSmem, "ocr_ctc_beam_search: shared memory %d too large", CL.total);
        static int configured_cta = -1;
        int dev_cta = 0;
        OCR_CHECK_CUDA(cudaGetDevice(&dev_cta));
        if (configured_cta != dev_cta) {
            OCR_CHECK_CUDA(cudaFuncSetAttribute(ctc_beam_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
            configured_cta = dev_cta;
        }
        ctc_beam_cta_kernel<<<B, kCtaThreads, CL.total, static_cast<cudaStream_t>(stream)>>>(
            logits, T, B, C, seq_len, beam_width, top_paths, merge_repeated, normalize, decoded, decoded_len, log_prob,
            static_cast<int2*>(workspace));
        OCR_CHECK_LAUNCH();
        return OCR_OK;
    }
    const BeamLayout L = beam_layout(C);
    const int smem = L.per_warp * kBeamWarps;
    OCR_CHECK_ARG(smem <= kMaxDynSmem, "ocr_ctc_beam_search: shared memory %d too large", smem);
    static int configured = -1;
    int dev = 0;
    OCR_CHECK_CUDA(cudaGetDevice(&dev));
    if (configured != dev) {
        OCR_CHECK_CUDA(cudaFuncSetAttribute(ctc_beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
        configured = dev;
    }
    const int grid = (B + kBeamWarps - 1) / kBeamWarps;
    ctc_beam_kernel<<<grid, kBeamWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
        logits, T, B, C, seq_len, beam_width, top_paths, merge_repeated, normalize, decoded, decoded_len, log_prob,
        static_cast<int2*>(workspace));
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}
