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

}  // namespace ocr

using namespace ocr;

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
