// Internal interface of the tcgen05 TF32 GEMM (gemm_tf32.cu): plan once (two TMA tensor maps), run many times.
// The recurrent layers re-launch the same h * W_h^T product every frame with unchanged buffers.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace ocr {

struct GemmPlan {
    CUtensorMap tmA, tmB;
    CUtensorMap tmD;       // output [M, N] as 32-column x 32-row boxes, 128-byte swizzle (tma_store)
    int tma_store = 0;     // 1: the epilogue stages 32x32 blocks in shared memory and stores them with cp.async.bulk.tensor (plain products with 16-byte aligned output rows)
    const float* bias;
    float* D;
    int M, N, K, ldd, relu, bn;
    // split-K / shifted-operand form used by the weight-gradient contractions (gemm_plan_wgrad):
    //   D[batch][split] (M x N) = sum over the split's k range of A[m, k + a_shift[batch]] * W[n, k]
    int splits = 1, ksteps_per_split = 0, nbatch = 1, a_box_rows = 128;
    int h16 = 0;             // 1: bfloat16 operands (gemm_plan_dirs_h16), 2: binary16 (gemm_plan_f16); K counted in elements of 64 per 128-byte row
    int tp = 1, ntaps = 0;   // stacked views (gemm_wgrad, shallow layers): tp views of a_box_rows rows per A tile, ntaps views in all
    int pdl = 0;       // 1: launch with programmatic stream serialization (the kernel orders itself with griddepcontrol.wait after its prologue)
    int blocked = 0;   // 1: K-blocked operands (gemm_wgrad, blocked != 0): 3-D tensor maps [K/32][rows][32], a_shift in blocks of 32
    int shallow = 0;   // 1: take the shallow pipeline (small shared-memory footprint) even for a one-wave grid, to leave room for co-running kernels
    int a_shift[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, a_row[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, b_row[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long split_stride = 0, batch_stride = 0;   // elements between the partial outputs of consecutive splits / batches
};

// D[M,N] = act(A[M,K] * W[N,K]^T + bias[N]); see ocr_gemm_tf32 in include/ocr_b200.h for the operand rules
int gemm_plan(GemmPlan* p, const float* A, int lda, const float* W, int ldw, const float* bias, float* D, int ldd, int M, int N,
              int K, int relu);
int gemm_run(const GemmPlan& p, cudaStream_t st);
// ndir independent products in one launch: D[d][split] (M x N dense) = A[d*M.., :] * W[d*N.., :]^T over the split's k range
int gemm_plan_dirs(GemmPlan* p, const float* A, int lda, const float* W, int ldw, float* D, int M, int N, int K, int ndir, int splits, int bn);
// the same with bfloat16 operands (gradients need float32's exponent range): A [ndir*M, K], W [ndir*N, K] bfloat16, float32 sums
// (tcgen05.mma.kind::f16, K = 16 per instruction; half the operand bytes through L2 -> shared memory); bn is 32 or 64
int gemm_plan_dirs_h16(GemmPlan* p, const void* A, int lda, const void* W, int ldw, float* D, int M, int N, int K, int ndir, int splits, int bn);
// binary16 operands, otherwise gemm_plan (bias, ReLU, TMA-store epilogue)
int gemm_plan_f16(GemmPlan* p, const void* A, int lda, const void* W, int ldw, const float* bias, float* D, int ldd, int M, int N, int K, int relu);
// Weight-gradient contraction over a long row dimension R (both operands R-contiguous, i.e. transposed activations):
//   D[batch] (M x N, row pitch ldd, batches batch_stride apart) = sum_r A[a_row[batch] + m, r + a_shift[batch]] * W[n, r]
// (A has a_rows rows in total; a_shift must be a multiple of 4: TMA box origins are 16-byte aligned)
// split over the CTAs along R; partial tiles go to `partials` ([nbatch][splits][M][N], caller scratch) and a second
// kernel adds them up in a fixed order (deterministic).  Out-of-range r (negative too) reads as zero (TMA fill).
size_t gemm_wgrad_scratch_floats(int M, int N, long long R, int nbatch);
int gemm_wgrad(const float* A, long long lda, const float* W, long long ldw, float* D, int ldd, long long batch_stride, int M, int N,
               long long R, int nbatch, const int* a_shift, const int* a_row, long long a_rows, float* partials, cudaStream_t st,
               int blocked = 0, long long w_rows = 0);
// blocked != 0: the operands are K-BLOCKED -- element (row, r) of A lies at ((r / 32) * a_rows + row) * 32 + r % 32 (W likewise
// with w_rows rows), R and every a_shift are multiples of 32, lda / ldw are unused.  The 128 bytes x box_rows a k-step needs of
// an operand are then ONE contiguous run of memory instead of box_rows runs a whole row pitch apart (channel planes of the
// shallow layers lie 2-8 MB apart: a k-step touched 160 pages, and conv2's filter gradient moved 25 GB/s per SM).
int tma_map_blocked(CUtensorMap* tm, const float* base, long long rows, long long nblk, int box_rows);

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 32;  // fp32 elements = one 128-byte swizzle row

// 2-D fp32 tensor [rows, K] row-major with row pitch ld elements; box = [box_rows, 32 floats], 128-byte swizzle
int tma_map_2d(CUtensorMap* tm, const float* base, long long rows, long long K, long long ld, int box_rows);

// 3-D view [K/32][rows][32]: one request loads box_chunks consecutive swizzled k-chunk tiles of box_rows rows
int tma_map_chunks(CUtensorMap* tm, const float* base, long long rows, long long K, long long ld, int box_rows, int box_chunks);
// output tensor [rows, cols] fp32 with row pitch ld elements; box = [32 rows, 32 floats], 128-byte swizzle (TMA-store epilogues)
int tma_map_out(CUtensorMap* tm, float* base, long long rows, long long cols, long long ld);

// the same two views of a half-precision (IEEE binary16) matrix: a 128-byte swizzle row holds 64 elements
constexpr int kGemmBKh = 64;
int tma_map_2d_h(CUtensorMap* tm, const void* base, long long rows, long long K, long long ld, int box_rows);
int tma_map_chunks_h(CUtensorMap* tm, const void* base, long long rows, long long K, long long ld, int box_rows, int box_chunks);

#ifdef __CUDACC__
__device__ __forceinline__ unsigned g_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void g_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void g_mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void g_mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    for (unsigned it = 0; it < (1u << 28); ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* tm, int c0, int c1, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* tm, int c0, int c1, int c2, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, rows of 128 bytes, 8-row atoms 1024 bytes apart
__device__ __forceinline__ unsigned long long umma_desc_k128(unsigned smem_addr) {
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr >> 4) & 0x3FFF);
    d |= (unsigned long long)1 << 16;               // leading byte offset (unused for swizzled K-major)
    d |= (unsigned long long)(1024 >> 4) << 32;     // stride byte offset: next 8-row atom
    d |= (unsigned long long)1 << 46;               // descriptor version (sm_100)
    d |= (unsigned long long)2 << 61;               // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
// binary16 operands (K = 16 per instruction: the same 32 bytes of a swizzle row as eight tf32), float32 accumulation
__device__ __forceinline__ void umma_f16(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int c0, int c1, unsigned src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(src), "r"(c0), "r"(c1) : "memory");
}

// TMA-store epilogue: accumulator tile -> + bias, ReLU -> 32x32 blocks staged in shared memory (128-byte rows, 16-byte groups
// XOR (row & 7): the SWIZZLE_128B layout, conflict-free for a warp writing one row per lane) -> cp.async.bulk.tensor stores.
// Each epilogue warp owns its 32 rows and two 4 KB staging blocks at `slab` (1024-byte aligned; the CTA's pipeline stages,
// idle once the accumulator is complete): no CTA-wide barrier.  Against one 16-byte STG per lane and row (32 rows x 16 bytes
// per instruction, every 128-byte line touched by eight instructions) the stores leave as whole lines.  Rows >= M and columns
// >= N are clipped by the tensor map.
template <int BN>
__device__ __forceinline__ void gemm_epilogue_tma(unsigned tmem_d, int q, int lane, int m0, int n0, int N, const float* __restrict__ bias,
                                                  int relu, const CUtensorMap* tmD, unsigned slab)
{
    const unsigned my = slab + (unsigned)q * 8192u;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
        if (n0 + c0 >= N) break;
        unsigned r[32];
        const unsigned taddr = tmem_d + ((unsigned)(q * 32) << 16) + (unsigned)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const unsigned blk = my + (unsigned)((c0 >> 5) & 1) * 4096u;
        if (c0 >= 64) {   // the block is being reused: the store issued two steps ago must have read it
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(r[j]);
            const int col = n0 + c0 + j;
            if (bias != nullptr && col < N) x += __ldg(bias + col);
            v[j] = relu ? fmaxf(x, 0.0f) : x;
        }
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(blk + (unsigned)lane * 128u + (unsigned)((j4 ^ (lane & 7)) << 4)),
                         "f"(v[4 * j4]), "f"(v[4 * j4 + 1]), "f"(v[4 * j4 + 2]), "f"(v[4 * j4 + 3]) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async-proxy (TMA) reads
        __syncwarp();
        if (lane == 0) {
            tma_store_2d(tmD, n0 + c0, m0 + q * 32, blk);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory must outlive the reads
    __syncwarp();
}

// accumulator tile (TMEM lanes q*32.., BN columns) -> + bias, ReLU -> global rows; called by the four epilogue warps
template <int BN>
__device__ __forceinline__ void gemm_epilogue(unsigned tmem_d, int q, int lane, int m0, int n0, int M, int N,
                                              const float* __restrict__ bias, float* __restrict__ D, int ldd, int relu)
{
    const int row = m0 + q * 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
        unsigned r[32];
        const unsigned taddr = tmem_d + ((unsigned)(q * 32) << 16) + (unsigned)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < M) {
            float* drow = D + (size_t)row * ldd + n0 + c0;
            const bool vec = ((reinterpret_cast<uintptr_t>(drow) & 15) == 0) && (n0 + c0 + 32 <= N);
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float x = __uint_as_float(r[j]);
                const int col = n0 + c0 + j;
                if (bias != nullptr && col < N) x += __ldg(bias + col);
                v[j] = relu ? fmaxf(x, 0.0f) : x;
            }
            if (vec) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(drow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (n0 + c0 + j < N) drow[j] = v[j];
            }
        }
    }
}

#endif

}  // namespace ocr
