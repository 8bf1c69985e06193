// TF32 tensor-core GEMM with fused bias / ReLU epilogue for sm_100a:  D[M,N] = act(A[M,K] * W[N,K]^T + bias[N]).
//
// The dense contractions of the recognizer -- the 3x3 convolutions as (im2col patches) x (filters)
// (/root/reference/src/weinman/model.py:84-109, tf.layers.conv2d), the LSTM/GRU input projections
// (model.py:167-199, the [x, h] * kernel product of tf.contrib.rnn cells split into its x part) and the
// logits layer (model.py:216-220, tf.layers.dense + ReLU) -- all have this shape with K-major operands.
//
// Blackwell mapping (one 128 x BN output tile per CTA, 6 warps, warp specialised):
//   warp 0 / one lane : TMA producer.  cp.async.bulk.tensor.2d loads a 128x32 fp32 box of A and a BNx32 box of W
//                       per k-step into a 128B-swizzled shared-memory stage; completion on the stage's mbarrier.
//   warp 1            : allocates BN TMEM columns; one lane issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8, four per
//                       k-step, shared-memory descriptors, fp32 accumulator in TMEM) and tcgen05.commit's the stage
//                       back to the producer; the last commit signals the epilogue.
//   warps 2..5        : epilogue.  tcgen05.ld (32 lanes x 32 columns per warp) -> registers -> + bias, ReLU ->
//                       128-byte row segments to global memory.
// Out-of-range rows / columns / k are zero-filled by TMA, so M, N, K need no padding (K*4 bytes must be a
// multiple of 16 for the tensor map).  TF32 keeps fp32 storage end to end: activations and weights stay the
// reference's float32 tensors, products are rounded to 10-bit mantissas inside the tensor core, sums are fp32.
#include <cuda_fp16.h>
#include "gemm_tf32.cuh"

namespace ocr {

constexpr int kGemmThreads = 192;

struct GemmSplit {
    int ksteps_per_split, nbatch;
    int a_shift[9], a_row[9], b_row[9];
    int a_box_bytes;   // bytes one A box delivers (boxes shorter than 128 rows when M < 128: no over-fetch of foreign rows)
    int tma_store;     // 1: TMA-store epilogue through tmD
    int tp, ntaps;     // tp > 1: a tile stacks the a_box_rows-row slices of tp consecutive views (of ntaps in all) in its 128 rows
    int blocked;       // 1: K-blocked operands, 3-D tensor maps [K/32][rows][32]; a_shift counts blocks
    long long split_stride, batch_stride;
};

template <int BN, int STAGES>
struct GemmSmem {
    static constexpr int kA = kGemmBM * kGemmBK * 4;
    static constexpr int kB = BN * kGemmBK * 4;
    static constexpr int kStage = kA + kB;
    static constexpr int kBars = STAGES * kStage;
    static constexpr int kTotal = kBars + (2 * STAGES + 1) * 8 + 16 + 1024;  // + alignment slack
};

template <int BN, int STAGES, int H16 = 0>   // H16: 16-bit operands, 64 elements per swizzle row (1: bfloat16, 2: binary16)
__global__ void __launch_bounds__(kGemmThreads)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmD,
                 const float* __restrict__ bias, float* __restrict__ D, int M, int N, int K, int ldd, int relu,
                 const GemmSplit sp)
{
    using S = GemmSmem<BN, STAGES>;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // (programmatic dependent launch: the successor may be scheduled; see below)
    extern __shared__ unsigned char gemm_smem_raw[];
    // 128-byte swizzle wants 1024-byte aligned tiles
    unsigned char* smem = gemm_smem_raw + ((1024u - (g_smem_u32(gemm_smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    const unsigned s_base = g_smem_u32(smem);
    const unsigned bar_full = s_base + S::kBars;
    const unsigned bar_empty = bar_full + STAGES * 8;
    const unsigned bar_acc = bar_empty + STAGES * 8;
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(smem + S::kBars + (2 * STAGES + 1) * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // blockIdx.x = m-tile * nbatch + batch (the batches of one tile run side by side and share operand tiles in L2)
    const int batch = blockIdx.x % sp.nbatch;
    const int m0 = (blockIdx.x / sp.nbatch) * kGemmBM, n0 = blockIdx.y * BN;
    constexpr int kBKe = H16 ? kGemmBKh : kGemmBK;   // elements per k-step (one 128-byte row)
    const int nk_all = (K + kBKe - 1) / kBKe;
    const int k_begin = sp.ksteps_per_split > 0 ? blockIdx.z * sp.ksteps_per_split : 0;
    const int nk = sp.ksteps_per_split > 0 ? min(sp.ksteps_per_split, nk_all - k_begin) : nk_all;
    // Stacked views (weight gradients of the shallow layers, C_in <= 64): the A tile's 128 rows are tp slices of a_box_rows rows,
    // one per tap view, each its own TMA box; ONE MMA then contracts tp taps against the same d-output tile (a tile per tap
    // left 3/4 of the MMA rows and of the B-operand traffic unused: conv2's filter gradient ran 1184 us).
    const int tp = sp.tp;
    const int v0 = batch * tp;                                            // first view of this tile
    const int nview = tp > 1 ? min(tp, sp.ntaps - v0) : 1;
    const int b_row = sp.b_row[v0];
    D += (size_t)batch * sp.batch_stride + (size_t)blockIdx.z * sp.split_stride;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { g_mbar_init(bar_full + s * 8, 1); g_mbar_init(bar_empty + s * 8, 1); }
        g_mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(g_smem_u32(tmem_slot)), "r"((unsigned)BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_d = *tmem_slot;
    // Programmatic dependent launch (GemmPlan::pdl): the prologue above (barriers, TMEM) may have run while the predecessor in
    // the stream was still finishing; nothing below touches global memory before the predecessor's results are visible.
    // (No-ops in an ordinary launch.)
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0) {
        if (lane == 0) {
            for (int k = 0; k < nk; ++k) {
                const int s = k % STAGES;
                if (k >= STAGES) g_mbar_wait(bar_empty + s * 8, ((k / STAGES) - 1) & 1);
                g_mbar_expect_tx(bar_full + s * 8, (unsigned)(nview * sp.a_box_bytes + S::kB));
                if (sp.blocked) {     // K-blocked operands: a box is one contiguous run of box_rows x 128 bytes
                    for (int i = 0; i < nview; ++i)
                        tma_load_3d(s_base + s * S::kStage + i * sp.a_box_bytes, &tmA, 0, m0 + sp.a_row[v0 + i], k_begin + k + sp.a_shift[v0 + i],
                                    bar_full + s * 8);
                    tma_load_3d(s_base + s * S::kStage + S::kA, &tmB, 0, n0 + b_row, k_begin + k, bar_full + s * 8);
                    continue;
                }
                for (int i = 0; i < nview; ++i)
                    tma_load_2d(s_base + s * S::kStage + i * sp.a_box_bytes, &tmA, (k_begin + k) * kBKe + sp.a_shift[v0 + i], m0 + sp.a_row[v0 + i],
                                bar_full + s * 8);
                tma_load_2d(s_base + s * S::kStage + S::kA, &tmB, (k_begin + k) * kBKe, n0 + b_row, bar_full + s * 8);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: fp32 accumulate, tf32 x tf32 (H16: bfloat16 x bfloat16 -- measured: mixing bfloat16 and
            // binary16 operands in one tcgen05.mma.kind::f16 is an illegal instruction), both K-major, N = BN, M = 128
            // (operand format field: kind::tf32 2 = TF32; kind::f16 1 = bfloat16, 0 = binary16)
            constexpr unsigned kFmt = H16 == 0 ? 2u : (H16 == 1 ? 1u : 0u);
            const unsigned idesc = (1u << 4) | (kFmt << 7) | (kFmt << 10) | ((unsigned)(BN >> 3) << 17) | ((unsigned)(kGemmBM >> 4) << 24);
            for (int k = 0; k < nk; ++k) {
                const int s = k % STAGES;
                g_mbar_wait(bar_full + s * 8, (k / STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned a_addr = s_base + s * S::kStage, b_addr = a_addr + S::kA;
                const unsigned long long da = umma_desc_k128(a_addr), db = umma_desc_k128(b_addr);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {  // 8 tf32 / 16 halves = 32 bytes per MMA: advance the start address inside the swizzle row
                    if constexpr (H16) umma_f16(tmem_d, da + (unsigned long long)(kk * 2), db + (unsigned long long)(kk * 2), idesc, (k | kk) ? 1u : 0u);
                    else umma_tf32(tmem_d, da + (unsigned long long)(kk * 2), db + (unsigned long long)(kk * 2), idesc, (k | kk) ? 1u : 0u);
                }
                umma_commit(bar_empty + s * 8);  // frees the stage when the MMAs that read it retire
            }
            umma_commit(bar_acc);
        }
    } else {
        // epilogue warps 2..5: TMEM lane quarter = warp % 4
        const int q = warp & 3;
        g_mbar_wait(bar_acc, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // (the pipeline stages are idle by now: every TMA load has been consumed and every MMA has retired)
        if (sp.tma_store) gemm_epilogue_tma<BN>(tmem_d, q, lane, m0, n0, N, bias, relu, &tmD, s_base);
        else gemm_epilogue<BN>(tmem_d, q, lane, m0, n0, tp > 1 ? nview * (sp.a_box_bytes >> 7) : M, N, bias, D, ldd, relu);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((unsigned)BN) : "memory");
    }
}

}  // namespace ocr

using namespace ocr;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D fp32 tensor [rows, K] row-major with row pitch ld elements; box = [box_rows, 32 floats], 128-byte swizzle
namespace ocr {
static int g_tma_promo256 = 0;   // tuning: L2 promotion of the 2-D operand maps (0: 128 bytes, 1: 256 bytes)
int tma_map_2d(CUtensorMap* tm, const float* base, long long rows, long long K, long long ld, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return OCR_ECUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)kGemmBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     g_tma_promo256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld K=%lld ld=%lld", (int)r, rows, K, ld); return OCR_ECUDA; }
    return OCR_OK;
}
}  // namespace ocr

namespace ocr {
// K-blocked fp32 operand [nblk][rows][32]: box = [1][box_rows][32 floats] = box_rows x 128 contiguous bytes, 128-byte swizzle
int tma_map_blocked(CUtensorMap* tm, const float* base, long long rows, long long nblk, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return OCR_ECUDA; }
    cuuint64_t dims[3] = {(cuuint64_t)kGemmBK, (cuuint64_t)rows, (cuuint64_t)nblk};
    cuuint64_t strides[2] = {(cuuint64_t)kGemmBK * 4, (cuuint64_t)rows * kGemmBK * 4};
    cuuint32_t box[3] = {(cuuint32_t)kGemmBK, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (blocked) failed (%d) rows=%lld nblk=%lld", (int)r, rows, nblk); return OCR_ECUDA; }
    return OCR_OK;
}
}  // namespace ocr

// 3-D view of a [rows, K] fp32 matrix as [K/32 chunks][rows][32 floats]: ONE request fetches `box_chunks` consecutive
// 128-byte-swizzled k-chunk tiles of `box_rows` rows (each tile laid out exactly as tma_map_2d's boxes)
namespace ocr {
int tma_map_chunks(CUtensorMap* tm, const float* base, long long rows, long long K, long long ld, int box_rows, int box_chunks) {
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return OCR_ECUDA; }
    cuuint64_t dims[3] = {(cuuint64_t)kGemmBK, (cuuint64_t)rows, (cuuint64_t)(K / kGemmBK)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)kGemmBK * 4};
    cuuint32_t box[3] = {(cuuint32_t)kGemmBK, (cuuint32_t)box_rows, (cuuint32_t)box_chunks};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (3-D) failed (%d) rows=%lld K=%lld ld=%lld", (int)r, rows, K, ld); return OCR_ECUDA; }
    return OCR_OK;
}
}  // namespace ocr

// binary16 forms of the two views (64 elements per 128-byte swizzle row)
namespace ocr {
int tma_map_2d_h(CUtensorMap* tm, const void* base, long long rows, long long K, long long ld, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return OCR_ECUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)kGemmBKh, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (binary16) failed (%d) rows=%lld K=%lld ld=%lld", (int)r, rows, K, ld); return OCR_ECUDA; }
    return OCR_OK;
}
int tma_map_chunks_h(CUtensorMap* tm, const void* base, long long rows, long long K, long long ld, int box_rows, int box_chunks) {
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return OCR_ECUDA; }
    cuuint64_t dims[3] = {(cuuint64_t)kGemmBKh, (cuuint64_t)rows, (cuuint64_t)(K / kGemmBKh)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)kGemmBKh * 2};
    cuuint32_t box[3] = {(cuuint32_t)kGemmBKh, (cuuint32_t)box_rows, (cuuint32_t)box_chunks};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (3-D, binary16) failed (%d) rows=%lld K=%lld ld=%lld", (int)r, rows, K, ld); return OCR_ECUDA; }
    return OCR_OK;
}
}  // namespace ocr

namespace ocr {
int tma_map_out(CUtensorMap* tm, float* base, long long rows, long long cols, long long ld) {
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return OCR_ECUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (output) failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, rows, cols, ld); return OCR_ECUDA; }
    return OCR_OK;
}
static int g_gemm_tma_store = 1;   // TMA-store epilogue where the plan allows it (0: one STG per lane and row everywhere)
int gemm_set_tma_store(int on) {
    g_gemm_tma_store = on ? 1 : 0;
    return OCR_OK;
}
}  // namespace ocr

template <int BN, int STAGES, int H16 = 0>
static int launch_planned(const GemmPlan& p, cudaStream_t st)
{
    using S = GemmSmem<BN, STAGES>;
    static int configured = -1;
    int dev = 0;
    OCR_CHECK_CUDA(cudaGetDevice(&dev));
    if (configured != dev) {
        OCR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<BN, STAGES, H16>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
        configured = dev;
    }
    dim3 grid(((p.M + kGemmBM - 1) / kGemmBM) * p.nbatch, (p.N + BN - 1) / BN, p.splits);
    GemmSplit sp;
    sp.ksteps_per_split = p.splits > 1 ? p.ksteps_per_split : 0;
    sp.nbatch = p.nbatch;
    for (int i = 0; i < 9; ++i) { sp.a_shift[i] = p.a_shift[i]; sp.a_row[i] = p.a_row[i]; sp.b_row[i] = p.b_row[i]; }
    sp.split_stride = p.split_stride;
    sp.a_box_bytes = p.a_box_rows * kGemmBK * 4;
    sp.batch_stride = p.batch_stride;
    sp.tma_store = (p.tma_store && g_gemm_tma_store && p.nbatch == 1 && p.splits == 1) ? 1 : 0;
    sp.tp = p.tp;
    sp.ntaps = p.ntaps;
    sp.blocked = p.blocked;
    if (p.pdl) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(kGemmThreads);
        cfg.dynamicSmemBytes = S::kTotal;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        OCR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tf32_kernel<BN, STAGES, H16>, p.tmA, p.tmB, p.tmD, p.bias, p.D, p.M, p.N, p.K, p.ldd, p.relu, sp));
        count_launch();
        return OCR_OK;
    }
    gemm_tf32_kernel<BN, STAGES, H16><<<grid, kGemmThreads, S::kTotal, st>>>(p.tmA, p.tmB, p.tmD, p.bias, p.D, p.M, p.N, p.K, p.ldd, p.relu, sp);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

namespace ocr {

int gemm_plan(GemmPlan* p, const float* A, int lda, const float* W, int ldw, const float* bias, float* D, int ldd, int M, int N,
              int K, int relu)
{
    OCR_CHECK_ARG(M >= 1 && N >= 1 && K >= 1, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
    OCR_CHECK_ARG(A && W && D, "gemm: NULL argument");
    OCR_CHECK_ARG(lda >= K && ldw >= K && ldd >= N, "gemm: leading dimensions too small");
    OCR_CHECK_ARG((lda % 4) == 0 && (ldw % 4) == 0 && ((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0,
                  "gemm: A and W need 16-byte aligned rows (pointer and leading dimension * 4 bytes)");
    const long long mt = (M + kGemmBM - 1) / kGemmBM;
    // tile width: the main loop is bound by the bytes an SM pulls in per k-step (16 KB of A + BN * 128 B of W), so minimise
    // (waves of CTAs over the 148 SMs) x (bytes per k-step); see conv_igemm.cu
    int bn = 32;
    long long best = -1;
    for (int cand = 32; cand <= 256; cand *= 2) {
        if (cand > 32 && cand / 2 >= N) break;
        const long long tiles = mt * ((N + cand - 1) / cand);
        const long long cost = ((tiles + 147) / 148) * (128 + cand);
        if (best < 0 || cost < best) { best = cost; bn = cand; }
    }
    p->bn = bn; p->bias = bias; p->D = D; p->M = M; p->N = N; p->K = K; p->ldd = ldd; p->relu = relu;
    int rc = tma_map_2d(&p->tmA, A, M, K, lda, kGemmBM);
    if (rc != OCR_OK) return rc;
    // TMA-store epilogue: output rows must be 16-byte aligned (tensor-map stride rule) and N a multiple of 4 (a bulk tensor
    // store writes whole 16-byte units: with N = 50 it zeroed columns 50 and 51 of a wider row; measured)
    p->tma_store = 0;
    if ((ldd % 4) == 0 && (N % 4) == 0 && ((uintptr_t)D % 16) == 0) {
        rc = tma_map_out(&p->tmD, D, M, N, ldd);
        if (rc != OCR_OK) return rc;
        p->tma_store = 1;
    }
    return tma_map_2d(&p->tmB, W, N, K, ldw, bn);
}

// gemm_plan with binary16 operands A [M, K], W [N, K] (row pitches in elements, multiples of 8): tcgen05.mma.kind::f16, float32
// sums, bias, output and epilogue as gemm_plan.  binary16 carries the 10 mantissa bits TF32 keeps, so for operands inside its
// range (|v| < 65504, steps of 6e-8 near zero) the products are the TF32 products at twice the tensor rate and half the bytes.
int gemm_plan_f16(GemmPlan* p, const void* A, int lda, const void* W, int ldw, const float* bias, float* D, int ldd, int M, int N, int K, int relu)
{
    OCR_CHECK_ARG(M >= 1 && N >= 1 && K >= 1, "gemm_f16: bad shape M=%d N=%d K=%d", M, N, K);
    OCR_CHECK_ARG(A && W && D, "gemm_f16: NULL argument");
    OCR_CHECK_ARG(lda >= K && ldw >= K && ldd >= N, "gemm_f16: leading dimensions too small");
    OCR_CHECK_ARG((lda % 8) == 0 && (ldw % 8) == 0 && ((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0,
                  "gemm_f16: A and W need 16-byte aligned rows (pointer and leading dimension * 2 bytes)");
    const long long mt = (M + kGemmBM - 1) / kGemmBM;
    int bn = 32;
    long long best = -1;
    for (int cand = 32; cand <= 256; cand *= 2) {
        if (cand > 32 && cand / 2 >= N) break;
        const long long tiles = mt * ((N + cand - 1) / cand);
        const long long cost = ((tiles + 147) / 148) * (128 + cand);
        if (best < 0 || cost < best) { best = cost; bn = cand; }
    }
    p->bn = bn; p->bias = bias; p->D = D; p->M = M; p->N = N; p->K = K; p->ldd = ldd; p->relu = relu; p->h16 = 2;
    int rc = tma_map_2d_h(&p->tmA, A, M, K, lda, kGemmBM);
    if (rc != OCR_OK) return rc;
    p->tma_store = 0;
    if ((ldd % 4) == 0 && (N % 4) == 0 && ((uintptr_t)D % 16) == 0) {
        rc = tma_map_out(&p->tmD, D, M, N, ldd);
        if (rc != OCR_OK) return rc;
        p->tma_store = 1;
    }
    return tma_map_2d_h(&p->tmB, W, N, K, ldw, bn);
}

// The two directions of a recurrent layer as ONE launch: batch d multiplies rows [d*M, (d+1)*M) of A with rows
// [d*N, (d+1)*N) of W; outputs are dense [M, N] tiles batch_stride apart; optional split over K (partials split_stride apart).
int gemm_plan_dirs(GemmPlan* p, const float* A, int lda, const float* W, int ldw, float* D, int M, int N, int K, int ndir, int splits, int bn)
{
    OCR_CHECK_ARG(M >= 1 && N >= 1 && K >= 1 && ndir >= 1 && ndir <= 9 && splits >= 1, "gemm_plan_dirs: bad shape");
    OCR_CHECK_ARG((lda % 4) == 0 && (ldw % 4) == 0 && ((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0, "gemm_plan_dirs: operands need 16-byte aligned rows");
    const int nk = (K + kGemmBK - 1) / kGemmBK;
    if (splits > nk) splits = nk;
    const int kps = (nk + splits - 1) / splits;
    splits = (nk + kps - 1) / kps;
    p->bn = bn; p->bias = nullptr; p->D = D; p->M = M; p->N = N; p->K = K; p->ldd = N; p->relu = 0;
    p->splits = splits; p->ksteps_per_split = splits > 1 ? kps : 0; p->nbatch = ndir;
    for (int d = 0; d < ndir; ++d) { p->a_shift[d] = 0; p->a_row[d] = d * M; p->b_row[d] = d * N; }
    p->split_stride = (long long)M * N;
    p->batch_stride = (long long)splits * M * N;
    int rc = tma_map_2d(&p->tmA, A, (long long)ndir * M, K, lda, kGemmBM);
    if (rc != OCR_OK) return rc;
    return tma_map_2d(&p->tmB, W, (long long)ndir * N, K, ldw, bn);
}

int gemm_plan_dirs_h16(GemmPlan* p, const void* A, int lda, const void* W, int ldw, float* D, int M, int N, int K, int ndir, int splits, int bn)
{
    OCR_CHECK_ARG(M >= 1 && N >= 1 && K >= 1 && ndir >= 1 && ndir <= 9 && splits >= 1 && (bn == 32 || bn == 64), "gemm_plan_dirs_h16: bad shape");
    OCR_CHECK_ARG((lda % 8) == 0 && (ldw % 8) == 0 && ((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0, "gemm_plan_dirs_h16: operands need 16-byte aligned rows");
    const int nk = (K + kGemmBKh - 1) / kGemmBKh;
    if (splits > nk) splits = nk;
    const int kps = (nk + splits - 1) / splits;
    splits = (nk + kps - 1) / kps;
    p->bn = bn; p->bias = nullptr; p->D = D; p->M = M; p->N = N; p->K = K; p->ldd = N; p->relu = 0; p->h16 = 1;
    p->splits = splits; p->ksteps_per_split = splits > 1 ? kps : 0; p->nbatch = ndir;
    for (int d = 0; d < ndir; ++d) { p->a_shift[d] = 0; p->a_row[d] = d * M; p->b_row[d] = d * N; }
    p->split_stride = (long long)M * N;
    p->batch_stride = (long long)splits * M * N;
    int rc = tma_map_2d_h(&p->tmA, A, (long long)ndir * M, K, lda, kGemmBM);
    if (rc != OCR_OK) return rc;
    return tma_map_2d_h(&p->tmB, W, (long long)ndir * N, K, ldw, bn);
}

int gemm_run(const GemmPlan& p, cudaStream_t st)
{
    if (p.h16 == 1) return p.bn == 32 ? launch_planned<32, 8, 1>(p, st) : launch_planned<64, 6, 1>(p, st);
    if (p.h16 == 2) {             // binary16 operands (gemm_plan_f16): the many-tile projections, shallow rings as below
        switch (p.bn) {
            case 32: return launch_planned<32, 3, 2>(p, st);
            case 64: return launch_planned<64, 3, 2>(p, st);
            case 128: return launch_planned<128, 3, 2>(p, st);
            default: return launch_planned<256, 2, 2>(p, st);
        }
    }
    // Two shapes of pipeline.  Grids that put at most one CTA on an SM (split-K contractions, the small per-frame
    // products) get a deep TMA ring: the k loop is all there is.  Grids of many short tiles get a shallow ring so that
    // 2-3 CTAs share an SM (shared memory and the 512 TMEM columns allow it) and the epilogue of one tile -- one thread per
    // row draining TMEM to global memory, which is what bounds the K <= 1024 projections -- overlaps the main loop of another.
    const long long ctas = (long long)((p.M + kGemmBM - 1) / kGemmBM) * p.nbatch * ((p.N + p.bn - 1) / p.bn) * p.splits;
    if (ctas <= 148 && !p.shallow) {
        switch (p.bn) {
            case 32: return launch_planned<32, 8>(p, st);
            case 64: return launch_planned<64, 6>(p, st);
            case 128: return launch_planned<128, 5>(p, st);
            default: return launch_planned<256, 4>(p, st);
        }
    }
    switch (p.bn) {
        case 32: return launch_planned<32, 3>(p, st);     // 60 KB  -> 3 CTAs / SM
        case 64: return launch_planned<64, 3>(p, st);     // 72 KB  -> 3
        case 128: return launch_planned<128, 3>(p, st);   // 96 KB  -> 2
        default: return launch_planned<256, 2>(p, st);    // 96 KB  -> 2 (2 x 256 TMEM columns)
    }
}

// D[b][i] = sum_z partials[b][z][i] in a fixed order (deterministic split-K)
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partials, int splits, int M, int N, int ldd, long long batch_stride, int nbatch,
                     float* __restrict__ D, long long in_batch_stride, long long in_split_stride)
{
    const long long per = (long long)M * N, total = per * nbatch;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(idx / per);
        const long long i = idx - (long long)b * per;
        const float* src = partials + (size_t)b * in_batch_stride + i;
        float acc = 0.0f;
        for (int z = 0; z < splits; ++z) acc += src[(size_t)z * in_split_stride];
        D[(size_t)b * batch_stride + (size_t)(i / N) * ldd + (i % N)] = acc;
    }
}

static int g_wgrad_stack = 1;   // stack the tap views of shallow layers in one tile (0: one tile per tap, the round-1 form)
// taps per tile: views of at most 64 rows (a multiple of 8) are stacked in the 128 rows of an A tile
static int wgrad_tp(int M, int nbatch) {
    if (!g_wgrad_stack || nbatch < 2 || M > 64 || (M % 8) != 0) return 1;
    return kGemmBM / M;
}
static int wgrad_splits(int M, int N, long long R, int nbatch, int* ksteps_per_split) {
    const long long nk = (R + kGemmBK - 1) / kGemmBK;
    const int bn = N > 128 ? 256 : (N > 64 ? 128 : (N > 32 ? 64 : 32));
    const int tp = wgrad_tp(M, nbatch);
    const long long tiles = tp > 1 ? (long long)((nbatch + tp - 1) / tp) * ((N + bn - 1) / bn)
                                   : (long long)((M + kGemmBM - 1) / kGemmBM) * ((N + bn - 1) / bn) * nbatch;
    long long want = 148 / tiles;                               // one full wave of CTAs (one CTA per SM): no ragged last wave
    if (want > nk / 8) want = nk / 8;                           // at least 8 k-steps per split
    if (want < 1) want = 1;
    const long long kps = (nk + want - 1) / want;
    *ksteps_per_split = (int)kps;
    return (int)((nk + kps - 1) / kps);
}

size_t gemm_wgrad_scratch_floats(int M, int N, long long R, int nbatch) {
    int kps;
    const int splits = wgrad_splits(M, N, R, nbatch, &kps);
    return (size_t)splits * nbatch * M * N;
}

int gemm_wgrad(const float* A, long long lda, const float* W, long long ldw, float* D, int ldd, long long batch_stride, int M, int N,
               long long R, int nbatch, const int* a_shift, const int* a_row, long long a_rows, float* partials, cudaStream_t st,
               int blocked, long long w_rows)
{
    OCR_CHECK_ARG(M >= 1 && N >= 1 && R >= 1 && nbatch >= 1 && nbatch <= 9, "gemm_wgrad: bad shape M=%d N=%d R=%lld nbatch=%d", M, N, R, nbatch);
    OCR_CHECK_ARG(A && W && D && partials, "gemm_wgrad: NULL argument");
    if (blocked) {
        OCR_CHECK_ARG((R % kGemmBK) == 0 && w_rows >= N && ((uintptr_t)A % 128) == 0 && ((uintptr_t)W % 128) == 0 && R < 0x7fffffffLL,
                      "gemm_wgrad: blocked operands need R %% 32 == 0 and 128-byte aligned bases");
    } else {
        OCR_CHECK_ARG((lda % 4) == 0 && (ldw % 4) == 0 && ((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0 && R < 0x7fffffffLL,
                      "gemm_wgrad: operands need 16-byte aligned rows");
    }
    GemmPlan p;
    p.blocked = blocked ? 1 : 0;
    p.bn = N > 128 ? 256 : (N > 64 ? 128 : (N > 32 ? 64 : 32));
    p.bias = nullptr; p.M = M; p.N = N; p.K = (int)R; p.relu = 0;
    p.splits = wgrad_splits(M, N, R, nbatch, &p.ksteps_per_split);
    p.nbatch = nbatch;
    for (int i = 0; i < nbatch; ++i) { p.a_shift[i] = a_shift ? a_shift[i] : 0; p.a_row[i] = a_row ? a_row[i] : 0; }
    if (blocked)
        for (int i = 0; i < nbatch; ++i) {
            OCR_CHECK_ARG((p.a_shift[i] % kGemmBK) == 0, "gemm_wgrad: blocked operands need shifts that are multiples of 32 (a_shift[%d] = %d)", i, p.a_shift[i]);
            p.a_shift[i] /= kGemmBK;
        }
    // partial tiles are dense M x N
    p.D = partials; p.ldd = N;
    p.split_stride = (long long)M * N;
    p.batch_stride = (long long)p.splits * M * N;
    long long in_batch_stride = p.batch_stride, in_split_stride = p.split_stride;
    p.a_box_rows = M >= kGemmBM ? kGemmBM : (M + 7) / 8 * 8;   // rows past M belong to other views: do not fetch them
    const int tp = wgrad_tp(M, nbatch);
    if (tp > 1) {
        // tile = tp stacked views; partials laid out [split][view][M][N] so that a tile's rows are contiguous
        p.tp = tp; p.ntaps = nbatch;
        p.nbatch = (nbatch + tp - 1) / tp;
        p.batch_stride = (long long)tp * M * N;
        p.split_stride = (long long)nbatch * M * N;
        in_batch_stride = (long long)M * N;
        in_split_stride = p.split_stride;
    }
    int rc = blocked ? tma_map_blocked(&p.tmA, A, a_rows, R / kGemmBK, p.a_box_rows) : tma_map_2d(&p.tmA, A, a_rows, R, lda, p.a_box_rows);
    if (rc != OCR_OK) return rc;
    rc = blocked ? tma_map_blocked(&p.tmB, W, w_rows, R / kGemmBK, p.bn) : tma_map_2d(&p.tmB, W, N, R, ldw, p.bn);
    if (rc != OCR_OK) return rc;
    if (p.splits == 1) p.ksteps_per_split = 0;
    rc = gemm_run(p, st);
    if (rc != OCR_OK) return rc;
    const long long total = (long long)M * N * nbatch;
    long long g = (total + 255) / 256;
    splitk_reduce_kernel<<<(int)(g > 148 * 8 ? 148 * 8 : g), 256, 0, st>>>(partials, p.splits, M, N, ldd, batch_stride, nbatch, D, in_batch_stride, in_split_stride);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

}  // namespace ocr

// Tuning aid: bit 0 = TMA-store epilogue of ocr_gemm_tf32 (default on), bit 1 = one tile per tap view in ocr_gemm_tf32_wgrad
// (default off: the views of layers with at most 64 input channels are stacked in one tile).
extern "C" int ocr_debug_gemm_tma_store(int on) {
    g_wgrad_stack = (on & 2) ? 0 : 1;
    g_tma_promo256 = (on & 4) ? 1 : 0;
    return gemm_set_tma_store(on & 1);
}

__global__ void __launch_bounds__(256) float_to_half_kernel(const float4* __restrict__ in, uint2* __restrict__ out, long long n4)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = in[i];
        const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
        out[i] = make_uint2(*reinterpret_cast<const unsigned*>(&lo), *reinterpret_cast<const unsigned*>(&hi));
    }
}

extern "C" int ocr_float_to_half(const float* in, void* out, long long n, ocr_stream_t stream)
{
    OCR_CHECK_ARG(n >= 0 && (n % 4) == 0, "ocr_float_to_half: n must be a multiple of 4");
    if (n == 0) return OCR_OK;
    OCR_CHECK_ARG(in && out && ((uintptr_t)in % 16) == 0 && ((uintptr_t)out % 8) == 0, "ocr_float_to_half: NULL or misaligned argument");
    long long g = (n / 4 + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    float_to_half_kernel<<<(unsigned)g, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<uint2*>(out), n / 4);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

extern "C" int ocr_gemm_f16(const void* A, int lda, const void* W, int ldw, const float* bias, float* D, int ldd, int M, int N, int K, int relu,
                            ocr_stream_t stream)
{
    OCR_CHECK_ARG(M >= 0 && N >= 0 && K >= 1, "ocr_gemm_f16: bad shape M=%d N=%d K=%d", M, N, K);
    if (M == 0 || N == 0) return OCR_OK;
    GemmPlan p;
    int rc = gemm_plan_f16(&p, A, lda, W, ldw, bias, D, ldd, M, N, K, relu);
    if (rc != OCR_OK) return rc;
    return gemm_run(p, static_cast<cudaStream_t>(stream));
}

extern "C" int ocr_gemm_tf32(const float* A, int lda, const float* W, int ldw, const float* bias, float* D, int ldd, int M, int N,
                             int K, int relu, ocr_stream_t stream)
{
    OCR_CHECK_ARG(M >= 0 && N >= 0 && K >= 1, "ocr_gemm_tf32: bad shape M=%d N=%d K=%d", M, N, K);
    if (M == 0 || N == 0) return OCR_OK;
    GemmPlan p;
    int rc = gemm_plan(&p, A, lda, W, ldw, bias, D, ldd, M, N, K, relu);
    if (rc != OCR_OK) return rc;
    return gemm_run(p, static_cast<cudaStream_t>(stream));
}
