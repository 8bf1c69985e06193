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
#include "gemm_tf32.cuh"

namespace ocr {

constexpr int kGemmThreads = 192;

template <int BN, int STAGES>
struct GemmSmem {
    static constexpr int kA = kGemmBM * kGemmBK * 4;
    static constexpr int kB = BN * kGemmBK * 4;
    static constexpr int kStage = kA + kB;
    static constexpr int kBars = STAGES * kStage;
    static constexpr int kTotal = kBars + (2 * STAGES + 1) * 8 + 16 + 1024;  // + alignment slack
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kGemmThreads)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const float* __restrict__ bias, float* __restrict__ D, int M, int N, int K, int ldd, int relu)
{
    using S = GemmSmem<BN, STAGES>;
    extern __shared__ unsigned char gemm_smem_raw[];
    // 128-byte swizzle wants 1024-byte aligned tiles
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gemm_smem_raw) + 1023) & ~(uintptr_t)1023);
    const unsigned s_base = g_smem_u32(smem);
    const unsigned bar_full = s_base + S::kBars;
    const unsigned bar_empty = bar_full + STAGES * 8;
    const unsigned bar_acc = bar_empty + STAGES * 8;
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(smem + S::kBars + (2 * STAGES + 1) * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kGemmBM, n0 = blockIdx.y * BN;
    const int nk = (K + kGemmBK - 1) / kGemmBK;

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

    if (warp == 0) {
        if (lane == 0) {
            for (int k = 0; k < nk; ++k) {
                const int s = k % STAGES;
                if (k >= STAGES) g_mbar_wait(bar_empty + s * 8, ((k / STAGES) - 1) & 1);
                g_mbar_expect_tx(bar_full + s * 8, (unsigned)S::kStage);
                tma_load_2d(s_base + s * S::kStage, &tmA, k * kGemmBK, m0, bar_full + s * 8);
                tma_load_2d(s_base + s * S::kStage + S::kA, &tmB, k * kGemmBK, n0, bar_full + s * 8);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: fp32 accumulate, tf32 x tf32, both K-major, N = BN, M = 128
            const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(BN >> 3) << 17) | ((unsigned)(kGemmBM >> 4) << 24);
            for (int k = 0; k < nk; ++k) {
                const int s = k % STAGES;
                g_mbar_wait(bar_full + s * 8, (k / STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned a_addr = s_base + s * S::kStage, b_addr = a_addr + S::kA;
                const unsigned long long da = umma_desc_k128(a_addr), db = umma_desc_k128(b_addr);
#pragma unroll
                for (int kk = 0; kk < kGemmBK / 8; ++kk)  // 8 tf32 = 32 bytes per MMA: advance the start address inside the swizzle row
                    umma_tf32(tmem_d, da + (unsigned long long)(kk * 2), db + (unsigned long long)(kk * 2), idesc, (k | kk) ? 1u : 0u);
                umma_commit(bar_empty + s * 8);  // frees the stage when the MMAs that read it retire
            }
            umma_commit(bar_acc);
        }
    } else {
        // epilogue warps 2..5: TMEM lane quarter = warp % 4
        const int q = warp & 3;
        g_mbar_wait(bar_acc, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        gemm_epilogue<BN>(tmem_d, q, lane, m0, n0, M, N, bias, D, ldd, relu);
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
int tma_map_2d(CUtensorMap* tm, const float* base, long long rows, long long K, long long ld, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return OCR_ECUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)kGemmBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld K=%lld ld=%lld", (int)r, rows, K, ld); return OCR_ECUDA; }
    return OCR_OK;
}
}  // namespace ocr

template <int BN, int STAGES>
static int launch_planned(const GemmPlan& p, cudaStream_t st)
{
    using S = GemmSmem<BN, STAGES>;
    static int configured = -1;
    int dev = 0;
    OCR_CHECK_CUDA(cudaGetDevice(&dev));
    if (configured != dev) {
        OCR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
        configured = dev;
    }
    dim3 grid((p.M + kGemmBM - 1) / kGemmBM, (p.N + BN - 1) / BN);
    gemm_tf32_kernel<BN, STAGES><<<grid, kGemmThreads, S::kTotal, st>>>(p.tmA, p.tmB, p.bias, p.D, p.M, p.N, p.K, p.ldd, p.relu);
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
    // widest tile that still gives the 148 SMs something to do
    int bn = 32;
    if (N > 32) bn = 64;
    if (N > 64 && mt * ((N + 127) / 128) >= 120) bn = 128;
    if (N > 128 && mt * ((N + 255) / 256) >= 120) bn = 256;
    p->bn = bn; p->bias = bias; p->D = D; p->M = M; p->N = N; p->K = K; p->ldd = ldd; p->relu = relu;
    int rc = tma_map_2d(&p->tmA, A, M, K, lda, kGemmBM);
    if (rc != OCR_OK) return rc;
    return tma_map_2d(&p->tmB, W, N, K, ldw, bn);
}

int gemm_run(const GemmPlan& p, cudaStream_t st)
{
    switch (p.bn) {
        case 32: return launch_planned<32, 8>(p, st);
        case 64: return launch_planned<64, 6>(p, st);
        case 128: return launch_planned<128, 5>(p, st);
        default: return launch_planned<256, 4>(p, st);
    }
}

}  // namespace ocr

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
