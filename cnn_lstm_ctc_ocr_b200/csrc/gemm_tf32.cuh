// Internal interface of the tcgen05 TF32 GEMM (gemm_tf32.cu): plan once (two TMA tensor maps), run many times.
// The recurrent layers re-launch the same h * W_h^T product every frame with unchanged buffers.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace ocr {

struct GemmPlan {
    CUtensorMap tmA, tmB;
    const float* bias;
    float* D;
    int M, N, K, ldd, relu, bn;
};

// D[M,N] = act(A[M,K] * W[N,K]^T + bias[N]); see ocr_gemm_tf32 in include/ocr_b200.h for the operand rules
int gemm_plan(GemmPlan* p, const float* A, int lda, const float* W, int ldw, const float* bias, float* D, int ldd, int M, int N,
              int K, int relu);
int gemm_run(const GemmPlan& p, cudaStream_t st);

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 32;  // fp32 elements = one 128-byte swizzle row

// 2-D fp32 tensor [rows, K] row-major with row pitch ld elements; box = [box_rows, 32 floats], 128-byte swizzle
int tma_map_2d(CUtensorMap* tm, const float* base, long long rows, long long K, long long ld, int box_rows);

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
__device__ __forceinline__ void umma_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

#endif

}  // namespace ocr
