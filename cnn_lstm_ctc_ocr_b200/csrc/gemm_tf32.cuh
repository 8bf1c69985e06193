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

}  // namespace ocr
