// Probe: where do the rows of an M=64 tcgen05.mma (cta_group::1, kind::tf32) accumulator land in TMEM?
#include <cstdio>
#include <cuda_runtime.h>
#include "../cnn_lstm_ctc_ocr_b200/csrc/gemm_tf32.cuh"
using namespace ocr;
__global__ void probe(float* out /*[128][32]*/, int M) {
    extern __shared__ unsigned char raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    const unsigned s_base = g_smem_u32(smem);
    float* A = reinterpret_cast<float*>(smem);            // [128 rows][32 floats] swizzled K-major tile
    float* Bt = reinterpret_cast<float*>(smem + 16384);   // [32 rows][32 floats]
    __shared__ unsigned tmem_slot;
    __shared__ __align__(8) unsigned long long bar;
    for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) A[i] = 0.f;
    for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) Bt[i] = 0.f;
    __syncthreads();
    // element (row r, k): 16-byte group (k/4) ^ (r & 7)
    if (threadIdx.x < 128) { const int r = threadIdx.x; A[r * 32 + ((0 ^ (r & 7)) << 2) + 0] = (float)(r + 1); }
    if (threadIdx.x < 32) { const int n = threadIdx.x; Bt[n * 32 + ((0 ^ (n & 7)) << 2) + 0] = (float)(1 + n * 0); }
    const unsigned b = g_smem_u32(&bar);
    if (threadIdx.x == 0) { g_mbar_init(b, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(g_smem_u32(&tmem_slot)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_d = tmem_slot;
    // pre-fill TMEM with a marker through an M=128 MMA of zeros? simply read garbage: mark by first issuing M=128 with A=0 -> zeros
    if (threadIdx.x == 0) {
        const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(32 >> 3) << 17) | ((unsigned)(M >> 4) << 24);
        umma_tf32(tmem_d, umma_desc_k128(s_base), umma_desc_k128(s_base + 16384), idesc, 0u);
        umma_commit(b);
    }
    g_mbar_wait(b, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    unsigned r[32];
    const unsigned taddr = tmem_d + ((unsigned)(warp * 32) << 16);
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
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 32 + j] = __uint_as_float(r[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(32u) : "memory");
}
int main() {
    float* d; cudaMalloc(&d, 128 * 32 * 4);
    float h[128 * 32];
    for (int M : {128, 64}) {
        cudaMemset(d, 0xff, 128 * 32 * 4);
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        probe<<<1, 128, 48 * 1024>>>(d, M);
        cudaError_t e = cudaDeviceSynchronize();
        printf("M=%d: %s\n", M, cudaGetErrorString(e));
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        for (int l = 0; l < 128; ++l) { printf("lane %3d: c0=%6.1f c1=%6.1f c16=%6.1f c31=%6.1f\n", l, h[l * 32], h[l * 32 + 1], h[l * 32 + 16], h[l * 32 + 31]); }
    }
    return 0;
}
