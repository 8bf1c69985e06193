// 3x3 'same' convolution for the wide, shallow layers (conv2..conv4 of /root/reference/src/weinman/model.py:47-54,84-109 and
// their input gradients): implicit GEMM on tcgen05 with NO gather at all.
//
// conv_igemm.cu builds every tap's patch tile with per-pixel cp.async copies: each input element crosses L2 -> SM nine
// times, and for C = 32 / 64 channels that traffic, not the tensor pipe, is the limit (ncu: L2 66 %, tensor pipe 5 % on
// conv2).  Here a CTA owns a 16 x 8 patch of output pixels and TMA loads the 18 x 16-pixel input halo of each 32-channel
// chunk ONCE (4-D tensor map over NHWC, out-of-image pixels zero-filled = the 'same' padding) as a 128B-swizzled tile with
// one 128-byte row per pixel.  Row group ty of the A operand of tap (dy, dx) -- the 8 pixels (y0+ty+dy-1, x0+dx-1 .. +7) --
// is 8 consecutive rows of that tile, and consecutive groups are one halo row (16 pixels = 2048 bytes) apart: every tap's
// patch matrix IS a view of the halo tile, described to the MMA by a shared-memory descriptor with stride-byte-offset
// 2048 whose start address sits dx rows into a 1024-byte swizzle atom.  (Measured: the MMA un-swizzles by the ABSOLUTE
// shared-memory address bits [7:9], exactly as TMA swizzled on the way in, so the view needs matrix-base-offset 0; a
// base offset of dx is subtracted from those bits and scrambles the 16-byte channel groups.)  Nine taps x C/32 chunks x 4
// MMAs read the same staged bytes; nothing is copied, no thread touches the input.
//   warp 0 / one lane : TMA: halo tiles (once), then the filter tiles of each k-step (tap-major, chunk inner: the same
//                       k order as conv_igemm.cu, so both kernels produce bit-identical sums) through a small ring
//   warp 1 / one lane : tcgen05.mma.kind::tf32, accumulator in TMEM
//   warps 2..5        : epilogue, TMEM lane m = ty*8 + tx -> out[b, y0+ty, x0+tx, :] (+ bias, ReLU)
//
// Round 2: the epilogue stages each warp's 4 x 8 pixels x 32 channels in shared memory (the halo tiles are idle by then) and
// stores them with ONE cp.async.bulk.tensor (4-D map over the NHWC output, box {32 ch, 8 x, 4 y, 1}: clipping at the image
// border comes with the map), and it can apply the max-pool that follows the layer (model.py:111-116, pool_layer) on the way:
//   POOL 1 = window 2x2, stride (2,2) (pool2): a warp's 4 x 8 pixels hold eight whole windows (lanes l, l^1, l^8, l^9): two
//            shuffles per value, the warp stores a 2 x 4 block of the pooled tensor;
//   POOL 2 = window 2x2, stride (2,1) (pool4): horizontal stride 1 needs the right-hand neighbour, so the patches overlap by
//            one column (x0 = 7 * tile): 7 pooled columns per patch, a 2 x 7 block per warp.
// conv -> (folded batch-norm) -> ReLU -> pool then is one launch and the unpooled activation never reaches memory.
#include "gemm_tf32.cuh"

namespace ocr {

constexpr int kHaloTY = 16, kHaloTX = 8;                       // output patch
constexpr int kHaloRows = kHaloTY + 2, kHaloCols = 16;         // staged pixels: 18 rows x 16 columns (10 used)
constexpr int kHaloBytes = kHaloRows * kHaloCols * 128;        // per 32-channel chunk: 36 KB
constexpr int kHaloThreads = 192;

__device__ __forceinline__ void tma_load_4d(unsigned dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// K-major, 128-byte swizzle, 8-row groups `sbo` bytes apart; base_off = matrix-base-offset field (0 for TMA-written tiles)
__device__ __forceinline__ unsigned long long umma_desc_view(unsigned smem_addr, unsigned sbo, unsigned base_off) {
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr >> 4) & 0x3FFF);
    d |= (unsigned long long)1 << 16;
    d |= (unsigned long long)(sbo >> 4) << 32;
    d |= (unsigned long long)1 << 46;
    d |= (unsigned long long)(base_off & 7) << 49;
    d |= (unsigned long long)2 << 61;
    return d;
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, int c0, int c1, int c2, int c3, unsigned src) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

template <int BN, int WSTAGES, int POOL>   // POOL: 0 = none, 1 = max 2x2 stride (2,2), 2 = max 2x2 stride (2,1) fused into the epilogue
__global__ void __launch_bounds__(kHaloThreads)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmOut,
                    int B, int H, int W, int C, const float* __restrict__ bias, float* __restrict__ out, int Cout, int relu, int tiles_x,
                    int tiles_y, int tma_store)
{
    constexpr int kWB = BN * kGemmBK * 4;                       // one filter tile
    extern __shared__ unsigned char halo_smem_raw[];
    unsigned char* smem = halo_smem_raw + ((1024u - (g_smem_u32(halo_smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    const unsigned s_base = g_smem_u32(smem);
    const int cpt = C / kGemmBK;                                // channel chunks
    const unsigned s_halo = s_base;                             // cpt halo tiles
    const unsigned s_w = s_halo + (unsigned)cpt * kHaloBytes;   // filter ring
    const unsigned s_bar = s_w + WSTAGES * kWB;
    const unsigned bar_full = s_bar, bar_empty = s_bar + WSTAGES * 8, bar_halo = bar_empty + WSTAGES * 8, bar_acc = bar_halo + 8;
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(smem + (s_bar - s_base) + (2 * WSTAGES + 2) * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int tile = blockIdx.x;
    const int tx_i = tile % tiles_x; tile /= tiles_x;
    const int ty_i = tile % tiles_y;
    const int b = tile / tiles_y;
    const int x0 = tx_i * (POOL == 2 ? kHaloTX - 1 : kHaloTX), y0 = ty_i * kHaloTY;   // POOL 2: patches overlap by one column
    const int n0 = blockIdx.y * BN;
    const int nk = 9 * cpt;

    if (threadIdx.x == 0) {
        for (int s = 0; s < WSTAGES; ++s) { g_mbar_init(bar_full + s * 8, 1); g_mbar_init(bar_empty + s * 8, 1); }
        g_mbar_init(bar_halo, 1);
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
            // halo of every channel chunk: box {32 ch, 16 x, 18 y, 1 image} at (c, x0-1, y0-1, b); outside the image -> 0
            g_mbar_expect_tx(bar_halo, (unsigned)cpt * kHaloBytes);
            for (int c = 0; c < cpt; ++c) tma_load_4d(s_halo + c * kHaloBytes, &tmIn, c * kGemmBK, x0 - 1, y0 - 1, b, bar_halo);
            for (int k = 0; k < nk; ++k) {
                const int s = k % WSTAGES;
                if (k >= WSTAGES) g_mbar_wait(bar_empty + s * 8, ((k / WSTAGES) - 1) & 1);
                g_mbar_expect_tx(bar_full + s * 8, (unsigned)kWB);
                tma_load_2d(s_w + s * kWB, &tmW, k * kGemmBK, n0, bar_full + s * 8);     // filter columns (tap, chunk) = k-step k
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(BN >> 3) << 17) | ((unsigned)(kGemmBM >> 4) << 24);
            g_mbar_wait(bar_halo, 0);
            for (int k = 0; k < nk; ++k) {
                const int s = k % WSTAGES;
                g_mbar_wait(bar_full + s * 8, (k / WSTAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int tap = k / cpt, c = k - tap * cpt;
                const int dy = tap / 3, dx = tap % 3;                       // halo row / column of the patch's first pixel
                const unsigned a_addr = s_halo + c * kHaloBytes + (unsigned)(dy * kHaloCols + dx) * 128u;
                const unsigned long long da = umma_desc_view(a_addr, kHaloCols * 128u, 0u);
                const unsigned long long db = umma_desc_k128(s_w + s * kWB);
#pragma unroll
                for (int kk = 0; kk < kGemmBK / 8; ++kk)
                    umma_tf32(tmem_d, da + (unsigned long long)(kk * 2), db + (unsigned long long)(kk * 2), idesc, (k | kk) ? 1u : 0u);
                umma_commit(bar_empty + s * 8);
            }
            umma_commit(bar_acc);
        }
    } else {
        const int q = warp & 3;
        const int m = q * 32 + lane;                    // TMEM lane = patch pixel ty*8 + tx
        const int y = y0 + (m >> 3), x = x0 + (m & 7);
        const bool live = y < H && x < W;
        g_mbar_wait(bar_acc, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float* drow = out + (((size_t)b * H + y) * W + x) * Cout + n0;
        // staging: two 4 KB blocks per warp in the (idle) halo tiles; rows of 128 bytes, 16-byte groups XOR (row & 7)
        const unsigned my = s_halo + (unsigned)q * 8192u;
        const int tyl = lane >> 3, tx = lane & 7;
        // the staged row this lane writes (-1: none) and the block's origin in the output tensor
        int srow, ox, oy;
        if (POOL == 0) { srow = lane; ox = x0; oy = y0 + 4 * q; }
        else if (POOL == 1) { srow = ((tyl | tx) & 1) ? -1 : (tyl >> 1) * 4 + (tx >> 1); ox = x0 >> 1; oy = (y0 + 4 * q) >> 1; }
        else { srow = ((tyl & 1) || tx == 7) ? -1 : (tyl >> 1) * 7 + tx; ox = x0; oy = (y0 + 4 * q) >> 1; }
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            if ((POOL != 0 || tma_store) && n0 + c0 >= Cout) break;
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
            if (POOL == 0 && !tma_store) {
                if (live) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        if (n0 + c0 + j < Cout) {            // Cout % 4 == 0
                            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + n0 + c0 + j));
                            float4 v = make_float4(__uint_as_float(r[j]) + bb.x, __uint_as_float(r[j + 1]) + bb.y,
                                                   __uint_as_float(r[j + 2]) + bb.z, __uint_as_float(r[j + 3]) + bb.w);
                            if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                            *reinterpret_cast<float4*>(drow + c0 + j) = v;
                        }
                    }
                }
                continue;
            }
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float xv = __uint_as_float(r[j]);
                if (n0 + c0 + j < Cout) xv += __ldg(bias + n0 + c0 + j);
                v[j] = relu ? fmaxf(xv, 0.0f) : xv;
            }
            if (POOL == 1) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    v[j] = fmaxf(v[j], __shfl_xor_sync(kFullMask, v[j], 1));   // (x, x+1)
                    v[j] = fmaxf(v[j], __shfl_xor_sync(kFullMask, v[j], 8));   // (y, y+1)
                }
            } else if (POOL == 2) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    v[j] = fmaxf(v[j], __shfl_down_sync(kFullMask, v[j], 1));  // (x, x+1); lanes with tx == 7 are not stored
                    v[j] = fmaxf(v[j], __shfl_xor_sync(kFullMask, v[j], 8));   // (y, y+1)
                }
            }
            const unsigned blk = my + (unsigned)((c0 >> 5) & 1) * 4096u;
            if (c0 >= 64) {   // the block is being reused: the store issued two steps ago must have read it
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
            }
            if (srow >= 0) {
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(blk + (unsigned)srow * 128u + (unsigned)((j4 ^ (srow & 7)) << 4)),
                                 "f"(v[4 * j4]), "f"(v[4 * j4 + 1]), "f"(v[4 * j4 + 2]), "f"(v[4 * j4 + 3]) : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async-proxy (TMA) reads
            __syncwarp();
            if (lane == 0) {
                tma_store_4d(&tmOut, n0 + c0, ox, oy, b, blk);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory must outlive the reads
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((unsigned)BN) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn4)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// NHWC activation as a 4-D tensor {C, W, H, B}; box {32 channels, 16 x, 18 y, 1}, 128-byte swizzle, zero fill outside
static int tma_map_nhwc_halo(CUtensorMap* tm, const float* base, int B, int H, int W, int C) {
    static EncodeTiledFn4 enc = nullptr;
    if (enc == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            enc = reinterpret_cast<EncodeTiledFn4>(p);
    }
    if (enc == nullptr) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return OCR_ECUDA; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    cuuint32_t box[4] = {(cuuint32_t)kGemmBK, (cuuint32_t)kHaloCols, (cuuint32_t)kHaloRows, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (NHWC halo) failed (%d) B=%d H=%d W=%d C=%d", (int)r, B, H, W, C); return OCR_ECUDA; }
    return OCR_OK;
}

// NHWC output [B, Ho, Wo, C] as a 4-D tensor {C, Wo, Ho, B}; box {32 channels, bx, by, 1}, 128-byte swizzle (TMA-store epilogue)
static int tma_map_nhwc_out(CUtensorMap* tm, float* base, int B, int Ho, int Wo, int C, int bx, int by) {
    EncodeTiledFn4 enc = nullptr;
    {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            enc = reinterpret_cast<EncodeTiledFn4>(p);
    }
    if (enc == nullptr) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return OCR_ECUDA; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)Wo * C * 4, (cuuint64_t)Ho * Wo * C * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)bx, (cuuint32_t)by, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (NHWC output) failed (%d) B=%d H=%d W=%d C=%d", (int)r, B, Ho, Wo, C); return OCR_ECUDA; }
    return OCR_OK;
}

static int g_halo_tma_store = 1;   // TMA-store epilogue (0: one STG per lane and 16 bytes; pooling always stores through TMA)
int conv_halo_set_tma_store(int on) {
    g_halo_tma_store = on ? 1 : 0;
    return OCR_OK;
}

template <int BN, int WSTAGES, int POOL>
static int launch_halo(const float* in, int B, int H, int W, int C, const float* w, const float* bias, int Cout, int relu, float* out,
                       cudaStream_t st)
{
    CUtensorMap tmIn, tmW, tmOut;
    int rc = tma_map_nhwc_halo(&tmIn, in, B, H, W, C);
    if (rc != OCR_OK) return rc;
    // output tensor: the layer's own, or the pooled one ('valid' pooling: floor)
    const int Ho = POOL == 0 ? H : (H - 2) / 2 + 1, Wo = POOL == 0 ? W : (POOL == 1 ? (W - 2) / 2 + 1 : W - 1);
    const int tma_store = (POOL != 0 || g_halo_tma_store) ? 1 : 0;
    rc = tma_map_nhwc_out(&tmOut, out, B, Ho, Wo, Cout, POOL == 0 ? 8 : (POOL == 1 ? 4 : 7), POOL == 0 ? 4 : 2);
    if (rc != OCR_OK) return rc;
    rc = tma_map_2d(&tmW, w, Cout, 9 * C, 9 * C, BN);
    if (rc != OCR_OK) return rc;
    const int cpt = C / kGemmBK;
    const size_t smem = (size_t)cpt * kHaloBytes + (size_t)WSTAGES * BN * kGemmBK * 4 + (2 * WSTAGES + 2) * 8 + 16 + 1024;
    static int configured = -1;
    int dev = 0;
    OCR_CHECK_CUDA(cudaGetDevice(&dev));
    if (configured != dev) {
        OCR_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel<BN, WSTAGES, POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
        configured = dev;
    }
    // POOL 2: a patch yields 7 pooled columns (the 8th output column is the right-hand neighbour of the 7th)
    const int tiles_x = POOL == 2 ? (W - 1 + kHaloTX - 2) / (kHaloTX - 1) : (W + kHaloTX - 1) / kHaloTX, tiles_y = (H + kHaloTY - 1) / kHaloTY;
    dim3 grid((unsigned)((long long)tiles_x * tiles_y * B), (unsigned)((Cout + BN - 1) / BN));
    conv3x3_halo_kernel<BN, WSTAGES, POOL><<<grid, kHaloThreads, smem, st>>>(tmIn, tmW, tmOut, B, H, W, C, bias, out, Cout, relu, tiles_x, tiles_y, tma_store);
    OCR_CHECK_LAUNCH();
    return OCR_OK;
}

// Is the halo kernel the better choice for this layer?  Shallow inputs (the gather of conv_igemm.cu is L2-bound), enough rows
// to fill the 16-row patch, channel counts whose halo tiles fit beside two more CTAs.
bool conv_halo_supported(int B, int H, int W, int C, int Cout) {
    if (C != 32 && C != 64) return false;
    if ((Cout % 4) != 0 || H < 12 || W < 8) return false;
    const long long tiles = (long long)((W + kHaloTX - 1) / kHaloTX) * ((H + kHaloTY - 1) / kHaloTY) * B;
    return tiles < 0x7fffffffLL;
}

// pool: 0 = none, 1 = max-pool 2x2 stride (2,2), 2 = max-pool 2x2 stride (2,1) applied to the layer's output in the epilogue
int conv_halo_run(const float* in, int B, int H, int W, int C, const float* w, const float* bias, int Cout, int relu, float* out, cudaStream_t st,
                  int pool) {
    if (pool == 1) return Cout <= 32 ? launch_halo<32, 4, 1>(in, B, H, W, C, w, bias, Cout, relu, out, st) : launch_halo<64, 4, 1>(in, B, H, W, C, w, bias, Cout, relu, out, st);
    if (pool == 2) return Cout <= 32 ? launch_halo<32, 4, 2>(in, B, H, W, C, w, bias, Cout, relu, out, st) : launch_halo<64, 4, 2>(in, B, H, W, C, w, bias, Cout, relu, out, st);
    if (Cout <= 32) return launch_halo<32, 4, 0>(in, B, H, W, C, w, bias, Cout, relu, out, st);
    return launch_halo<64, 4, 0>(in, B, H, W, C, w, bias, Cout, relu, out, st);
}

}  // namespace ocr
