"""GPU tier: the tcgen05/TMA TF32 GEMM (ocr_gemm_tf32) against a plain float64 reference of the same op.
Tolerance: TF32 rounds each operand to a 10-bit mantissa (relative 2^-11 per factor) and accumulates in fp32:
|error| <= ~1e-3 * sum_k |a_k w_k|; the test bounds it by 2e-3 * (|A| @ |W|^T)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(A, W, bias, relu, ldd=None):
    from cnn_lstm_ctc_ocr_b200 import _lib
    lib = _lib.load()
    M, K = A.shape
    N = W.shape[0]
    ldd = ldd or N
    D = torch.full((M, ldd), float("nan"), device=A.device)
    _lib.check(lib.ocr_gemm_tf32(_lib.ptr(A), A.stride(0), _lib.ptr(W), W.stride(0), _lib.ptr(bias), _lib.ptr(D), ldd, M, N, K,
                                 int(relu), _lib.stream_handle()), "ocr_gemm_tf32")
    torch.cuda.synchronize()
    return D


@pytest.mark.parametrize("M,N,K,relu,use_bias", [
    (128, 32, 32, False, False),      # one tile, one k-step
    (128, 64, 288, True, True),       # conv2-like K
    (3780, 32, 288, True, True),      # conv2 of one 32x128 crop
    (945, 64, 576, True, True),       # conv4
    (183 * 32, 256, 2304, True, True),  # conv8, batch 32
    (1952, 4096, 256, False, True),   # BiLSTM layer-1 input projection (T=61, B=32, both directions)
    (1952, 96, 1024, True, True),     # logits
    (1952, 63, 1024, True, True),     # N not a multiple of anything
    (200, 130, 36, False, True),      # ragged everything (K tail zero-filled by TMA)
    (1, 1, 4, False, False),
])
def test_gemm_tf32_matches_reference(M, N, K, relu, use_bias):
    g = torch.Generator(device="cuda")
    g.manual_seed(M * 7 + N)
    A = torch.randn((M, K), device="cuda", generator=g)
    W = torch.randn((N, K), device="cuda", generator=g) * 0.1
    bias = torch.randn((N,), device="cuda", generator=g) if use_bias else None
    D = _gemm(A, W, bias, relu)
    ref = A.double() @ W.double().t()
    if use_bias:
        ref = ref + bias.double()
    if relu:
        ref = ref.clamp_min(0)
    bound = 2e-3 * (A.abs().double() @ W.abs().double().t()) + 1e-6
    err = (D.double() - ref).abs()
    assert torch.isfinite(D).all()
    assert bool((err <= bound).all()), "max err %.3g (bound %.3g)" % (err.max().item(), bound.max().item())


def test_gemm_tf32_strided_operands_and_padding():
    """Row pitches larger than K / N; columns beyond N must stay untouched."""
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    Abuf = torch.randn((300, 72), device="cuda", generator=g)
    Wbuf = torch.randn((50, 80), device="cuda", generator=g)
    A, W = Abuf[:, :68], Wbuf[:, :68]
    D = _gemm(A, W, None, False, ldd=64)
    ref = A.double() @ W.double().t()
    assert (D[:, :50].double() - ref).abs().max() < 0.1
    assert torch.isnan(D[:, 50:]).all()


@pytest.mark.parametrize("M,N,K,relu,ldd", [(32000, 4096, 256, False, 4096), (1952, 96, 1024, True, 96), (777, 100, 64, True, 104), (130, 36, 40, False, 36),
                                            (257, 52, 36, False, 64)])
def test_gemm_tma_store_epilogue_equals_stg_epilogue(M, N, K, relu, ldd):
    """The epilogue that stages 32x32 blocks in shared memory and stores them with cp.async.bulk.tensor writes the same bits as
    the one-STG-per-lane-and-row epilogue, clips at M and N, and leaves the columns between N and the row pitch untouched."""
    from cnn_lstm_ctc_ocr_b200 import _lib as L
    lib = L.load()
    g = torch.Generator(device="cuda")
    g.manual_seed(M + N + K)
    A = torch.randn((M, K), device="cuda", generator=g)
    W = torch.randn((N, K), device="cuda", generator=g) * 0.1
    bias = torch.randn(N, device="cuda", generator=g)
    res = []
    for on in (1, 0):
        L.check(lib.ocr_debug_gemm_tma_store(on), "store path")
        D = torch.full((M + 3, ldd), float("nan"), device="cuda")
        L.check(lib.ocr_gemm_tf32(L.ptr(A), K, L.ptr(W), K, L.ptr(bias), L.ptr(D), ldd, M, N, K, int(relu), L.stream_handle()), "gemm")
        torch.cuda.synchronize()
        res.append(D)
    L.check(lib.ocr_debug_gemm_tma_store(1), "store path")
    assert torch.equal(res[0][:M, :N], res[1][:M, :N])
    assert torch.isnan(res[0][M:]).all() and torch.isnan(res[0][:, N:]).all()
    ref = torch.addmm(bias.double(), A.double(), W.double().t())
    if relu:
        ref = ref.clamp(min=0)
    assert (res[0][:M, :N].double() - ref).abs().max() <= 5e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("M,N,K,relu,use_bias", [
    (128, 32, 64, False, False),      # one tile, one k-step of 64 halves
    (32000, 4096, 512, False, True),  # cfg3 layer-1 input projection (T=125, B=256)
    (4000, 4096, 1024, False, True),  # layer 2 (B=32)
    (1952, 96, 1024, True, True),
    (200, 130, 40, False, True),      # ragged M / N, K tail zero-filled by TMA
])
def test_gemm_f16_matches_reference(M, N, K, relu, use_bias):
    """ocr_gemm_f16 (binary16 operands, float32 sums): exact products of the ROUNDED operands up to float32 accumulation
    (1e-5 of sum |a||w|), and the TF32 bar of this file (2e-3) against the unrounded float32 operands; ocr_float_to_half is
    round-to-nearest-even (torch's .half())."""
    from cnn_lstm_ctc_ocr_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda")
    g.manual_seed(M * 7 + N + 1)
    A = torch.randn((M, K), device="cuda", generator=g).abs()      # layer inputs are ReLU / bounded outputs
    W = torch.randn((N, K), device="cuda", generator=g) * 0.1
    bias = torch.randn((N,), device="cuda", generator=g) if use_bias else None
    A16 = torch.empty((M, K), dtype=torch.float16, device="cuda")
    W16 = torch.empty((N, K), dtype=torch.float16, device="cuda")
    sh = _lib.stream_handle()
    if (M * K) % 4 == 0 and (N * K) % 4 == 0:
        _lib.check(lib.ocr_float_to_half(_lib.ptr(A), _lib.ptr(A16), M * K, sh), "to_half")
        _lib.check(lib.ocr_float_to_half(_lib.ptr(W), _lib.ptr(W16), N * K, sh), "to_half")
        torch.cuda.synchronize()
        assert torch.equal(A16, A.half()) and torch.equal(W16, W.half())
    else:
        A16, W16 = A.half(), W.half()
    D = torch.full((M, N), float("nan"), device="cuda")
    _lib.check(lib.ocr_gemm_f16(_lib.ptr(A16), K, _lib.ptr(W16), K, _lib.ptr(bias), _lib.ptr(D), N, M, N, K, int(relu), sh), "ocr_gemm_f16")
    torch.cuda.synchronize()
    assert torch.isfinite(D).all()
    rows = slice(0, M) if M <= 4000 else torch.arange(0, M, 37, device="cuda")       # float64 reference on a row subsample
    def ref_of(a, w):
        r = a[rows].double() @ w.double().t()
        if use_bias:
            r = r + bias.double()
        return r.clamp_min(0) if relu else r
    mag = A[rows].abs().double() @ W.abs().double().t() + 1e-6
    err16 = (D[rows].double() - ref_of(A16, W16)).abs()
    assert bool((err16 <= 1e-5 * mag).all()), "vs rounded operands: max %.3g" % (err16 / mag).max().item()
    err32 = (D[rows].double() - ref_of(A, W)).abs()
    assert bool((err32 <= 2e-3 * mag).all()), "vs float32 operands: max %.3g" % (err32 / mag).max().item()
