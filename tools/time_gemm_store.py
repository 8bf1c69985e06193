"""ocr_gemm_tf32 with the TMA-store epilogue against the STG epilogue: same bits, time per call: python tools/time_gemm_store.py"""
import sys
sys.path.insert(0, ".")
import torch
from cnn_lstm_ctc_ocr_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
shapes = [(32000, 4096, 256, 0), (32000, 4096, 1024, 0), (32000, 1024, 4096, 0), (32000, 256, 4096, 0), (32000, 96, 1024, 1), (1952, 4096, 256, 0), (1952, 96, 1024, 1),
          (4003, 4096, 1024, 0), (777, 100, 64, 1), (130, 36, 40, 0)]
for M, N, K, relu in shapes:
    g = torch.Generator(device=dev); g.manual_seed(M + N)
    A = torch.randn((M, K), device=dev, generator=g)
    W = torch.randn((N, K), device=dev, generator=g) * 0.05
    bias = torch.randn(N, device=dev, generator=g)
    res = {}
    for on in (1, 0):
        lib.ocr_debug_gemm_tma_store(on)
        D = torch.full((M, N), float("nan"), device=dev)
        def run():
            _lib.check(lib.ocr_gemm_tf32(_lib.ptr(A), K, _lib.ptr(W), K, _lib.ptr(bias), _lib.ptr(D), N, M, N, K, relu, _lib.stream_handle()), "gemm")
        for _ in range(3): run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        res[on] = (e0.elapsed_time(e1) * 100, D.clone())
    print("M=%5d N=%4d K=%4d: TMA store %.1f us, STG %.1f us, same bits: %s" % (M, N, K, res[1][0], res[0][0], torch.equal(res[1][1], res[0][1])), flush=True)
lib.ocr_debug_gemm_tma_store(1)
