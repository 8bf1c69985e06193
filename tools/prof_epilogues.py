"""Driver for ncu captures of the epilogue work of round 2: the RNN input projection (ocr_gemm_tf32) and conv2 / conv6 / conv8 of the
training step's shapes, first with the TMA-store epilogues, then with the STG epilogues; then conv2 + pool2 fused and unfused:
python tools/prof_epilogues.py"""
import sys
sys.path.insert(0, ".")
import os
import torch
from cnn_lstm_ctc_ocr_b200 import _lib
lib = _lib.load()
REPS = int(os.environ.get("PROF_REPS", "2"))
dev = torch.device("cuda:0")
sh = _lib.stream_handle()
M, N, K = 32000, 4096, 256
A = torch.randn((M, K), device=dev); Wm = torch.randn((N, K), device=dev) * 0.05; bias = torch.randn(N, device=dev); D = torch.empty((M, N), device=dev)
convs = [(256, 30, 254, 32, 32), (256, 7, 126, 128, 128), (256, 3, 125, 256, 256)]
cx = [(torch.randn((B, H, W, C), device=dev), torch.randn((Co, 9 * C), device=dev) * 0.05, torch.randn(Co, device=dev), torch.empty((B, H, W, Co), device=dev)) for B, H, W, C, Co in convs]
for mode in (1, 0):
    lib.ocr_debug_gemm_tma_store(mode); lib.ocr_debug_conv_tma_store(mode)
    for _ in range(REPS):
        _lib.check(lib.ocr_gemm_tf32(_lib.ptr(A), K, _lib.ptr(Wm), K, _lib.ptr(bias), _lib.ptr(D), N, M, N, K, 0, sh), "gemm")
        for (B, H, W, C, Co), (x, w, b, o) in zip(convs, cx):
            _lib.check(lib.ocr_conv3x3_same(_lib.ptr(x), B, H, W, C, _lib.ptr(w), _lib.ptr(b), Co, 1, _lib.ptr(o), sh), "conv")
lib.ocr_debug_gemm_tma_store(1); lib.ocr_debug_conv_tma_store(1)
# conv2 + pool2 at the inference shape (B = 32, 30 x 126, 32 -> 32): fused, then conv followed by ocr_maxpool
B, H, W, C, Co = 32, 30, 126, 32, 32
x = torch.randn((B, H, W, C), device=dev); w = torch.randn((Co, 9 * C), device=dev) * 0.05; b = torch.randn(Co, device=dev)
full = torch.empty((B, H, W, Co), device=dev); pooled = torch.empty((B, H // 2, W // 2, Co), device=dev)
for _ in range(REPS):
    _lib.check(lib.ocr_conv3x3_same_pool(_lib.ptr(x), B, H, W, C, _lib.ptr(w), _lib.ptr(b), Co, 1, 2, _lib.ptr(pooled), sh), "fused")
    _lib.check(lib.ocr_conv3x3_same(_lib.ptr(x), B, H, W, C, _lib.ptr(w), _lib.ptr(b), Co, 1, _lib.ptr(full), sh), "conv")
    _lib.check(lib.ocr_maxpool(_lib.ptr(full), B, H, W, Co, 2, 2, 2, 2, _lib.ptr(pooled), sh), "pool")
torch.cuda.synchronize()
print("ok")
