"""Filter-gradient contraction of conv2 / conv3 / conv4 at the cfg3 shapes (planar operands, nine tap views): python tools/time_wgrad.py [flags ...]
flags = values for ocr_debug_gemm_tma_store (1 default, 3 one tile per tap, 5 256-byte L2 promotion)"""
import sys, ctypes
sys.path.insert(0, ".")
import torch
from cnn_lstm_ctc_ocr_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
sh = _lib.stream_handle()
flags = [int(a) for a in sys.argv[1:]] or [1, 3, 5]
for name, B, H, W, C, Co in [("conv2", 256, 30, 254, 32, 32), ("conv3", 256, 15, 127, 32, 64), ("conv4", 256, 15, 127, 64, 64), ("conv6", 256, 7, 126, 128, 128)]:
    Wp = lib.ocr_planar_pad_pitch(W)
    R = B * (H + 2) * Wp
    xt = torch.randn((3, C, R), device=dev)          # three pixel-shifted planar copies
    dyt = torch.randn((Co, R), device=dev)
    D = torch.zeros((9, C, Co), device=dev)
    shifts = (ctypes.c_int32 * 9)(*[(t // 3 - 1) * Wp for t in range(9)])
    rows = (ctypes.c_int32 * 9)(*[(t % 3) * C for t in range(9)])
    line = "%s wgrad (C_in %d, C_out %d, R = %d):" % (name, C, Co, R)
    for f in flags:
        lib.ocr_debug_gemm_tma_store(f)
        need = ctypes.c_size_t(0)
        _lib.check(lib.ocr_gemm_wgrad_scratch_bytes(C, Co, R, 9, ctypes.byref(need)), "scratch")
        scr = torch.empty(need.value, dtype=torch.uint8, device=dev)
        def run():
            _lib.check(lib.ocr_gemm_tf32_wgrad(_lib.ptr(xt), R, _lib.ptr(dyt), R, _lib.ptr(D), Co, C * Co, C, Co, R, 9, shifts, rows, 3 * C, _lib.ptr(scr), need.value, sh), "wgrad")
        for _ in range(2): run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run()
        e1.record(); torch.cuda.synchronize()
        line += "  flags %d: %.0f us" % (f, e0.elapsed_time(e1) * 200)
    print(line, flush=True)
lib.ocr_debug_gemm_tma_store(1)
