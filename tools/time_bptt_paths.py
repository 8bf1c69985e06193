"""BPTT of one LSTM layer (H=512, T=125): frame-by-frame launches against the persistent kernel: python tools/time_bptt_paths.py [B ...]"""
import sys, ctypes
sys.path.insert(0, ".")
import torch
from cnn_lstm_ctc_ocr_b200 import _lib
lib = _lib.load()
T, H = 125, 512
dev = torch.device("cuda:0")
for B in [int(a) for a in sys.argv[1:]] or [32, 64, 128, 256]:
    g = torch.Generator(device=dev); g.manual_seed(0)
    act = torch.rand((T * B, 8 * H), device=dev, generator=g) * 0.8 + 0.1
    cs = torch.randn((T, B, 2 * H), device=dev, generator=g) * 0.5
    dout = torch.randn((T, B, 2 * H), device=dev, generator=g) * 0.01
    wh_rows = torch.randn((2 * H, 4 * H), device=dev, generator=g) * 0.02
    sl = torch.full((B,), T, dtype=torch.int32, device=dev)
    need = ctypes.c_size_t(0)
    lib.ocr_birnn_lstm_train_workspace_bytes(T, B, H, ctypes.byref(need))
    ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
    a = act.clone()
    out = {}
    for path in (1, 129, 257, 3):
        lib.ocr_birnn_set_path(min(path, 3) if path <= 3 else 1)
        lib.ocr_debug_bptt_pdl(1 if path <= 3 else path)
        def run():
            a.copy_(act)
            _lib.check(lib.ocr_birnn_lstm_bwd(_lib.ptr(dout), T, B, H, _lib.ptr(sl), _lib.ptr(a), _lib.ptr(cs), _lib.ptr(wh_rows), _lib.ptr(ws), need.value,
                                              _lib.stream_handle()), "bwd")
        for _ in range(2): run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run()
        e1.record(); torch.cuda.synchronize()
        out[path] = (e0.elapsed_time(e1) / 5, a.clone())
    lib.ocr_birnn_set_path(0); lib.ocr_debug_bptt_pdl(1)
    print("   tile width 128: %.2f us/frame, 256: %.2f us/frame" % (out[129][0] * 1e3 / T, out[257][0] * 1e3 / T))
    d = (out[1][1] - out[3][1]).abs().max().item() / out[1][1].abs().max().item()
    print("B=%3d: frame-by-frame %.3f ms (%.2f us/frame), persistent %.3f ms (%.2f us/frame), max rel diff %.2e" % (
        B, out[1][0], out[1][0] * 1e3 / T, out[3][0], out[3][0] * 1e3 / T, d), flush=True)
