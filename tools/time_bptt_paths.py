"""BPTT of one LSTM layer (H=512, T=125): frame-by-frame launches (16-bit and TF32 operands of the recurrent product) against the
persistent kernel: python tools/time_bptt_paths.py [B ...]"""
import sys, ctypes
sys.path.insert(0, ".")
import torch
from cnn_lstm_ctc_ocr_b200 import _lib
lib = _lib.load()
T, H = 125, 512
dev = torch.device("cuda:0")
modes = [("frames, bf16", 1, 1), ("frames, tf32", 1, 3), ("persistent", 3, 1)]
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
    for name, path, flags in modes:
        lib.ocr_birnn_set_path(path)
        lib.ocr_debug_bptt_pdl(flags)
        def run():
            a.copy_(act)
            _lib.check(lib.ocr_birnn_lstm_bwd(_lib.ptr(dout), T, B, H, _lib.ptr(sl), _lib.ptr(a), _lib.ptr(cs), _lib.ptr(wh_rows), _lib.ptr(ws), need.value,
                                              _lib.stream_handle()), "bwd")
        for _ in range(2): run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run()
        e1.record(); torch.cuda.synchronize()
        out[name] = (e0.elapsed_time(e1) / 5, a.clone())
    lib.ocr_birnn_set_path(0); lib.ocr_debug_bptt_pdl(1)
    ref = out["frames, tf32"][1]
    print("B=%3d: " % B + "; ".join("%s %.2f us/frame (max rel diff to tf32 frames %.1e)" % (n, out[n][0] * 1e3 / T, (out[n][1] - ref).abs().max().item() / ref.abs().max().item())
                                     for n, _, _ in modes), flush=True)
