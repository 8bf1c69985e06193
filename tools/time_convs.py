"""The seven 'same' convolutions of the recognizer at one batch shape, one by one: python tools/time_convs.py [B] [W] [mode]
mode: value for ocr_debug_conv_tma_store (1 = default, 0 = STG epilogues, 3 = shallow gather ring)"""
import sys
sys.path.insert(0, ".")
import torch
from cnn_lstm_ctc_ocr_b200 import _lib
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
W = int(sys.argv[2]) if len(sys.argv) > 2 else 128
modes = [int(a) for a in sys.argv[3:]] or [1, 3, 0]
dev = torch.device("cuda:0")
Wc = W - 2
layers = [("conv2", 30, Wc, 32, 32), ("conv3", 15, Wc // 2, 32, 64), ("conv4", 15, Wc // 2, 64, 64), ("conv5", 7, Wc // 2 - 1, 64, 128),
          ("conv6", 7, Wc // 2 - 1, 128, 128), ("conv7", 3, Wc // 2 - 2, 128, 256), ("conv8", 3, Wc // 2 - 2, 256, 256)]
for name, H, Wl, C, Co in layers:
    x = torch.randn((B, H, Wl, C), device=dev)
    w = torch.randn((Co, 9 * C), device=dev) * 0.05
    b = torch.randn(Co, device=dev)
    out = torch.empty((B, H, Wl, Co), device=dev)
    line = "%s B=%d %dx%d %d->%d:" % (name, B, H, Wl, C, Co)
    ref = None
    for mode in modes:
        lib.ocr_debug_conv_tma_store(mode)
        def run():
            _lib.check(lib.ocr_conv3x3_same(_lib.ptr(x), B, H, Wl, C, _lib.ptr(w), _lib.ptr(b), Co, 1, _lib.ptr(out), _lib.stream_handle()), "conv")
        for _ in range(3): run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        if ref is None: ref = out.clone()
        line += "  mode %d %.1f us%s" % (mode, e0.elapsed_time(e1) * 50, "" if torch.equal(ref, out) else " (DIFFERENT BITS)")
    print(line, flush=True)
lib.ocr_debug_conv_tma_store(1)
