"""Per-frame phase timeline of the persistent LSTM kernel (CTA 0): python tools/timeline_lstm.py [B] [T] [f16 operands: 1|0]"""
import sys, ctypes
sys.path.insert(0, ".")
import numpy as np, torch
from cnn_lstm_ctc_ocr_b200 import _lib
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 61
I, H = 256, 512
F16 = int(sys.argv[3]) if len(sys.argv) > 3 else 1
lib.ocr_debug_lstm_operands(F16)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(0)
x = torch.randn((T, B, I), device=dev, generator=g)
wx = torch.randn((8 * H, I), device=dev, generator=g) * 0.05
wh = torch.randn((8 * H, H), device=dev, generator=g) * 0.05
bias = torch.zeros(8 * H, device=dev)
sl = torch.full((B,), T, dtype=torch.int32, device=dev)
out = torch.empty((T, B, 2 * H), device=dev)
need = ctypes.c_size_t(0)
lib.ocr_birnn_workspace_bytes(0, T, B, H, ctypes.byref(need))
ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
wh2 = torch.empty_like(wh)
_lib.check(lib.ocr_lstm_prepare_wh(_lib.ptr(wh), H, _lib.ptr(wh2), _lib.stream_handle()), "prep")
def run():
    _lib.check(lib.ocr_birnn_layer(0, _lib.ptr(x), T, B, I, H, _lib.ptr(sl), _lib.ptr(wx), _lib.ptr(wh), _lib.ptr(wh2), _lib.ptr(bias), _lib.ptr(out),
                                   _lib.ptr(ws), need.value, _lib.stream_handle()), "layer")
for _ in range(3): run()
tl = torch.zeros(T * 8, dtype=torch.int64, device=dev)
lib.ocr_debug_lstm_timeline(_lib.ptr(tl))
run(); torch.cuda.synchronize()
lib.ocr_debug_lstm_timeline(None)
a = tl.cpu().numpy().reshape(T, 8).astype(np.float64)
names = ["barrier->", "tma issued", "1st tile landed", "mma issued", "acc done", "tmem read", "cell+stores", "published"]
fr = a[5:T - 2]
print("operands: %s" % ("binary16" if F16 else "tf32"))
print("B=%d T=%d: frame period %.0f cycles (%.2f us at 1.965 GHz)" % (B, T, np.diff(a[5:T - 2, 0]).mean(), np.diff(a[5:T - 2, 0]).mean() / 1965))
for i in range(1, 8):
    print("  %-16s +%6.0f cycles after the previous mark" % (names[i], (fr[:, i] - fr[:, i - 1]).mean()))
print("  %-16s +%6.0f cycles (published -> next frame's barrier passed)" % ("grid barrier", (a[6:T - 1, 0] - a[5:T - 2, 7]).mean()))
