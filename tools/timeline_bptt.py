"""Per-frame phase timeline of the persistent BPTT kernel (CTA 0): python tools/timeline_bptt.py [B] [T]"""
import sys, ctypes
sys.path.insert(0, ".")
import numpy as np, torch
from cnn_lstm_ctc_ocr_b200 import _lib
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 125
H = 512
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(0)
act = torch.rand((T * B, 8 * H), device=dev, generator=g) * 0.8 + 0.1
cs = torch.randn((T, B, 2 * H), device=dev, generator=g) * 0.5
dout = torch.randn((T, B, 2 * H), device=dev, generator=g) * 0.01
wh_rows = torch.randn((2 * H, 4 * H), device=dev, generator=g) * 0.02
sl = torch.full((B,), T, dtype=torch.int32, device=dev)
need = ctypes.c_size_t(0)
lib.ocr_birnn_lstm_train_workspace_bytes(T, B, H, ctypes.byref(need))
ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
lib.ocr_birnn_set_path(3)
def run():
    a = act.clone()
    _lib.check(lib.ocr_birnn_lstm_bwd(_lib.ptr(dout), T, B, H, _lib.ptr(sl), _lib.ptr(a), _lib.ptr(cs), _lib.ptr(wh_rows), _lib.ptr(ws), need.value,
                                      _lib.stream_handle()), "bwd")
for _ in range(2): run()
tl = torch.zeros(T * 8, dtype=torch.int64, device=dev)
lib.ocr_debug_lstm_timeline(_lib.ptr(tl))
run(); torch.cuda.synchronize()
lib.ocr_debug_lstm_timeline(None); lib.ocr_birnn_set_path(0)
a = tl.cpu().numpy().reshape(T, 8).astype(np.float64)
fr = a[5:T - 3]
per = np.diff(a[5:T - 3, 0]).mean()
print("B=%d T=%d: frame period %.0f cycles (%.2f us at 1.965 GHz)" % (B, T, per, per / 1965))
names = ["barrier passed", "partials summed", "cell + A tile + fence", "MMAs issued (MMA lane)", "accumulator ready", "scatter done", "published"]
for i in range(1, 7):
    print("  %-26s +%6.0f cycles after the previous mark" % (names[i], (fr[:, i] - fr[:, i - 1]).mean()))
print("  %-26s +%6.0f cycles after the barrier (first batch of 16 partial loads per thread summed)" % ("(first batch)", (fr[:, 7] - fr[:, 0]).mean()))
print("  %-26s +%6.0f cycles (published -> next frame's barrier passed)" % ("grid barrier", (a[6:T - 2, 0] - a[5:T - 3, 6]).mean()))
