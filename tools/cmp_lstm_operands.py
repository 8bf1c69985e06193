"""binary16 against TF32 recurrent operands of the persistent LSTM kernel: outputs of one layer (both against a float64 torch
recursion) and the time per layer: python tools/cmp_lstm_operands.py [B] [T] [H] [I]"""
import sys, ctypes
sys.path.insert(0, ".")
import numpy as np, torch
from cnn_lstm_ctc_ocr_b200 import _lib
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 61
H = int(sys.argv[3]) if len(sys.argv) > 3 else 512
I = int(sys.argv[4]) if len(sys.argv) > 4 else 256
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(0)
x = torch.randn((T, B, I), device=dev, generator=g)
wx = torch.randn((8 * H, I), device=dev, generator=g) * 0.05
wh = torch.randn((8 * H, H), device=dev, generator=g) * 0.05
bias = torch.zeros(8 * H, device=dev)
sl = torch.randint(T // 2, T + 1, (B,), dtype=torch.int32, device=dev)
need = ctypes.c_size_t(0)
lib.ocr_birnn_workspace_bytes(0, T, B, H, ctypes.byref(need))
ws = torch.empty(need.value, dtype=torch.uint8, device=dev)

def ref64():
    xd, wxd, whd = x.double(), wx.double(), wh.double()
    out = torch.zeros((T, B, 2 * H), dtype=torch.float64, device=dev)
    for d in range(2):
        h = torch.zeros((B, H), dtype=torch.float64, device=dev); c = torch.zeros_like(h)
        for s in range(T):
            t = (sl.long() - 1 - s) if d else torch.full((B,), s, device=dev)
            live = (s < sl).unsqueeze(1)
            tt = t.clamp(min=0)
            xt = xd[tt, torch.arange(B, device=dev)]
            z = xt @ wxd[d * 4 * H:(d + 1) * 4 * H].T + h @ whd[d * 4 * H:(d + 1) * 4 * H].T
            zi, zj, zf, zo = z.split(H, dim=1)
            cn = torch.sigmoid(zf + 1.0) * c + torch.sigmoid(zi) * torch.tanh(zj)
            hn = torch.sigmoid(zo) * torch.tanh(cn)
            c = torch.where(live, cn, c); h = torch.where(live, hn, h)
            idx = torch.nonzero(live.squeeze(1)).squeeze(1)
            out[tt[idx], idx, d * H:(d + 1) * H] = hn[idx]
    return out
want = ref64()
res = {}
for f16 in (1, 0):
    lib.ocr_debug_lstm_operands(f16)
    wh2 = torch.empty_like(wh)
    _lib.check(lib.ocr_lstm_prepare_wh(_lib.ptr(wh), H, _lib.ptr(wh2), _lib.stream_handle()), "prep")
    out = torch.empty((T, B, 2 * H), device=dev)
    def run():
        _lib.check(lib.ocr_birnn_layer(0, _lib.ptr(x), T, B, I, H, _lib.ptr(sl), _lib.ptr(wx), _lib.ptr(wh), _lib.ptr(wh2), _lib.ptr(bias), _lib.ptr(out),
                                       _lib.ptr(ws), need.value, _lib.stream_handle()), "layer")
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    err = (out.double() - want).abs().max().item()
    print("%s operands: %.1f us per layer call (input projection included), max |out - float64| = %.3e" % ("binary16" if f16 else "tf32    ", e0.elapsed_time(e1) * 50, err), flush=True)
    res[f16] = out.clone()
print("max |binary16 - tf32| = %.3e" % (res[1] - res[0]).abs().max().item())
lib.ocr_debug_lstm_operands(1)
