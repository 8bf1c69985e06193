#!/usr/bin/env python
"""Phase timeline of ctc_loss_stream_kernel (tuning aid): python tools/timeline_ctc_stream.py [B] [T] [C]."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from cnn_lstm_ctc_ocr_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
C = int(sys.argv[3]) if len(sys.argv) > 3 else 63
G = 4
W = 2 * G
dev = torch.device("cuda:0"); lib = _lib.load()
lib.ocr_ctc_loss_set_path(8)   # the streaming kernel is opt-in
g = torch.Generator(device=dev); g.manual_seed(7)
x = torch.randn((T, B, C), device=dev, generator=g)
sl = torch.randint(T // 2, T + 1, (B,), device=dev, generator=g, dtype=torch.int32)
lens = torch.minimum(torch.randint(1, 17, (B,), device=dev, generator=g, dtype=torch.int32), sl // 2).clamp_(min=1)
off = torch.zeros(B + 1, dtype=torch.int32, device=dev); off[1:] = torch.cumsum(lens, 0)
flat = torch.randint(0, C - 1, (int(off[-1].item()),), device=dev, generator=g, dtype=torch.int32)
loss = torch.empty(B, device=dev); grad = torch.empty_like(x); status = torch.empty(B, dtype=torch.int32, device=dev)
need = ctypes.c_size_t(0); lib.ocr_ctc_loss_workspace_bytes(T, B, C, 16, ctypes.byref(need))
ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=dev)
nwarps = B // G * W
tlb = torch.zeros(nwarps * 12, dtype=torch.int64, device=dev)
def go():
    _lib.check(lib.ocr_ctc_loss(_lib.ptr(x), T, B, C, _lib.ptr(flat), _lib.ptr(off), _lib.ptr(sl), 16, _lib.ptr(loss),
                                _lib.ptr(grad), _lib.ptr(status), 1.0 / B, _lib.ptr(ws), need.value, _lib.stream_handle()), "ctc")
go(); go(); torch.cuda.synchronize()
_lib.check(lib.ocr_debug_ctc_timeline(_lib.ptr(tlb)), "timeline")
go(); torch.cuda.synchronize()
lib.ocr_debug_ctc_timeline(None)
tl = tlb.cpu().numpy().reshape(-1, W, 12).astype(np.float64)   # [cta, warp, slot]
t0 = tl[:, :, 0].min(1)
life = tl[:, :, 10].max(1) - t0
print("CTAs %d, mean CTA lifetime %.0f cycles (min %.0f max %.0f)" % (tl.shape[0], life.mean(), life.min(), life.max()))
rel = tl - t0[:, None, None]
comp = rel[:, :2 * G]
names = {11: "first box landed", 1: "front done", 2: "e block complete", 3: "ring free", 4: "stored half", 5: "partner met", 6: "consumed half",
         7: "chains done", 8: "boxes stored (wait)", 9: "fix-up done", 10: "CTA barrier"}
for role, nm in ((0, "alpha"), (1, "beta")):
    print("-- %s warps: mean time since CTA start (cycles), p90" % nm)
    for k in (11, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10):
        v = comp[:, role::2, k]; v = v[v > 0]
        if v.size: print("   %-24s %8.0f %8.0f" % (names[k], v.mean(), np.percentile(v, 90)))
