#!/usr/bin/env python
"""Phase timeline of ctc_loss_fast_kernel (tuning aid): python tools/timeline_ctc.py [B] [T] [C]."""
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
dev = torch.device("cuda:0"); lib = _lib.load()
g = torch.Generator(device=dev); g.manual_seed(7)
x = torch.randn((T, B, C), device=dev, generator=g)
sl = torch.randint(T // 2, T + 1, (B,), device=dev, generator=g, dtype=torch.int32)
lens = torch.minimum(torch.randint(1, 17, (B,), device=dev, generator=g, dtype=torch.int32), sl // 2).clamp_(min=1)
off = torch.zeros(B + 1, dtype=torch.int32, device=dev); off[1:] = torch.cumsum(lens, 0)
flat = torch.randint(0, C - 1, (int(off[-1].item()),), device=dev, generator=g, dtype=torch.int32)
loss = torch.empty(B, device=dev); grad = torch.empty_like(x); status = torch.empty(B, dtype=torch.int32, device=dev)
need = ctypes.c_size_t(0); lib.ocr_ctc_loss_workspace_bytes(T, B, C, 16, ctypes.byref(need))
ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=dev)
nwarps = (B + G - 1) // G * 2 * G
tlb = torch.zeros(nwarps * 12 * 2, dtype=torch.int64, device=dev)  # x2 slack in case the plan picks a smaller G
def go():
    _lib.check(lib.ocr_ctc_loss(_lib.ptr(x), T, B, C, _lib.ptr(flat), _lib.ptr(off), _lib.ptr(sl), 16, _lib.ptr(loss),
                                _lib.ptr(grad), _lib.ptr(status), 1.0 / B, _lib.ptr(ws), need.value, _lib.stream_handle()), "ctc")
go(); go(); torch.cuda.synchronize()
_lib.check(lib.ocr_debug_ctc_timeline(_lib.ptr(tlb)), "timeline")
go(); torch.cuda.synchronize()
lib.ocr_debug_ctc_timeline(None)
tl = tlb.cpu().numpy()[: nwarps * 12].reshape(-1, 2 * G, 12)   # [cta, warp, slot]
names = ["launch->loads landed", "pass1", "wait partner(p1)", "store half", "wait partner(mid)", "consume half",
         "wait partner(end)", "pass3+detect", "zero rows", "wait CTA", "bulk store issue+drain"]
ok = tl[:, :, 11] > 0
d = np.diff(tl, axis=2).astype(np.float64)
print("CTAs %d, mean CTA lifetime %.0f cycles (min %.0f max %.0f)" % (tl.shape[0], (tl[:, :, 11].max(1) - tl[:, :, 0].min(1)).mean(),
      (tl[:, :, 11].max(1) - tl[:, :, 0].min(1)).min(), (tl[:, :, 11].max(1) - tl[:, :, 0].min(1)).max()))
for i, n in enumerate(names):
    a = d[:, 0::2, i][ok[:, 0::2]]; b = d[:, 1::2, i][ok[:, 1::2]]
    print("%-26s alpha warp %8.0f  beta warp %8.0f   (p90 %6.0f / %6.0f)" % (n, a.mean(), b.mean(), np.percentile(a, 90), np.percentile(b, 90)))
