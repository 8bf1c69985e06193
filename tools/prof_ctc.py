#!/usr/bin/env python
"""Small driver for ncu captures of the CTC loss kernel: `python tools/prof_ctc.py [B] [T] [C] [reps]`.
Runs ocr_ctc_loss on one synthetic batch (same generator as bench.py's bandwidth regime)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from cnn_lstm_ctc_ocr_b200 import _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
C = int(sys.argv[3]) if len(sys.argv) > 3 else 63
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
dev = torch.device("cuda:0")
lib = _lib.load()
g = torch.Generator(device=dev)
g.manual_seed(7)
x = torch.randn((T, B, C), device=dev, generator=g)
sl = torch.randint(T // 2, T + 1, (B,), device=dev, generator=g, dtype=torch.int32)
lens = torch.minimum(torch.randint(1, 17, (B,), device=dev, generator=g, dtype=torch.int32), sl // 2).clamp_(min=1)
off = torch.zeros(B + 1, dtype=torch.int32, device=dev)
off[1:] = torch.cumsum(lens, 0)
flat = torch.randint(0, C - 1, (int(off[-1].item()),), device=dev, generator=g, dtype=torch.int32)
loss = torch.empty(B, device=dev)
grad = torch.empty_like(x)
status = torch.empty(B, dtype=torch.int32, device=dev)
need = ctypes.c_size_t(0)
_lib.check(lib.ocr_ctc_loss_workspace_bytes(T, B, C, 16, ctypes.byref(need)), "ws")
ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=dev)
sh = _lib.stream_handle()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(reps):
    if i == reps - 1:
        e0.record()
    _lib.check(lib.ocr_ctc_loss(_lib.ptr(x), T, B, C, _lib.ptr(flat), _lib.ptr(off), _lib.ptr(sl), 16, _lib.ptr(loss),
                                _lib.ptr(grad), _lib.ptr(status), 1.0 / B, _lib.ptr(ws), need.value, sh), "ocr_ctc_loss")
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3
print("B=%d T=%d C=%d: last call %.1f us, %.1f GB/s algorithmic, redo=%d, loss mean %.4f"
      % (B, T, C, us, 2 * T * B * C * 4 / us / 1e3, int((status == 100).sum()), float(loss.mean())))
