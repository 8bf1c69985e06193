#!/usr/bin/env python
"""L2 prefetch distance of ctc_loss_fast_kernel (tuning aid): python tools/time_ctc_prefetch.py [B] [strides...]
Times ocr_ctc_loss in the bandwidth regime for each distance (0 = off, -1 = automatic = resident CTAs)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cnn_lstm_ctc_ocr_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
strides = [int(a) for a in sys.argv[2:]] or [0, -1, 148, 296, 444, 592]
T, C = 64, 63
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def go():
    _lib.check(lib.ocr_ctc_loss(_lib.ptr(x), T, B, C, _lib.ptr(flat), _lib.ptr(off), _lib.ptr(sl), 16, _lib.ptr(loss),
                                _lib.ptr(grad), _lib.ptr(status), 1.0 / B, _lib.ptr(ws), need.value, _lib.stream_handle()), "ctc")
ref = None
for s in strides:
    spec = 1
    if s <= -2:      # -2: automatic distance, speculative box requests off
        spec, s = 0, -1
    lib.ocr_debug_ctc_speculate(spec)
    _lib.check(lib.ocr_debug_ctc_prefetch(s), "prefetch")
    go(); go(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); go(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    us = sorted(ts)[len(ts) // 2]
    if ref is None:
        ref = (loss.clone(), grad.clone())
    same = torch.equal(ref[0], loss) and torch.equal(ref[1], grad)
    print("speculate %d prefetch %4d: %.1f us, %.1f GB/s algorithmic, bit-identical to the first setting: %s" % (spec, s, us, 2 * T * B * C * 4 / us / 1e3, same))
lib.ocr_debug_ctc_prefetch(-1); lib.ocr_debug_ctc_speculate(1)
