#!/usr/bin/env python
"""cfg2 (B=256, T=64, C=63) CTC loss+gradient as a CUDA graph of K calls over a ring of distinct batches larger than L2,
with programmatic dependent launch on and off (tuning aid): python tools/time_ctc_cfg2.py [B] [K]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cnn_lstm_ctc_ocr_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 100
T, C = 64, 63
dev = torch.device("cuda:0"); lib = _lib.load()
g = torch.Generator(device=dev); g.manual_seed(7)
ring_n = max(2, -(-int(2.5 * 126e6) // (2 * T * B * C * 4)))
ring = []
for i in range(ring_n):
    x = torch.randn((T, B, C), device=dev, generator=g)
    sl = torch.randint(T // 2, T + 1, (B,), device=dev, generator=g, dtype=torch.int32)
    lens = torch.minimum(torch.randint(1, 17, (B,), device=dev, generator=g, dtype=torch.int32), sl // 2).clamp_(min=1)
    off = torch.zeros(B + 1, dtype=torch.int32, device=dev); off[1:] = torch.cumsum(lens, 0)
    flat = torch.randint(0, C - 1, (int(off[-1].item()),), device=dev, generator=g, dtype=torch.int32)
    ring.append(dict(x=x, sl=sl, off=off, flat=flat, loss=torch.empty(B, device=dev), grad=torch.empty_like(x),
                     status=torch.empty(B, dtype=torch.int32, device=dev)))
need = ctypes.c_size_t(0); lib.ocr_ctc_loss_workspace_bytes(T, B, C, 16, ctypes.byref(need))
ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=dev)
def call(r, sh):
    _lib.check(lib.ocr_ctc_loss(_lib.ptr(r["x"]), T, B, C, _lib.ptr(r["flat"]), _lib.ptr(r["off"]), _lib.ptr(r["sl"]), 16, _lib.ptr(r["loss"]),
                                _lib.ptr(r["grad"]), _lib.ptr(r["status"]), 1.0 / B, _lib.ptr(ws), need.value, sh), "ctc")
stream = torch.cuda.Stream(device=dev)
ref = None
for pdl, inl in ((1, 0), (1, 1), (0, 1), (2, 1), (1, 0), (1, 1), (2, 1)):
    lib.ocr_debug_ctc_inline_redo(inl)
    lib.ocr_debug_ctc_pdl(pdl)
    with torch.cuda.stream(stream):
        sh = _lib.stream_handle()
        for i in range(3): call(ring[i % ring_n], sh)
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for i in range(K): call(ring[i % ring_n], _lib.stream_handle())
        graph.replay(); stream.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); graph.replay(); e1.record(stream); stream.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / K)
    us = sorted(ts)[len(ts) // 2]
    out = (ring[0]["loss"].clone(), ring[0]["grad"].clone(), ring[(K - 1) % ring_n]["grad"].clone())
    if ref is None: ref = out
    same = all(torch.equal(a, b) for a, b in zip(ref, out))
    print("inline_redo %d pdl %d: %.2f us per call (graph of %d), %.1f GB/s algorithmic, outputs identical to the first run: %s, flagged %d"
          % (inl, pdl, us, K, 2 * T * B * C * 4 / us / 1e3, same, int((ring[0]["status"] == 100).sum())))
lib.ocr_debug_ctc_pdl(1); lib.ocr_debug_ctc_inline_redo(1)
