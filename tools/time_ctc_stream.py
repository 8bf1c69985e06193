#!/usr/bin/env python
"""CTC loss+gradient: streaming kernel (path 0) against the fast kernel (path 7) at cfg2 (B=256, graph of K calls over a ring
of batches larger than L2) and in the bandwidth regime (B=65536, single launches): python tools/time_ctc_stream.py [nbuf...]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cnn_lstm_ctc_ocr_b200 import _lib
T, C = 64, 63
dev = torch.device("cuda:0"); lib = _lib.load()

def batches(B, n):
    g = torch.Generator(device=dev); g.manual_seed(7)
    out = []
    for i in range(n):
        x = torch.randn((T, B, C), device=dev, generator=g)
        sl = torch.randint(T // 2, T + 1, (B,), device=dev, generator=g, dtype=torch.int32)
        lens = torch.minimum(torch.randint(1, 17, (B,), device=dev, generator=g, dtype=torch.int32), sl // 2).clamp_(min=1)
        off = torch.zeros(B + 1, dtype=torch.int32, device=dev); off[1:] = torch.cumsum(lens, 0)
        flat = torch.randint(0, C - 1, (int(off[-1].item()),), device=dev, generator=g, dtype=torch.int32)
        out.append(dict(x=x, sl=sl, off=off, flat=flat, loss=torch.empty(B, device=dev), grad=torch.empty_like(x),
                        status=torch.empty(B, dtype=torch.int32, device=dev)))
    return out

def run(B, K, configs):
    ring_n = max(2, -(-int(2.5 * 126e6) // (2 * T * B * C * 4)))
    ring = batches(B, min(ring_n, 40))
    need = ctypes.c_size_t(0); lib.ocr_ctc_loss_workspace_bytes(T, B, C, 16, ctypes.byref(need))
    ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=dev)
    def call(r, sh):
        _lib.check(lib.ocr_ctc_loss(_lib.ptr(r["x"]), T, B, C, _lib.ptr(r["flat"]), _lib.ptr(r["off"]), _lib.ptr(r["sl"]), 16, _lib.ptr(r["loss"]),
                                    _lib.ptr(r["grad"]), _lib.ptr(r["status"]), 1.0 / B, _lib.ptr(ws), need.value, sh), "ctc")
    stream = torch.cuda.Stream(device=dev)
    ref = None
    for path, nbuf in configs:
        lib.ocr_ctc_loss_set_path(path); lib.ocr_debug_ctc_stream_nbuf(nbuf)
        with torch.cuda.stream(stream):
            for i in range(3): call(ring[i % len(ring)], _lib.stream_handle())
            stream.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                for i in range(K): call(ring[i % len(ring)], _lib.stream_handle())
            graph.replay(); stream.synchronize()
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); graph.replay(); e1.record(stream); stream.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3 / K)
        us = sorted(ts)[len(ts) // 2]
        out = (ring[0]["loss"].clone(), ring[0]["grad"].clone())
        if ref is None: ref = out
        dl = float((out[0] - ref[0]).abs().max()); dg = float((out[1] - ref[1]).abs().max())
        gbs = 2 * T * B * C * 4 / us / 1e3
        print("B=%d path %d nbuf %d: %.2f us per call, %.1f GB/s algorithmic (frac %.3f of 6555.8), max |dloss| %.2e |dgrad| %.2e vs first config, flagged %d"
              % (B, path, nbuf, us, gbs, gbs / 6555.8, dl, dg, int((ring[0]["status"] == 100).sum())), flush=True)
    lib.ocr_ctc_loss_set_path(0); lib.ocr_debug_ctc_stream_nbuf(0)

nb = [int(a) for a in sys.argv[1:]] or [0]
run(256, 100, [(0, 0), (7, 0)] + [(8, n) for n in nb])
run(65536, 4, [(0, 0), (7, 0)] + [(8, n) for n in nb])
