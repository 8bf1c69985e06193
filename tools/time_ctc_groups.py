"""cfg2 step time of ocr_ctc_loss for each group size / load path: python tools/time_ctc_groups.py [B]"""
import sys, ctypes
sys.path.insert(0, ".")
import numpy as np, torch
from cnn_lstm_ctc_ocr_b200 import _lib
import bench
lib = _lib.load()
T, C = 64, 63
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
ring = []
for i in range(40 if B <= 1024 else 2):
    x, flat, off, sl, lens = bench.make_ctc_batch(i, T, B, C)
    xt = torch.from_numpy(x).to(dev)
    ring.append(dict(x=xt, flat=torch.from_numpy(flat).to(dev), off=torch.from_numpy(off).to(dev), sl=torch.from_numpy(sl).to(dev),
                     loss=torch.empty(B, device=dev), grad=torch.empty_like(xt), st=torch.empty(B, dtype=torch.int32, device=dev)))
need = ctypes.c_size_t(0)
lib.ocr_ctc_loss_workspace_bytes(T, B, C, 16, ctypes.byref(need))
ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=dev)
def step(i, sh):
    r = ring[i % len(ring)]
    _lib.check(lib.ocr_ctc_loss(_lib.ptr(r["x"]), T, B, C, _lib.ptr(r["flat"]), _lib.ptr(r["off"]), _lib.ptr(r["sl"]), 16, _lib.ptr(r["loss"]),
                                _lib.ptr(r["grad"]), _lib.ptr(r["st"]), 1.0 / B, _lib.ptr(ws), need.value, sh), "ctc")
ref = None
for path in (0, 2):
    for G in (0, 1, 2, 4, 8):
        lib.ocr_ctc_loss_set_path(path); lib.ocr_debug_ctc_group(G)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            sh = _lib.stream_handle()
            for i in range(5): step(i, sh)
            torch.cuda.synchronize()
            K = 200 if B <= 1024 else 10
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                gh = _lib.stream_handle()
                for i in range(K): step(i, gh)
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s); g.replay(); e1.record(s); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / K
        l = ring[0]["loss"].cpu().numpy().copy()
        if ref is None: ref = l
        print("path %d G %d: %.2f us/step  %.1f GB/s  max|dloss| %.2e" % (path, G, us, 2 * T * B * C * 4 / us / 1e3, np.abs(l - ref).max()))
lib.ocr_ctc_loss_set_path(0); lib.ocr_debug_ctc_group(0)
