#!/usr/bin/env python
"""A/B of two builds of the CTC fast kernel: `python tools/ab_ctc_variant.py [lib.so]` times ocr_ctc_loss in the bandwidth
regime (B=65536) and at cfg2 (B=256) with the given library (default: the in-tree one) and prints a checksum of the
results, so two processes (one per library) can be compared.  The variant library is built by
`python tools/ab_ctc_variant.py --build -DOCR_CTC_RESCALE_ALT=0 scratch/libocr_b200_var.so`."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "--build":
    from cnn_lstm_ctc_ocr_b200 import build as B
    define, out = sys.argv[2], os.path.join(ROOT, sys.argv[3])
    B.build()
    obj = out[:-3] + "_ctc_loss.o"
    subprocess.check_call([B._nvcc()] + B.ARCH + B.NVCC_FLAGS + [define, "-c", os.path.join(B.CSRC, "ctc_loss.cu"), "-o", obj],
                          stderr=subprocess.DEVNULL)
    objs = [os.path.join(B.OBJ, f) for f in sorted(os.listdir(B.OBJ)) if f.endswith(".o") and f != "ctc_loss.o"] + [obj]
    subprocess.check_call([B._nvcc()] + B.ARCH + ["-shared", "-o", out] + objs + ["-cudart", "static"])
    print("built", out)
    sys.exit(0)

import numpy as np  # noqa: E402
import torch  # noqa: E402
from cnn_lstm_ctc_ocr_b200 import _lib  # noqa: E402
import bench  # noqa: E402

if len(sys.argv) > 1:
    _lib.SO_PATH = os.path.join(ROOT, sys.argv[1])
lib = _lib.load()
dev = torch.device("cuda:0")
T, C = 64, 63
for rep in range(3):
    r = bench.bandwidth_regime(lib, _lib, dev, T, C, 65536, [])
    print("%s bw regime: %.1f us  frac %.4f  redo %d" % (os.path.basename(_lib.SO_PATH), r["kernel_us"], r["frac"], r["redo_sequences"]))
# cfg2 in a CUDA graph of 200 calls over a ring of 40 batches
B = 256
ring = []
for i in range(40):
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


s = torch.cuda.Stream()
with torch.cuda.stream(s):
    sh = _lib.stream_handle()
    for i in range(5):
        step(i, sh)
    torch.cuda.synchronize()
    K = 200
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        gh = _lib.stream_handle()
        for i in range(K):
            step(i, gh)
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        g.replay()
        e1.record(s)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / K)
l = torch.stack([r["loss"] for r in ring]).double().cpu().numpy()
gsum = float(sum(r["grad"].double().abs().sum().item() for r in ring))
print("cfg2: %.2f us/call  loss sum %.9f  |grad| sum %.9f  redo %d" % (best, l.sum(), gsum, sum(int((r["st"] == 100).sum()) for r in ring)))
