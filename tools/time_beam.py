#!/usr/bin/env python
"""Beam search (width 128) timing: CTA-per-sequence kernel (path 0) against the replay kernel (path 1): python tools/time_beam.py [B...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from cnn_lstm_ctc_ocr_b200 import _lib, ctc
lib = _lib.load()
dev = torch.device("cuda:0")
T, C = 64, 63
for B in [int(a) for a in sys.argv[1:]] or [1024, 128]:
    rng = np.random.default_rng(2)
    x = torch.from_numpy((rng.standard_normal((T, B, C)) * 3).astype(np.float32)).to(dev)
    sl = torch.from_numpy(rng.integers(T // 2, T + 1, B).astype(np.int32)).to(dev)
    outs = []
    for path in (0, 1):
        lib.ocr_debug_beam_path(path)
        for _ in range(2):
            o = ctc.ctc_beam_search_raw(x, sl, 128, 1, True, True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            o = ctc.ctc_beam_search_raw(x, sl, 128, 1, True, True)
        e1.record(); torch.cuda.synchronize()
        outs.append(o)
        print("B=%d path %d: %.3f ms per call" % (B, path, e0.elapsed_time(e1) / 5), flush=True)
    print("   identical:", all(torch.equal(a, b) for a, b in zip(outs[0], outs[1])))
lib.ocr_debug_beam_path(0)
import ctypes
buf = (ctypes.c_longlong * 16)()
lib.ocr_debug_beam_profile(None, 1)
ctc.ctc_beam_search_raw(x, sl, 128, 1, True, True); torch.cuda.synchronize()
lib.ocr_debug_beam_profile(buf, 1)
names = ['scores+sort', 'advance+exclusions+rank', 'selection', 'resets', 'winners+ranking', 'rebuild']
print('CTA 0: frames %d, frames with a second event pass %d, frames with resets %d, events per frame %.1f (first pass) %.1f (second pass, when run)' % (int(sl[0]), buf[8], buf[9], buf[10] / max(int(sl[0]), 1), buf[11] / max(buf[8], 1)))
print('   inside resets: a_in scans %.0f, event passes %.0f, settling (per occurrence) %.0f' % (buf[12] / max(int(sl[0]), 1), buf[13] / max(int(sl[0]), 1), buf[14] / max(buf[8], 1)))
print('CTA 0 cycles per frame:', ', '.join('%s %.0f' % (n, buf[i] / max(int(sl[0]), 1)) for i, n in enumerate(names)))
