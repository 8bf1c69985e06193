#!/usr/bin/env python
"""Driver for ncu captures of the beam-search kernel: python tools/prof_beam.py [B] [reps]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from cnn_lstm_ctc_ocr_b200 import ctc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
T, C = 64, 63
rng = np.random.default_rng(2)
dev = torch.device("cuda:0")
x = torch.from_numpy((rng.standard_normal((T, B, C)) * 3).astype(np.float32)).to(dev)
sl = torch.from_numpy(rng.integers(T // 2, T + 1, B).astype(np.int32)).to(dev)
for _ in range(reps):
    ctc.ctc_beam_search_raw(x, sl, 128, 1, True, True)
torch.cuda.synchronize()
print("ok")
