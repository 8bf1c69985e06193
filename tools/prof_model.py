#!/usr/bin/env python
"""Driver for ncu captures of the recognizer forward: python tools/prof_model.py [B] [W] [cell] [reps]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cnn_lstm_ctc_ocr_b200 import model
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
W = int(sys.argv[2]) if len(sys.argv) > 2 else 128
cell = sys.argv[3] if len(sys.argv) > 3 else "lstm"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
sizes = (512, 512) if cell == "lstm" else (512, 256)
m = model.Model(model.init_params(0, cell, sizes), cell_type=cell, rnn_sizes=sizes)
dev = torch.device("cuda:0")
img = torch.randint(0, 256, (B, 32, W, 1), dtype=torch.uint8, device=dev)
widths = torch.full((B,), W, dtype=torch.int32, device=dev)
for _ in range(reps):
    f, sl = m.convnet_layers(img, widths)
    lg = m.rnn_layers(f, sl)
    out = m.get_output(lg, sl)
torch.cuda.synchronize()
print("ok", tuple(lg.shape), tuple(out[0].shape))
