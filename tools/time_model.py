#!/usr/bin/env python
"""Stage timing of the recognizer forward pass: python tools/time_model.py [B] [W] [cell]."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from cnn_lstm_ctc_ocr_b200 import model, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
W = int(sys.argv[2]) if len(sys.argv) > 2 else 128
cell = sys.argv[3] if len(sys.argv) > 3 else "lstm"
sizes = (512, 512) if cell == "lstm" else (512, 256)
m = model.Model(model.init_params(0, cell, sizes), cell_type=cell, rnn_sizes=sizes)
dev = torch.device("cuda:0")
img = torch.randint(0, 256, (B, 32, W, 1), dtype=torch.uint8, device=dev)
widths = torch.full((B,), W, dtype=torch.int32, device=dev)
def run():
    f, sl = m.convnet_layers(img, widths)
    lg = m.rnn_layers(f, sl)
    return m.get_output(lg, sl)
for _ in range(3): run()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
n0 = _lib.launch_count()
ev[0].record(); f, sl = m.convnet_layers(img, widths); ev[1].record()
seq = f.transpose(0, 1).contiguous(); r1 = m.rnn_layer(seq, sl.to(dev), 0); ev[2].record(); r2 = m.rnn_layer(r1, sl.to(dev), 1); ev[3].record()
lg = m.rnn_layers(f, sl); ev[4].record()
torch.cuda.synchronize()
print("B=%d W=%d %s: convnet %.3f ms, rnn1 %.3f ms, rnn2 %.3f ms (eager, %d launches so far)" % (B, W, cell, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), _lib.launch_count() - n0))
t0 = time.perf_counter(); reps = 10
for _ in range(reps): out = run()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
print("eager end-to-end %.3f ms per batch -> %.0f crops/s" % (dt * 1e3, B / dt))
# CUDA graph of the whole forward
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    run(); torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        f, sl = m.convnet_layers(img, widths); lg = m.rnn_layers(f, sl)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(reps): g.replay()
    e1.record(s); torch.cuda.synchronize()
print("graph replay %.3f ms per batch -> %.0f crops/s" % (e0.elapsed_time(e1) / reps, B * reps / (e0.elapsed_time(e1) * 1e-3)))
