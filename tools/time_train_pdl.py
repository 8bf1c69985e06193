"""Captured training step (cfg3 shape) with the frame-by-frame BPTT chain launched with and without programmatic dependent launch:
python tools/time_train_pdl.py [B] [W] [steps]"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from cnn_lstm_ctc_ocr_b200 import train, _lib
from cnn_lstm_ctc_ocr_b200 import model as _model
sys.path.insert(0, "tests")
from util import make_labels
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
W = int(sys.argv[2]) if len(sys.argv) > 2 else 256
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
lib = _lib.load()
rng = np.random.default_rng(0)
params = _model.init_params(0, "lstm", (512, 512))
img = torch.tensor(rng.integers(0, 256, (B, 32, W, 1)).astype(np.uint8), device="cuda")
widths = np.full(B, W)
T = (W - 2) // 2 - 2
labels = make_labels(rng, B, np.full(B, T), 24, 95)
res = {}
for pdl in (1, 0, 1, 0):
    lib.ocr_debug_bptt_pdl(pdl)
    tr = train.Trainer(params)
    tr.capture(B, W, max_label_len=24)
    for _ in range(3):
        losses = tr.train_step_captured(img, widths, labels)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        losses = tr.train_step_captured(img, widths, labels)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print("pdl %d: %.3f ms/step, loss %.6f" % (pdl, ms, float(losses.mean())), flush=True)
    res.setdefault(pdl, []).append((ms, losses.clone(), tr.theta.clone()))
    del tr
    torch.cuda.empty_cache()
a, b = res[1][0], res[0][0]
print("same losses:", torch.equal(a[1], b[1]), " same parameters after the steps:", torch.equal(a[2], b[2]))
lib.ocr_debug_bptt_pdl(1)
