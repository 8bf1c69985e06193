"""Times Trainer.train_step (cfg3-shaped) on one GPU: python tools/time_train.py [B] [W] [steps] [birnn path: 0 automatic, 3 persistent BPTT forced]."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from cnn_lstm_ctc_ocr_b200 import train, _lib
from cnn_lstm_ctc_ocr_b200 import model as _model
sys.path.insert(0, "tests")
from util import make_labels

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
W = int(sys.argv[2]) if len(sys.argv) > 2 else 256
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
if len(sys.argv) > 4:
    _lib.load().ocr_birnn_set_path(int(sys.argv[4]))
rng = np.random.default_rng(0)
params = _model.init_params(0, "lstm", (512, 512))
tr = train.Trainer(params)
img = torch.tensor(rng.integers(0, 256, (B, 32, W, 1)).astype(np.uint8), device="cuda")
widths = np.full(B, W)
T = (W - 2) // 2 - 2
labels = make_labels(rng, B, np.full(B, T), 24, 95)
for _ in range(2):
    loss = tr.train_step(img, widths, labels)
torch.cuda.synchronize()
print("warm loss", float(loss), "mem GB", torch.cuda.max_memory_allocated() / 2**30)
n0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time()
e0.record()
for _ in range(steps):
    loss = tr.train_step(img, widths, labels)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print("B=%d W=%d: %.2f ms/step (host %.2f ms), %.0f crops/s, launches/step %d, loss %.4f" % (B, W, ms, (time.time() - t0) * 1e3 / steps, B / ms * 1e3,
      (_lib.launch_count() - n0) // steps, float(loss)))
