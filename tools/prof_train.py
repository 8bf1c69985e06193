"""Driver for ncu launch lists of the training step: python tools/prof_train.py [B] [W] [steps]."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from cnn_lstm_ctc_ocr_b200 import train
from cnn_lstm_ctc_ocr_b200 import model as _model
sys.path.insert(0, "tests")
from util import make_labels
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
W = int(sys.argv[2]) if len(sys.argv) > 2 else 256
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rng = np.random.default_rng(0)
tr = train.Trainer(_model.init_params(0, "lstm", (512, 512)))
img = torch.tensor(rng.integers(0, 256, (B, 32, W, 1)).astype(np.uint8), device="cuda")
labels = make_labels(rng, B, np.full(B, (W - 2) // 2 - 2), 24, 95)
for _ in range(steps):
    loss = tr.train_step(img, np.full(B, W), labels)
torch.cuda.synchronize()
print("ok", float(loss))
