"""Captured training step (CUDA-graph replay, crops resident in HBM) at several per-GPU batch sizes on one GPU:
python tools/time_train_captured.py [B[:birnn_path] ...]   (default 256 32; birnn_path 3 forces the persistent BPTT kernel).  The numbers
bench.py's training block reports as `value`."""
import sys
sys.path.insert(0, ".")
import torch
import bench
from cnn_lstm_ctc_ocr_b200 import train
from cnn_lstm_ctc_ocr_b200 import model as _model

dev = torch.device("cuda:0")
from cnn_lstm_ctc_ocr_b200 import _lib
for arg in (sys.argv[1:] or ["256", "32"]):
    B, path = (int(v) for v in (arg.split(":") + ["0"])[:2])
    _lib.load().ocr_birnn_set_path(path)
    params = _model.init_params(0, "lstm", (512, 512))
    tr = train.Trainer(params, device=dev)
    batches = [bench.make_train_batch(i, B, 256) for i in range(3)]
    dimg = [torch.from_numpy(b[0]).to(dev) for b in batches]
    tr.capture(B, 256, max_label_len=24)
    for i in range(5):
        tr.train_step_captured(dimg[i % 3], batches[i % 3][1], batches[i % 3][2])
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            losses = tr.train_step_captured(dimg[i % 3], batches[i % 3][1], batches[i % 3][2])
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    print("B=%d path %d: %.3f ms/step  loss %.5f" % (B, path, best, float(losses.mean())))
    del tr
    torch.cuda.empty_cache()
