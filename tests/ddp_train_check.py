"""Data-parallel training check, launched by tests/test_train_gpu.py::test_data_parallel_step under torchrun on 2 GPUs:
two replicas with half the batch each (sync_bn=True, NCCL gradient all-reduce in two buckets) must reproduce the
single-replica step on the whole batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from cnn_lstm_ctc_ocr_b200 import train
    from oracle import model_oracle as mo
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(0)
    B, W = 8, 48
    params = mo.init_params(0, "lstm", (32, 32), 19, np.float64, randomize_bn=True)
    for k in list(params):
        if "lstm_cell/kernel" in k:
            params[k] = params[k] * 8
    img = rng.integers(0, 256, (B, 32, W, 1)).astype(np.uint8)
    widths = np.array([W - (3 * i) % 11 for i in range(B)])
    labels = [[int(v) for v in rng.integers(0, 19, rng.integers(1, 5))] for _ in range(B)]
    # single replica on the whole batch (every rank computes it for itself)
    ref = train.Trainer(params, rnn_sizes=(32, 32), device=dev)
    ref.forward_backward(torch.tensor(img, device=dev), widths, labels)
    ref.apply_gradients()
    # two replicas, half the batch each
    n = B // world
    sl = slice(rank * n, (rank + 1) * n)
    for mode in ("eager", "captured"):
        tr = train.Trainer(params, rnn_sizes=(32, 32), device=dev, process_group=True, sync_bn=(mode == "eager"))
        if mode == "eager":
            tr.forward_backward(torch.tensor(img[sl], device=dev), widths[sl], labels[sl])
            tr.apply_gradients()
            # identical statistics and gradients up to summation order: variables agree to float32 noise, except where
            # Adam's first step amplifies the sign of a near-zero gradient (|update| = lr either way)
            d = (tr.theta - ref.theta).abs()
            frac_off = float((d > 2e-5).float().mean())
            assert frac_off < 0.02, "eager data-parallel step differs from the single-replica step: %.4f of the variables" % frac_off
            for k in ref.stats:
                assert torch.allclose(tr.stats[k], ref.stats[k], rtol=1e-4, atol=1e-6), k
        else:
            tr.capture(n, W, max_label_len=8)
            losses = tr.train_step_captured(torch.tensor(img[sl], device=dev), widths[sl], labels[sl])
            assert bool(torch.isfinite(losses).all())
            # replicas stay in lock-step: all ranks hold the same variables after the all-reduced update
            mine = tr.theta.clone()
            other = tr.theta.clone()
            dist.broadcast(other, src=0)
            assert torch.equal(mine, other), "replicas diverged"
    dist.barrier()
    if rank == 0:
        print("DDP_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
