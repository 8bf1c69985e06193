"""CPU tier: the N>1 plumbing of bench.py on the gloo backend, world size 2 (no GPU needed).
The CTC path shards by sequence with no data-path collective; what crosses ranks
is the max-over-ranks timing reduction and the reference arm's rank-0-only rule."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ms = bench.max_over_ranks(10.0 + 5.0 * rank, world, torch.device("cpu"))
    seeds = [bench.shard_seed(rank, i) for i in range(4)]
    gathered = [None] * world
    dist.all_gather_object(gathered, seeds)
    if rank == 0:
        flat = sum(gathered, [])
        json.dump({"ms": ms, "disjoint": len(set(flat)) == len(flat),
                   "value": bench.whole_job_value(256, 10, ms, world)}, open(out, "w"))
    dist.barrier()
    dist.destroy_process_group()


def test_max_over_ranks_and_sharding_gloo(tmp_path):
    out = str(tmp_path / "r.json")
    mp.spawn(_worker, args=(2, 29513, out), nprocs=2, join=True)
    r = json.load(open(out))
    assert r["ms"] == 15.0                      # the slowest rank's time
    assert r["disjoint"]                        # ranks own disjoint synthetic batches
    assert abs(r["value"] - 2 * 256 * 10 / 15e-3) < 1e-6


def test_reference_arm_rank0_only():
    """Under torchrun the reference arm runs on rank 0 alone; other ranks exit 0 without output."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env["RANK"] = "0"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
                        "--warmup", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["gpu_launches"] == 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0


def _ddp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import numpy as np
    from cnn_lstm_ctc_ocr_b200 import model, train
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shapes = {k: v.shape for k, v in model.init_params(0, "lstm", (32, 32), num_classes=19).items()}
    names, offsets, n_first, n_total = train.flat_layout(shapes, "lstm")
    # every replica holds the gradient of ITS half batch's mean loss; after the exchange all hold the global-batch mean
    g = torch.arange(n_total, dtype=torch.float32) * (rank + 1)
    scale = train.allreduce_buckets(g, n_first, world, lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM))
    g *= scale
    if rank == 0:
        json.dump({"ok": bool(torch.allclose(g, torch.arange(n_total, dtype=torch.float32) * 1.5)), "n_first": n_first, "n_total": n_total,
                   "first_is_rnn": all(n.startswith("rnn/") for n in names if offsets[n] < n_first),
                   "rest_is_conv": all(n.startswith("convnet/") for n in names if offsets[n] >= n_first),
                   "aligned": all(o % 64 == 0 for o in offsets.values())}, open(out, "w"))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_bucket_exchange_gloo(tmp_path):
    """The training step's one exchange: two contiguous buckets of the flat gradient buffer (logits + RNN, then
    the convolutional stack), SUM all-reduced and scaled by 1/world = the gradient of the global-batch mean loss."""
    out = str(tmp_path / "g.json")
    mp.spawn(_ddp_worker, args=(2, 29517, out), nprocs=2, join=True)
    r = json.load(open(out))
    assert r["ok"] and r["first_is_rnn"] and r["rest_is_conv"] and r["aligned"] and 0 < r["n_first"] < r["n_total"]
