#!/usr/bin/env python
"""bench.py -- measures the hot path on B200 (contract: see DESIGN.md "Measurement").

Headline workload = BASELINE.json configs[2], the recognizer's training step: conv stack + BiLSTM + CTC forward/backward +
Adam, global batch 256 of 32x256 crops, batch-sharded data parallel over the GPUs (per-GPU batch 256/N, NCCL gradient
all-reduce in two buckets) -- strong scaling.  A "step" is one training step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  value = line-crops/s of the whole job with the crops resident in HBM (CUDA-graph replay of
the step); e2e = the same through Trainer.train_step_captured from pinned host crops + labels to the host loss; roofline =
the metric's named kernel (BASELINE.json: "CTC fwd-bwd % HBM peak"): ctc_loss_fast_kernel in the bandwidth regime
(B=65536, 2.1 GB >> L2) with the configs[1] figure (B=256, one launch, latency-bound) beside it; cpu_baseline = the same
training step on the host cores (torch CPU float32 restatement, oracle/train_oracle.py, bounded sample); blocks = short
results of the other BASELINE configs (CTC configs[1], inference configs[0] LSTM and GRU, beam search configs[3], width
sweep configs[4]); the full per-block detail goes to stderr and to gpurun_out/bench_detail_n<N>.json.
`--impl reference` times the reference's CPU path for the same step: TensorFlow cannot run here, so it is
the restatement in oracle/train_oracle.py on all host threads -- kind "port".
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "line-crops/sec"
UNIT = "crops/s"
L2_BYTES = 126 * 1024 * 1024
_emit = print


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(workload)
    return None


# --------------------------------------------------------------------------- synthetic inputs
def make_ctc_batch(seed, T, B, C, max_label=16):
    import numpy as np
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((T, B, C), dtype=np.float32)
    seq_len = rng.integers(T // 2, T + 1, B).astype(np.int32)
    lens = np.minimum(rng.integers(1, max_label + 1, B), seq_len // 2).astype(np.int32)
    lens = np.maximum(lens, 1)
    off = np.zeros(B + 1, np.int32)
    off[1:] = np.cumsum(lens)
    flat = rng.integers(0, C - 1, int(off[-1])).astype(np.int32)
    rep = rng.random(flat.shape[0]) < 0.15
    first = np.zeros(flat.shape[0], bool)
    first[off[:-1]] = True
    idx = np.nonzero(rep & ~first)[0]
    flat[idx] = flat[idx - 1]
    return x, flat, off, seq_len, lens


# --------------------------------------------------------------------------- clocks sampler
class ClockSampler(threading.Thread):
    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.period = period
        self.samples = []  # (t, sm_mhz, reasons_mask)
        self.max_mhz = None
        self.ok = False
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.time(), mhz, rs))
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()

    def summary(self, windows):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        inside = [s for s in self.samples if any(a <= s[0] <= b for a, b in windows)]
        use = inside if inside else self.samples
        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
                 0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        mask = 0
        for s in use:
            mask |= s[2]
        reasons = [n for b, n in names.items() if mask & b and n != "gpu_idle"]
        return {"sm_mhz": statistics.median(s[1] for s in use), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples_under_load": len(inside), "samples": len(self.samples)}


# --------------------------------------------------------------------------- CPU baseline (oracle port)
def cpu_baseline_ctc(T, B, C, min_wall=10.0, min_reps=3):
    from oracle import ctc_oracle
    ctc_oracle.build()
    import numpy as np
    x, flat, off, seq_len, lens = make_ctc_batch(1, T, B, C)
    labels = [flat[off[b]:off[b + 1]].tolist() for b in range(B)]
    cores = ctc_oracle.max_threads()
    ctc_oracle.ctc_loss(x, labels, seq_len, nthreads=cores)  # warm
    reps, t0 = 0, time.perf_counter()
    while True:
        ctc_oracle.ctc_loss(x, labels, seq_len, nthreads=cores)
        reps += 1
        dt = time.perf_counter() - t0
        if reps >= min_reps and dt >= min_wall:
            break
    return {"value": B * reps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d passes over one batch (B=%d,T=%d,C=%d) of oracle/ctc_oracle.c CTC loss+grad, %d threads, %.2f s wall"
                      % (reps, B, T, C, cores, dt)}


REF_CROPS = 16   # crops per CPU step (bounded sample of the 256-crop step: the CPU cost is linear in the batch)


def cpu_train_step_timer(width, n_crops=REF_CROPS):
    """Returns (step_fn, cores): one call = one training step (forward TRAIN, backward, Adam) on n_crops crops of 32 x width
    through oracle/train_oracle.py in float32, torch CPU on all host threads."""
    import numpy as np
    import torch
    from oracle import model_oracle as mo
    from oracle import train_oracle as to
    params = mo.init_params(0, "lstm", (512, 512), 95, np.float32)
    img, widths, labels = make_train_batch(7, n_crops, width)

    def step():
        to.train_step_reference(params, img, widths, labels, step=0, cell_type="lstm", sizes=(512, 512), dtype=np.float32)
    return step, torch.get_num_threads()


def run_reference(args, cfg):
    """The reference's CPU implementation of the path, on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step, cores = cpu_train_step_timer(cfg["W"])
    for _ in range(max(min(args.warmup, 2), 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = REF_CROPS * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg["config"],
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "each step = one training step on %d of the 256 crops (32x%d, LSTM 512/512, 96 logits) through "
                                       "oracle/train_oracle.py: torch CPU float32 forward + autograd backward + Adam (TensorFlow itself "
                                       "cannot run in this image), %d threads" % (REF_CROPS, cfg["W"], cores)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(json.dumps(line))


# --------------------------------------------------------------------------- N > 1 plumbing (also exercised on CPU/gloo)
def shard_seed(rank, i):
    """Seed of the i-th synthetic batch of a rank: ranks own disjoint batches (weak scaling, no data-path collective)."""
    return 1000 * rank + i


def max_over_ranks(ms, world, device):
    """A multi-GPU time is the slowest rank's device time."""
    if world <= 1:
        return float(ms)
    import torch
    import torch.distributed as dist
    tm = torch.tensor([float(ms)], device=device)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    return float(tm.item())


def whole_job_value(units_per_rank_step, steps, ms, world):
    """Whole-job throughput: units all ranks processed / max-over-ranks time."""
    return world * units_per_rank_step * steps / (ms * 1e-3)


# --------------------------------------------------------------------------- our arm
def ctc_cfg2_block(args, dev, rank, world, windows, T=64, B=256, C=63, K=200):
    """configs[1]: CTC loss + gradient only, batch 256 per GPU (weak scaling, no collective).  value = K calls captured in one
    CUDA graph over a ring of distinct batches larger than L2; e2e = ctc.ctc_loss from pinned host logits/labels to host losses."""
    import ctypes
    import torch
    from cnn_lstm_ctc_ocr_b200 import _lib, ctc
    lib = _lib.load()
    bytes_per_batch = 2 * T * B * C * 4
    ring_n = max(2, -(-int(2.5 * L2_BYTES) // bytes_per_batch))
    host = [make_ctc_batch(shard_seed(rank, i), T, B, C) for i in range(min(ring_n, 8))]
    ring = []
    for i in range(ring_n):
        x, flat, off, seq_len, lens = host[i % len(host)]
        xt = torch.from_numpy(x).to(dev)
        if i >= len(host):
            xt = xt + 1e-3 * i  # distinct memory, same distribution
        ring.append(dict(x=xt, flat=torch.from_numpy(flat).to(dev), off=torch.from_numpy(off).to(dev),
                         sl=torch.from_numpy(seq_len).to(dev), Lmax=int(lens.max()),
                         loss=torch.empty(B, device=dev), grad=torch.empty_like(xt),
                         status=torch.empty(B, dtype=torch.int32, device=dev)))
    stream = torch.cuda.Stream(device=dev)
    need = ctypes.c_size_t(0)
    _lib.check(lib.ocr_ctc_loss_workspace_bytes(T, B, C, 16, ctypes.byref(need)), "ocr_ctc_loss_workspace_bytes")
    ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=dev)

    def step(i, st):
        r = ring[i % ring_n]
        _lib.check(lib.ocr_ctc_loss(_lib.ptr(r["x"]), T, B, C, _lib.ptr(r["flat"]), _lib.ptr(r["off"]), _lib.ptr(r["sl"]),
                                    r["Lmax"], _lib.ptr(r["loss"]), _lib.ptr(r["grad"]), _lib.ptr(r["status"]),
                                    1.0 / B, _lib.ptr(ws), need.value, st), "ocr_ctc_loss")
    with torch.cuda.stream(stream):
        sh = _lib.stream_handle()
        n0 = _lib.launch_count()
        step(0, sh)
        launches_per_step = _lib.launch_count() - n0
        for i in range(5):
            step(i, sh)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            gh = _lib.stream_handle()
            for i in range(K):
                step(5 + i, gh)
        graph.replay()  # warm the instantiated graph once (untimed)
        torch.cuda.synchronize()
        R = 20    # K x R launches: a timed region of tens of milliseconds
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_a = time.time()
        e0.record(stream)
        for _ in range(R):
            graph.replay()
        e1.record(stream)
        torch.cuda.synchronize()
        windows.append((t_a, time.time()))
        ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
        kernel_us = ms * 1e3 / (K * R)
        ok_status = int(ring[0]["status"].sum().item()) == 0
        # ---- end to end through the public API with host buffers
        hx = [torch.from_numpy(h[0]).pin_memory() for h in host]
        hlab = [(torch.from_numpy(h[1]).pin_memory(), torch.from_numpy(h[4]).pin_memory()) for h in host]
        hsl = [torch.from_numpy(h[3]).pin_memory() for h in host]
        hloss = [torch.empty(B, dtype=torch.float32).pin_memory() for _ in range(2)]
        dx = torch.empty((T, B, C), device=dev)
        dx.requires_grad_(True)

        def e2e_enqueue(i):
            j = i % len(host)
            with torch.no_grad():
                dx.copy_(hx[j], non_blocking=True)
            loss = ctc.ctc_loss(hlab[j], dx, hsl[j])  # labels/seq_len are host tensors: copied inside
            hloss[i % 2].copy_(loss.detach(), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            return ev

        def e2e_run(n):
            """n steps with one step in flight: step i+1 is queued before step i's losses are read on the host."""
            acc, prev = 0.0, None
            for i in range(n):
                ev = e2e_enqueue(i)
                if prev is not None:
                    prev.synchronize()
                    acc += float(hloss[(i - 1) % 2][0])
                prev = ev
            prev.synchronize()
            return acc + float(hloss[(n - 1) % 2][0])
        e2e_run(3)
        KE = 200
        torch.cuda.synchronize()
        t_a = time.time()
        e0.record(stream)
        e2e_run(KE)
        e1.record(stream)
        torch.cuda.synchronize()
        windows.append((t_a, time.time()))
        ms_e = max_over_ranks(e0.elapsed_time(e1), world, dev)
    h = host[0]
    peak, peak_src = _peaks()
    alg_bytes = 2 * T * B * C * 4
    ach = alg_bytes / (kernel_us * 1e-6) / 1e9
    out = {"workload": "BASELINE configs[1]: CTC loss + gradient only, batch 256 per GPU, T=64 frames, 63-class alphabet (blank=62), "
                       "seq_len U{32..64}, label length U{1..16}, fp32 logits ~N(0,1); weak scaling, no collective",
           "value": world * B / (kernel_us * 1e-6), "unit": UNIT, "us_per_call": kernel_us, "gpu_launches_per_step": launches_per_step,
           "timed_region": "%d replays of a CUDA graph of %d calls over a ring of %d distinct batches (%d MB > 126 MB L2)" % (R, K, ring_n, ring_n * bytes_per_batch >> 20),
           "e2e": {"value": world * B * KE / (ms_e * 1e-3), "unit": UNIT,
                   "h2d_bytes_per_step": int(h[0].nbytes + h[1].nbytes + h[2].nbytes + h[3].nbytes), "d2h_bytes_per_step": int(B * 4),
                   "api": "cnn_lstm_ctc_ocr_b200.ctc.ctc_loss (loss + gradient), pinned host logits/labels in, losses out; one step in flight"},
           "roofline": {"bound": "hbm", "kernel": "ctc_loss_fast_kernel", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                        "traffic": _traffic("ctc_cfg2"), "algorithmic_bytes_per_launch": alg_bytes, "kernel_us": kernel_us,
                        "note": "8.3 MB in 64 CTAs of 2x24 dependent lattice frames on 148 SMs: latency-bound by construction"},
           "parity_status_ok": ok_status}
    if rank == 0 and world == 1:
        out["cpu_baseline"] = cpu_baseline_ctc(T, B, C, min_wall=5.0)
    del ring
    torch.cuda.empty_cache()
    return out


def run_ours(args, cfg):
    import torch
    import torch.distributed as dist
    from cnn_lstm_ctc_ocr_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    K, W = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(local)
    sampler.start()
    windows = []

    # ---- headline: the training step (configs[2])
    if args.blocks_only:
        blocks = {}
        if not args.skip_ctc:
            blocks["ctc_cfg2"] = ctc_cfg2_block(args, dev, rank, world, windows)
        if not args.skip_infer:
            blocks["inference_gru"] = inference_block(dev, world, windows, cell="gru", with_cpu=False)
        if not args.skip_extra:
            blocks["beam_search"] = beam_block(dev, rank, world, windows, with_cpu=False)
            blocks["sweep"] = sweep_block(dev, rank, world, windows)
            blocks["sweep_b128"] = sweep_block(dev, rank, world, windows, bucket_size=128)
        sampler.stop()
        if rank == 0:
            _emit(json.dumps({"blocks_only": True, "n_gpus": world, "blocks": {k: _short(v) for k, v in blocks.items()}}))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    training = training_block(dev, rank, world, windows, global_batch=cfg["B"], W=cfg["W"], steps=K, warmup=W,
                              with_cpu=(rank == 0 and world == 1))
    # ---- the metric's named kernel: CTC loss + gradient, bandwidth regime (rank 0) and configs[1]
    bw = None
    if rank == 0 and not args.skip_bw:
        try:
            bw = bandwidth_regime(lib, _lib, dev, 64, 63, args.bw_batch, windows)
        except torch.cuda.OutOfMemoryError:
            bw = {"error": "out of memory"}
    if world > 1:
        dist.barrier()
    blocks = {}
    if not args.skip_ctc:
        blocks["ctc_cfg2"] = ctc_cfg2_block(args, dev, rank, world, windows)
    if not args.skip_infer:
        blocks["inference_lstm"] = inference_block(dev, world, windows, cell="lstm", with_cpu=(rank == 0 and world == 1))
        blocks["inference_gru"] = inference_block(dev, world, windows, cell="gru", with_cpu=False)
    if not args.skip_extra:
        blocks["beam_search"] = beam_block(dev, rank, world, windows, with_cpu=(rank == 0 and world == 1))
        blocks["sweep"] = sweep_block(dev, rank, world, windows)
        # "batch per bucket 32 (and a tuned size)": the recurrence costs a frame the same at 32 and at 128 rows
        blocks["sweep_b128"] = sweep_block(dev, rank, world, windows, bucket_size=128)
    sampler.stop()
    if rank == 0:
        cfg2 = blocks.get("ctc_cfg2")
        roof = dict(bw) if bw and "error" not in bw else {"bound": "hbm", "kernel": "ctc_loss_fast_kernel", "achieved": None, "peak": _peaks()[0],
                                                          "unit": "GB/s", "frac": None, "traffic": None}
        if cfg2:
            r2 = cfg2["roofline"]
            roof["cfg2"] = {"workload": "configs[1]: B=256, one launch of 64 CTAs (latency-bound)", "kernel_us": r2["kernel_us"],
                            "achieved": r2["achieved"], "frac": r2["frac"], "traffic": r2["traffic"]}
        line = {
            "metric": METRIC, "value": training["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": training["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "tf32 products (16-bit operands in the recurrent, projection and BPTT products), fp32 accumulate/storage/optimizer", "data": "synthetic",
            "config": dict(cfg["config"], per_gpu_batch=cfg["B"] // world, timed_region=training["timed_region"],
                           l2="per-step working set (activations + 43 MB of weights, gradients and Adam slots) >> 126 MB L2"),
            "e2e": training["e2e"], "gpu_launches": training["gpu_launches_per_step"] * K,
            "roofline": roof, "cpu_baseline": training.get("cpu_baseline"), "clocks": sampler.summary(windows),
            "train": {k: training[k] for k in ("loss", "loss_check", "allreduce_bytes_per_step", "exposed_allreduce_ms", "tensor_roofline") if k in training},
            "blocks": {k: _short(v) for k, v in blocks.items()},
        }
        detail = dict(line, blocks=blocks, training=training, roofline_bw_regime=bw)
        sys.stderr.write("BENCH_DETAIL " + json.dumps(detail) + "\n")
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "bench_detail_n%d.json" % world), "w") as f:
                json.dump(detail, f, indent=1)
        except OSError:
            pass
        _emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _short(b):
    """The few numbers of a block that go on the JSON line (full detail: stderr / gpurun_out/bench_detail_n<N>.json)."""
    out = {}
    for k in ("value", "ms_per_step", "us_per_call", "padded_pixel_overhead", "ms_total"):
        if k in b:
            out[k] = round(b[k], 4) if isinstance(b[k], float) else b[k]
    if isinstance(b.get("e2e"), dict):
        out["e2e"] = round(b["e2e"]["value"], 1)
    if isinstance(b.get("cpu_baseline"), dict):
        out["cpu"] = round(b["cpu_baseline"]["value"], 1)
        out["cpu_cores"] = b["cpu_baseline"]["cores"]
    if isinstance(b.get("roofline"), dict) and b["roofline"].get("frac") is not None:
        out["roofline_frac"] = round(b["roofline"]["frac"], 4)
        out["bound"] = b["roofline"].get("bound")
    return out


# --------------------------------------------------------------------------- recognizer inference (BASELINE configs[0])
INFER_FLOP_PER_CROP = {("lstm", 128): 1855.5e6, ("gru", 128): 1225.9e6}  # forward FLOPs, SURVEY.md section 8(d)


def inference_block(dev, world, windows, B=32, W=128, cell="lstm", steps=50, with_cpu=True):
    """configs[0]: weinman CNN-BiLSTM-CTC inference, batch 32 synthetic 32x128 grayscale crops, greedy decode,
    random-init weights.  value = graph replay with the crops resident in HBM; e2e = Model.recognize() from pinned
    host uint8 crops to host strings."""
    import numpy as np
    import torch
    from cnn_lstm_ctc_ocr_b200 import _lib, model
    sizes = (512, 512) if cell == "lstm" else (512, 256)
    params = model.init_params(0, cell, sizes)
    m = model.Model(params, cell_type=cell, rnn_sizes=sizes, device=dev)
    rng = np.random.default_rng(0)
    host_img = torch.from_numpy(rng.integers(0, 256, (B, 32, W, 1)).astype(np.uint8)).pin_memory()
    host_w = torch.full((B,), W, dtype=torch.int32).pin_memory()
    img = host_img.to(dev)
    widths = host_w.to(dev)

    def fwd():
        f, sl = m.convnet_layers(img, widths)
        lg = m.rnn_layers(f, sl)
        return m.get_output(lg, sl)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        for _ in range(3):
            fwd()
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        fwd()
        launches = _lib.launch_count() - n0
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            f, sl = m.convnet_layers(img, widths)
            lg = m.rnn_layers(f, sl)
            dec, ln, ns = __import__("cnn_lstm_ctc_ocr_b200").ctc.ctc_greedy_decode_raw(lg, sl)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_a = time.time()
        e0.record(stream)
        for _ in range(steps):
            g.replay()
        e1.record(stream)
        torch.cuda.synchronize()
        windows.append((t_a, time.time()))
        ms = max_over_ranks(e0.elapsed_time(e1), world, dev) / steps
        # end to end: host crops -> strings
        for _ in range(2):
            m.recognize(host_img, host_w)
        torch.cuda.synchronize()
        t_a = time.time()
        e0.record(stream)
        for _ in range(steps):
            texts = m.recognize(host_img, host_w)
        e1.record(stream)
        torch.cuda.synchronize()
        windows.append((t_a, time.time()))
        ms_e = max_over_ranks(e0.elapsed_time(e1), world, dev) / steps
    flop = INFER_FLOP_PER_CROP.get((cell, W))
    peaks = {}
    pth = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pth):
        with open(pth) as fjs:
            peaks = json.load(fjs)
    tf32_peak = float(peaks.get("bf16_tflops_sustained", 1400.0)) / 2.0
    out = {"workload": "BASELINE configs[0]: weinman CNN-Bi%s-CTC inference, batch %d synthetic 32x%d grayscale crops, greedy CTC decode, "
                       "random-init weights, 96 logits" % (cell.upper(), B, W),
           "value": world * B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "gpu_launches_per_step": launches,
           "timed_region": "cuda graph replay of the whole forward + greedy decode, crops resident in HBM",
           "dtype": "tf32 products, fp32 accumulate/storage",
           "e2e": {"value": world * B / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(host_img.numel() + host_w.numel() * 4),
                   "d2h_bytes_per_step": int(B * 61 * 8), "api": "cnn_lstm_ctc_ocr_b200.model.Model.recognize (uint8 crops -> strings)"},
           "roofline": {"bound": "tensor", "achieved": (flop * B / (ms * 1e-3) / 1e12) if flop else None, "peak": tf32_peak, "unit": "TFLOP/s",
                        "frac": (flop * B / (ms * 1e-3) / 1e12 / tf32_peak) if flop else None, "traffic": None,
                        "peak_source": "half of the measured sustained bf16 GEMM peak (TF32 runs at half the bf16 rate); "
                                       "the step is launch/latency bound by the %d-frame recurrence, not tensor bound" % 61}}
    if with_cpu:
        from oracle import model_oracle as mo   # the CPU baseline leg: the only use of oracle/ in this block
        t0 = time.perf_counter()
        p64 = {k: v.astype(np.float32) for k, v in params.items()}
        x = mo.preprocess_image(host_img.numpy()).astype(np.float32)
        feats, sl_ = mo.convnet_layers(x, np.full(B, W), p64)
        logits = mo.rnn_layers(feats, sl_, p64, cell, sizes)
        from oracle import ctc_oracle
        ctc_oracle.ctc_greedy_decoder(logits.astype(np.float32), sl_)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": B / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                               "sample": "one batch of %d crops through oracle/model_oracle.py (numpy float32, BLAS threads) + "
                                         "the C greedy decoder, %.1f s wall (TensorFlow itself cannot run in this image)" % (B, dt)}
    return out


# --------------------------------------------------------------------------- beam search (configs[3]) and width sweep (configs[4])
def beam_block(dev, rank, world, windows, batch=1024, T=64, C=63, beam=128, reps=5, with_cpu=True):
    """configs[3]: CTC beam-search decode (beam width 128, top path) over batch 1024 logits, sharded by crop across the
    GPUs (no collective).  Latency/dependency bound: T frames x up to 128 expansions per sequence, each of which changes
    the candidate set the next one is judged against (TensorFlow's order-dependent acceptance)."""
    import numpy as np
    import torch
    from cnn_lstm_ctc_ocr_b200 import ctc
    B = batch // world
    rng = np.random.default_rng(shard_seed(rank, 2))
    xh = (rng.standard_normal((T, B, C)) * 3).astype(np.float32)
    slh = rng.integers(T // 2, T + 1, B).astype(np.int32)
    x = torch.from_numpy(xh).to(dev)
    sl = torch.from_numpy(slh).to(dev)
    for _ in range(2):
        dec, ln, lp = ctc.ctc_beam_search_raw(x, sl, beam, 1, True, True)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_a = time.time()
    e0.record()
    for _ in range(reps):
        dec, ln, lp = ctc.ctc_beam_search_raw(x, sl, beam, 1, True, True)
    e1.record()
    torch.cuda.synchronize()
    windows.append((t_a, time.time()))
    ms = max_over_ranks(e0.elapsed_time(e1), world, dev) / reps
    frames = float(slh.sum())
    out = {"workload": "BASELINE configs[3]: CTC beam search, beam width %d, top path, merge_repeated, batch %d logits [T=%d, C=%d] ~3*N(0,1), "
                       "sharded by crop over %d GPU(s)" % (beam, batch, T, C, world),
           "value": batch / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "per_gpu_batch": B, "scaling": "strong",
           "mean_decoded_len": float(ln.float().mean().item()),
           "roofline": {"bound": "latency", "note": "dependency-chain bound, not HBM (logits are read once: %.1f MB) and not tensor: per sequence and frame the "
                                                    "beam entries are expanded in descending-score order, each acceptance changing the list the next is judged "
                                                    "against" % (T * B * C * 4 / 1e6),
                        "achieved": frames / (ms * 1e-3) / 1e6, "unit": "M sequence-frames/s per GPU", "peak": None, "frac": None, "traffic": None,
                        "us_per_sequence_frame_per_wave": ms * 1e3 / float(slh.max())}}
    if with_cpu:
        from oracle import ctc_oracle     # CPU baseline leg
        ctc_oracle.build()
        cores = ctc_oracle.max_threads()
        nb = min(B, 256)
        t0 = time.perf_counter()
        rd, rl, rp = ctc_oracle.ctc_beam_search_decoder(xh[:, :nb], slh[:nb], beam_width=beam, top_paths=1, merge_repeated=True, nthreads=cores)
        dt = time.perf_counter() - t0
        same = bool((dec[:nb].cpu().numpy() == rd).all() and (lp[:nb].cpu().numpy().view(np.int32) == rp.view(np.int32)).all())
        out["cpu_baseline"] = {"value": nb / dt, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": "%d of the %d sequences through oracle/ctc_oracle.c beam search (width %d), %d threads, %.2f s wall" % (nb, B, beam, cores, dt)}
        out["bit_exact_vs_oracle"] = same
    return out


def sweep_block(dev, rank, world, windows, n_crops=10000, bucket_size=32):
    """configs[4]: receipt-line recognition sweep, 10k synthetic variable-width crops (32x64 .. 32x1024), bucketed by
    width exactly as the reference's LocalServer does (32-px buckets, zero right-padding, fixed batch with filler crops),
    inference + greedy decode to strings; crops are dealt round-robin to the GPUs, no collective.  End to end: host
    uint8 crops in, host strings out."""
    import numpy as np
    import torch
    from cnn_lstm_ctc_ocr_b200 import model, server
    params = model.init_params(0, "lstm", (512, 512))
    m = model.Model(params, cell_type="lstm", rnn_sizes=(512, 512), device=dev)
    rng = np.random.default_rng(3)
    widths = rng.integers(64, 1025, n_crops)
    # every process is offered the same crops; LocalServer(shard=...) keeps the crops of its own BATCHES only
    # (bucket k, batch j -> process (k + j) % world), so the fillers are those of a single server
    pix = rng.integers(0, 256, (32, 1024), dtype=np.uint8)
    crops = [pix[:, :int(w)] for w in widths]
    # like the reference's server, which builds its graph once at start-up: one batch per bucket shape before the clock
    # starts (records the per-shape CUDA graphs in the model -- on every process, whichever batches it will be dealt)
    server.BatchLinePredictor(server.LocalServer(m, bucket_size=bucket_size, device=dev)).predict_batch(
        "warm", [np.zeros((32, w), np.uint8) for w in range(64, 1025, 32)])
    srv = server.LocalServer(m, bucket_size=bucket_size, device=dev, shard=(rank, world))
    pred = server.BatchLinePredictor(srv)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_a = time.time()
    e0.record()
    texts = pred.predict_batch("sweep", crops)
    e1.record()
    torch.cuda.synchronize()
    windows.append((t_a, time.time()))
    ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
    cnt = torch.tensor([float(srv.padded_pixels), float(srv.real_pixels), float(len(texts))], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(cnt)
    padded, real, nstr = (float(v) for v in cnt.tolist())
    return {"workload": "BASELINE configs[4]: %d synthetic crops, widths U{64..1024}, bucketed per server.py (32-px buckets, batch %d), "
                        "CNN-BiLSTM inference + greedy decode to strings, whole batches dealt to %d GPU(s)" % (n_crops, bucket_size, world),
            "value": n_crops / (ms * 1e-3), "unit": UNIT,
            "e2e": {"value": n_crops / (ms * 1e-3), "unit": UNIT, "api": "server.BatchLinePredictor.predict_batch (host uint8 crops -> strings)"},
            "ms_total": ms, "strings": int(nstr), "scaling": "strong",
            "padded_pixel_overhead": padded / max(real, 1.0) - 1.0}


# --------------------------------------------------------------------------- training step (BASELINE configs[2])
TRAIN_FLOP_PER_CROP = 3 * 3793.1e6   # W=256, LSTM 512/512: 3 x forward


def make_train_batch(seed, B, W, num_labels=95, max_label=24):
    import numpy as np
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (B, 32, W, 1)).astype(np.uint8)
    T = (W - 2) // 2 - 2
    labels = []
    for _ in range(B):
        n = int(rng.integers(1, max_label + 1))
        l = rng.integers(0, num_labels, n)
        for i in range(1, n):
            if rng.random() < 0.15:
                l[i] = l[i - 1]
        labels.append([int(v) for v in l])
    return img, np.full(B, W), labels


def training_block(dev, rank, world, windows, global_batch=256, W=256, steps=20, warmup=5, with_cpu=True):
    """configs[2]: full training step (conv + BiLSTM + CTC forward/backward + Adam), global batch 256 of 32x256 crops,
    data parallel over `world` GPUs (per-GPU batch 256/world) with the NCCL gradient all-reduce in two buckets, the
    first overlapped with the conv backward.  value: CUDA-graph replay with the crops resident in HBM;
    e2e: Trainer.train_step_captured from pinned host uint8 crops + labels to the host loss."""
    import numpy as np
    import torch
    from cnn_lstm_ctc_ocr_b200 import _lib, train
    from cnn_lstm_ctc_ocr_b200 import model as _model
    B = global_batch // world
    params = _model.init_params(0, "lstm", (512, 512))
    tr = train.Trainer(params, device=dev, process_group=(True if world > 1 else None))
    batches = [make_train_batch(shard_seed(rank, i), B, W) for i in range(3)]
    dimg = [torch.from_numpy(b[0]).to(dev) for b in batches]
    himg = [torch.from_numpy(b[0]).pin_memory() for b in batches]
    # parity guard inside the bench (rank 0): the first step's per-example losses against the float64 oracle on a 4-crop
    # subsample of the same batch.  Batch statistics couple the examples, so the oracle runs the same 4 crops as their own
    # batch and the product is asked for that batch too (eager path, before anything is captured).
    loss_check = None
    if rank == 0:
        from oracle import train_oracle as to     # checker only
        sub = (batches[0][0][:4], batches[0][1][:4], batches[0][2][:4])
        chk = train.Trainer(params, device=dev)
        got = chk.forward_backward(torch.from_numpy(sub[0]).to(dev), sub[1], sub[2]).cpu().numpy()
        ref = to.train_step_reference({k: v.astype(np.float64) for k, v in params.items()}, sub[0], sub[1], sub[2], step=0)["losses"]
        rel = float(np.max(np.abs(got - ref) / np.abs(ref)))
        loss_check = {"crops": 4, "max_rel_err_vs_float64_oracle": rel, "tolerance": 5e-3, "ok": bool(rel <= 5e-3)}
        del chk
    n0 = _lib.launch_count()
    tr.capture(B, W, max_label_len=24)
    launches = (_lib.launch_count() - n0) // 2     # warm-up pass + capture pass
    for i in range(max(warmup, 3)):
        tr.train_step_captured(dimg[i % 3], batches[i % 3][1], batches[i % 3][2])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(imgs, n, read_loss):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_a = time.time()
        e0.record()
        for i in range(n):
            losses = tr.train_step_captured(imgs[i % 3], batches[i % 3][1], batches[i % 3][2])
            if read_loss:
                hloss.copy_(losses.mean().reshape(1), non_blocking=True)
                torch.cuda.current_stream().synchronize()
        e1.record()
        barrier()
        windows.append((t_a, time.time()))
        return max_over_ranks(e0.elapsed_time(e1), world, dev) / n, losses
    hloss = torch.empty(1, dtype=torch.float32).pin_memory()
    ms, losses = timed(dimg, steps, False)
    loss_dev = float(losses.mean().item())
    # end to end: pinned host crops + labels in, host loss out, every step
    ms_e, _ = timed(himg, steps, True)
    exposed = None
    if world > 1:
        # the same steps with the two gradient all-reduces left out (timing only: the replicas drift apart afterwards, which is
        # why this comes last): the difference is the part of the collectives that the backward pass does not hide
        tr.skip_allreduce = True
        ms_no, _ = timed(dimg, steps, False)
        tr.skip_allreduce = False
        exposed = ms - ms_no
    peaks = {}
    pth = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pth):
        with open(pth) as fjs:
            peaks = json.load(fjs)
    tf32_peak = float(peaks.get("bf16_tflops_sustained", 1400.0)) / 2.0
    ach = TRAIN_FLOP_PER_CROP * global_batch / (ms * 1e-3) / 1e12 / world
    out = {"workload": "BASELINE configs[2]: full training step (conv+BiLSTM+CTC fwd/bwd + Adam), global batch %d of 32x%d crops, "
                       "LSTM 512/512, 96 logits, data-parallel over %d GPU(s)" % (global_batch, W, world),
           "value": global_batch / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "n_gpus": world, "per_gpu_batch": B, "scaling": "strong",
           "gpu_launches_per_step": launches, "steps": steps, "loss": loss_dev, "loss_check": loss_check,
           "timed_region": "K steps, each a CUDA-graph replay of the step (%s), crops resident in HBM, labels staged per step" %
                           ("one graph" if world == 1 else "three graphs, two NCCL bucket all-reduces between them, the first overlapped with the conv backward"),
           "e2e": {"value": global_batch / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(himg[0].numel() + B * 4 + B * 24 * 4 + (B + 1) * 4 + 4),
                   "d2h_bytes_per_step": 4, "api": "cnn_lstm_ctc_ocr_b200.train.Trainer.train_step_captured (pinned uint8 crops + labels -> loss)"},
           "allreduce_bytes_per_step": int(tr.n_floats * 4) if world > 1 else 0,
           "exposed_allreduce_ms": exposed,
           "tensor_roofline": {"bound": "tensor", "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s per GPU", "frac": ach / tf32_peak,
                               "peak_source": "half of the measured sustained bf16 GEMM peak (TF32 runs at half the bf16 rate)",
                               "algorithmic_flop_per_step": TRAIN_FLOP_PER_CROP * global_batch}}
    if with_cpu:
        step, cores = cpu_train_step_timer(W)
        step()
        reps, t0 = 0, time.perf_counter()
        while reps < 2 or time.perf_counter() - t0 < 10.0:
            step()
            reps += 1
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": REF_CROPS * reps / dt, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": "%d training steps on %d of the 256 crops through oracle/train_oracle.py (torch CPU float32 forward + "
                                         "autograd backward + Adam), %.1f s wall (TensorFlow itself cannot run in this image)" % (reps, REF_CROPS, dt)}
    del tr
    torch.cuda.empty_cache()
    return out


def bandwidth_regime(lib, _lib, dev, T, C, B, windows):
    import torch
    peak, peak_src = _peaks()
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    x = torch.randn((T, B, C), device=dev, generator=g)
    sl = torch.randint(T // 2, T + 1, (B,), device=dev, generator=g, dtype=torch.int32)
    lens = torch.minimum(torch.randint(1, 17, (B,), device=dev, generator=g, dtype=torch.int32), sl // 2).clamp_(min=1)
    off = torch.zeros(B + 1, dtype=torch.int32, device=dev)
    off[1:] = torch.cumsum(lens, 0)
    n = int(off[-1].item())
    flat = torch.randint(0, C - 1, (n,), device=dev, generator=g, dtype=torch.int32)
    loss = torch.empty(B, device=dev)
    grad = torch.empty_like(x)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    sh = _lib.stream_handle()
    import ctypes
    need = ctypes.c_size_t(0)
    _lib.check(lib.ocr_ctc_loss_workspace_bytes(T, B, C, 16, ctypes.byref(need)), "ocr_ctc_loss_workspace_bytes")
    ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=dev)

    def go():
        _lib.check(lib.ocr_ctc_loss(_lib.ptr(x), T, B, C, _lib.ptr(flat), _lib.ptr(off), _lib.ptr(sl), 16, _lib.ptr(loss),
                                    _lib.ptr(grad), _lib.ptr(status), 1.0 / B, _lib.ptr(ws), need.value, sh), "ocr_ctc_loss")
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_a = time.time()
    e0.record(torch.cuda.current_stream())
    for _ in range(reps):
        go()
    e1.record(torch.cuda.current_stream())
    torch.cuda.synchronize()
    windows.append((t_a, time.time()))
    us = e0.elapsed_time(e1) * 1e3 / reps
    alg = 2 * T * B * C * 4
    ach = alg / (us * 1e-6) / 1e9
    return {"bound": "hbm", "kernel": "ctc_loss_fast_kernel", "workload": "CTC loss+grad B=%d T=%d C=%d (%.2f GB moved, >> L2)" % (B, T, C, alg / 1e9),
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": _traffic("ctc_bw_regime"),
            "kernel_us": us, "crops_per_s": B / (us * 1e-6), "peak_source": peak_src,
            "redo_sequences": int((status == 100).sum().item())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-ctc", action="store_true", help="skip the CTC-only block (BASELINE configs[1])")
    ap.add_argument("--blocks-only", action="store_true", help="tuning: skip the headline training step, print the blocks")
    ap.add_argument("--skip-bw", action="store_true", help="skip the bandwidth-regime measurement")
    ap.add_argument("--bw-batch", type=int, default=65536)
    ap.add_argument("--skip-infer", action="store_true", help="skip the recognizer-inference block (BASELINE configs[0])")
    ap.add_argument("--skip-extra", action="store_true", help="skip the beam-search (configs[3]) and width-sweep (configs[4]) blocks")
    args = ap.parse_args()
    cfg = {"B": 256, "W": 256,
           "config": {"workload": "BASELINE configs[2]: full training step (conv+BiLSTM+CTC fwd/bwd + Adam), global batch 256 of synthetic "
                                  "32x256 uint8 crops, LSTM 512/512, 96 logits, labels U{1..24}, random-init weights",
                      "global_batch": 256, "crop": "32x256", "T": 125, "C": 96,
                      "parallelism": "dp%d: batch-sharded data parallel, NCCL gradient all-reduce (42.9 MB) in two buckets" % args.gpus}}
    # libraries (NCCL's version banner, for one) write to stdout: keep fd 1 clean for the ONE JSON line
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    global _emit

    def _emit(text):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(text)
        sys.stdout.flush()
        os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
