"""GPU tier: parity at BASELINE.json's OWN shapes (the code paths bench.py times), against the float64 oracles.

  cfg3 (training, LSTM 512/512, W=256 -> T=125): ocr_birnn_lstm_train_fwd / _bwd at H=512, I in {256, 1024},
       per-GPU batches 32 / 64 (persistent BPTT, N in two halves of 256) and 128 / 256 (two 128-row batch tiles forward,
       large-batch BPTT) -- model_bu.py:167-221, train.py:101-141;
       one whole Trainer step with rnn_sizes=(512,512) at W=256.
  cfg1 (inference, B=32, W=128) and cfg5's widest bucket (W=1024 -> T=509), LSTM(512,512) and GRU(512,256) -- model.py:126-221.

Batch rows of a recurrent layer are independent, so the float64 oracle runs on a SUBSET of rows (first/last row of every
128-row tile and of every 32-row TMEM lane group) with the same weights; outputs, gate-pre-activation gradients (through the
oracle's `probe`) and input gradients of those rows are compared with the full-batch GPU run.

Tolerances (TF32 products, fp32 sums, tanh.approx-based cell nonlinearities; see DESIGN.md "Tolerances"): layer outputs 5e-3,
gate gradients 2e-2, input gradients 2e-2 of the tensor's max magnitude."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _close(got, ref, rel, what=""):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    scale = max(np.abs(ref).max(), 1e-30)
    err = np.abs(got - ref).max()
    assert err <= rel * scale, "%s: max err %.3e vs scale %.3e (rel %.2e > %.1e)" % (what, err, scale, err / scale, rel)


def _t(a):
    return torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device=DEV)


@pytest.mark.parametrize("B,I", [(32, 256), (64, 1024), (128, 1024), (256, 256)])
def test_lstm_training_kernels_at_cfg3_shapes(B, I):
    from cnn_lstm_ctc_ocr_b200 import _lib as L
    from oracle import train_oracle as to
    lib, sh = L.load(), L.stream_handle()
    H, T = 512, 125
    rng = np.random.default_rng(100 + B)
    sl = rng.integers(T // 3, T + 1, B).astype(np.int32)
    sl[0], sl[-1] = T, 1
    x = rng.standard_normal((T, B, I)).astype(np.float32)
    # TruncNormal(0.01) recurrences barely recur: scale the kernels so that h_{t-1} matters over many frames
    ks = {d: (rng.standard_normal((I + H, 4 * H)) * (1.0 / np.sqrt(I + H))).astype(np.float32) for d in ("fw", "bw")}
    bs = {d: (rng.standard_normal(4 * H) * 0.1).astype(np.float32) for d in ("fw", "bw")}
    gout = rng.standard_normal((T, B, 2 * H)).astype(np.float32)
    rows = sorted(set([0, 1, 31, 32, B // 2 - 1, B // 2, B - 33, B - 32, B - 1] + ([127, 128] if B > 128 else [])))
    rows = [r for r in rows if 0 <= r < B]

    # ---- float64 oracle on the row subset
    tp = {}
    for d in ("fw", "bw"):
        tp["p/%s/lstm_cell/kernel" % d] = torch.tensor(ks[d].astype(np.float64))
        tp["p/%s/lstm_cell/bias" % d] = torch.tensor(bs[d].astype(np.float64))
    xt = torch.tensor(x[:, rows].astype(np.float64), requires_grad=True)
    probes = {d: torch.zeros((T, len(rows), 4 * H), dtype=torch.float64, requires_grad=True) for d in ("fw", "bw")}
    ref = torch.cat([to._run_direction(xt, sl[rows], tp, "p/%s/" % d, "lstm", H, rev, probe=probes[d]) for d, rev in (("fw", False), ("bw", True))], dim=2)
    ref.backward(torch.tensor(gout[:, rows].astype(np.float64)))

    # ---- the kernels on the full batch (automatic path: what Trainer / bench.py run at this batch size)
    wx = _t(np.concatenate([ks["fw"][:I].T, ks["bw"][:I].T], 0))
    wh = _t(np.concatenate([ks["fw"][I:].T, ks["bw"][I:].T], 0))
    bias = _t(np.concatenate([bs["fw"], bs["bw"]]))
    wh_rows = _t(np.concatenate([ks["fw"][I:], ks["bw"][I:]], 0))
    need = ctypes.c_size_t(0)
    L.check(lib.ocr_birnn_lstm_train_workspace_bytes(T, B, H, ctypes.byref(need)), "ws")
    ws = torch.empty(need.value, dtype=torch.uint8, device=DEV)
    out, gates, cs = torch.empty((T, B, 2 * H), device=DEV), torch.empty((T * B, 8 * H), device=DEV), torch.empty((T, B, 2 * H), device=DEV)
    dsl, dx_in, dgo = torch.tensor(sl, device=DEV), _t(x), _t(gout)
    L.check(lib.ocr_birnn_set_path(0), "path")
    L.check(lib.ocr_birnn_lstm_train_fwd(L.ptr(dx_in), T, B, I, H, L.ptr(dsl), L.ptr(wx), L.ptr(wh), L.ptr(bias), L.ptr(out), L.ptr(gates), L.ptr(cs),
                                         L.ptr(ws), need.value, sh), "fwd")
    o = out.cpu().numpy()
    _close(o[:, rows], ref.detach().numpy(), 5e-3, "lstm out B=%d I=%d" % (B, I))
    for b in range(B):                       # zero outputs past each example's length, all rows
        assert not o[sl[b]:, b].any()
    L.check(lib.ocr_birnn_lstm_bwd(L.ptr(dgo), T, B, H, L.ptr(dsl), L.ptr(gates), L.ptr(cs), L.ptr(wh_rows), L.ptr(ws), need.value, sh), "bwd")
    dG = gates.view(T, B, 8 * H).cpu().numpy()
    for d, dn in enumerate(("fw", "bw")):
        _close(dG[:, rows, d * 4 * H:(d + 1) * 4 * H], probes[dn].grad.numpy(), 2e-2, "d gates %s B=%d" % (dn, B))
    for b in range(B):
        assert not dG[sl[b]:, b].any()
    dGs = dG[:, rows].astype(np.float64)
    dx = dGs[..., :4 * H] @ ks["fw"][:I].T.astype(np.float64) + dGs[..., 4 * H:] @ ks["bw"][:I].T.astype(np.float64)
    _close(dx, xt.grad.numpy(), 2e-2, "dx")


@pytest.mark.parametrize("B,H,T", [(32, 512, 40), (23, 512, 17), (64, 512, 24), (40, 256, 19), (5, 64, 9)])
def test_bptt_row_copies_leave_every_bit_unchanged(B, H, T):
    """Persistent BPTT kernel: with B <= 32 (<= 64) rows the 128-row operand tile holds four (two) copies of every row and the
    copies share the partial-sum reads and the scatter (ocr_debug_bptt_copies).  The sums are added in the same fixed order
    either way: identical gate gradients, bit for bit."""
    from cnn_lstm_ctc_ocr_b200 import _lib as L
    lib, sh = L.load(), L.stream_handle()
    g = torch.Generator(device=DEV); g.manual_seed(B * 1000 + H)
    act = torch.rand((T * B, 8 * H), device=DEV, generator=g) * 0.8 + 0.1
    cs = torch.randn((T, B, 2 * H), device=DEV, generator=g) * 0.5
    dout = torch.randn((T, B, 2 * H), device=DEV, generator=g) * 0.01
    wh_rows = torch.randn((2 * H, 4 * H), device=DEV, generator=g) * 0.05
    sl = torch.randint(1, T + 1, (B,), dtype=torch.int32, device=DEV, generator=g)
    sl[0] = T
    need = ctypes.c_size_t(0)
    L.check(lib.ocr_birnn_lstm_train_workspace_bytes(T, B, H, ctypes.byref(need)), "ws")
    ws = torch.empty(need.value, dtype=torch.uint8, device=DEV)
    L.check(lib.ocr_birnn_set_path(3), "path")          # persistent BPTT whenever the shape fits
    res = []
    for copies in (1, 0):
        L.check(lib.ocr_debug_bptt_copies(copies), "copies")
        a = act.clone()
        L.check(lib.ocr_birnn_lstm_bwd(L.ptr(dout), T, B, H, L.ptr(sl), L.ptr(a), L.ptr(cs), L.ptr(wh_rows), L.ptr(ws), need.value, sh), "bwd")
        torch.cuda.synchronize()
        res.append(a)
    L.check(lib.ocr_debug_bptt_copies(1), "copies")
    L.check(lib.ocr_birnn_set_path(0), "path")
    assert torch.isfinite(res[0]).all()
    assert torch.equal(res[0], res[1])


def test_whole_training_step_lstm512_w256():
    """One step of train.py's graph at cfg3's layer sizes and crop width (LSTM 512/512, 32x256 crops, 96 logits), batch 8."""
    from cnn_lstm_ctc_ocr_b200 import train
    from oracle import model_oracle as mo
    from oracle import train_oracle as to
    B, W = 8, 256
    rng = np.random.default_rng(5)
    params = mo.init_params(3, "lstm", (512, 512), 95, np.float64, randomize_bn=True)
    img = rng.integers(0, 256, (B, 32, W, 1)).astype(np.uint8)
    widths = np.array([256, 250, 231, 256, 180, 256, 99, 256])
    labels = [[int(v) for v in rng.integers(0, 95, n)] for n in (24, 7, 12, 1, 9, 16, 3, 20)]
    ref = to.train_step_reference(params, img, widths, labels, step=0, cell_type="lstm", sizes=(512, 512))
    tr = train.Trainer(params, rnn_sizes=(512, 512))
    losses = tr.forward_backward(torch.tensor(img, device=DEV), widths, labels)
    _close(losses.cpu().numpy(), ref["losses"], 5e-3, "losses")
    _close(tr.last_logits.cpu().numpy(), ref["logits"], 1e-2, "logits")
    bad = {}
    for name, g in ref["grads"].items():
        got = tr.grads[name].cpu().numpy().astype(np.float64)
        if np.abs(g).max() < 1e-12:
            assert np.abs(got).max() < 1e-5, name
            continue
        if name.startswith("rnn/"):
            e = np.abs(got - g).max() / np.abs(g).max()
            if e > 3e-2:
                bad[name] = ("max", e)
        else:     # ReLU / max-pool masks make these discontinuous in the activations (test_train_oracle.py::test_conv_gradient_conditioning)
            l2 = np.linalg.norm(got - g) / np.linalg.norm(g)
            cos = float((got * g).sum() / (np.linalg.norm(got) * np.linalg.norm(g)))
            if l2 > 0.3 or cos < 0.95:
                bad[name] = ("l2", l2, "cos", cos)
    assert not bad, "gradients off: %s" % bad
    for name, v in tr.stats.items():
        _close(v.cpu().numpy(), ref["new_params"][name], 1e-3, name)


@pytest.mark.parametrize("cell,sizes,B,W", [("lstm", (512, 512), 32, 128), ("gru", (512, 256), 32, 128),
                                            ("lstm", (512, 512), 2, 1024), ("gru", (512, 256), 2, 1024)])
def test_inference_graph_at_cfg1_and_cfg5_shapes(cell, sizes, B, W):
    """cfg1 (B=32, W=128, T=61) and the widest cfg5 bucket (W=1024, T=509: halo-tile convolutions on long rows, the persistent
    recurrence over 509 frames) against oracle/model_oracle.py, ragged true widths inside the padded batch."""
    from cnn_lstm_ctc_ocr_b200 import model
    from oracle import ctc_oracle
    from oracle import model_oracle as mo
    params = mo.init_params(seed=1, cell_type=cell, sizes=sizes, num_classes=95, dtype=np.float64, randomize_bn=True)
    rng = np.random.default_rng(W + B)
    img = rng.integers(0, 256, (B, 32, W, 1)).astype(np.uint8)
    widths = rng.integers(W - 31, W + 1, B).astype(np.int32)        # the crops of one 32-px server bucket
    widths[0] = W
    feats_ref, sl_ref = mo.convnet_layers(mo.preprocess_image(img), widths, params)
    logits_ref = mo.rnn_layers(feats_ref, sl_ref, params, cell, sizes)
    T = (W - 2) // 2 - 2
    m = model.Model(params, cell_type=cell, rnn_sizes=sizes)
    feats, sl = m.convnet_layers(torch.tensor(img, device=DEV), torch.tensor(widths), model.ModeKeys.INFER)
    assert sl.cpu().tolist() == sl_ref.tolist() and feats.shape == (B, T, 256)
    _close(feats.cpu().numpy(), feats_ref, 5e-3, "features")
    logits = m.rnn_layers(feats, sl, 95)
    lg = logits.cpu().numpy()
    assert lg.shape == (T, B, 96)
    scale = np.abs(logits_ref).max()
    for b in range(B):
        n = sl_ref[b]
        assert np.abs(lg[:n, b] - logits_ref[:n, b]).max() <= 1e-2 * scale, "logits of crop %d" % b
    dense = m.get_output(logits, sl)[0].cpu().numpy()
    od, ol, _ = ctc_oracle.ctc_greedy_decoder(lg, sl_ref)
    assert (dense == ctc_oracle.densify(od, ol)).all()
    # the serving entry point gives the same strings as get_output + the reference's charset
    texts = m.recognize(torch.tensor(img), torch.tensor(widths))
    gold = open(__file__.rsplit("/", 1)[0] + "/golden/out_charset.txt", encoding="utf-8").read()
    assert texts == ["".join(gold[c] for c in row if c >= 0) for row in dense]


def test_checkpoint_round_trip_and_resume(tmp_path):
    """Trainer.save_npz -> Model.load_npz gives identical logits; Trainer.load_npz resumes with the Adam slots and
    global_step (train.py:185-201: the Supervisor's saver holds them), so a resumed run takes the same next step."""
    from cnn_lstm_ctc_ocr_b200 import model, train
    from oracle import model_oracle as mo
    rng = np.random.default_rng(0)
    params = mo.init_params(0, "lstm", (32, 32), 19, np.float64, randomize_bn=True)
    img = torch.tensor(rng.integers(0, 256, (4, 32, 48, 1)).astype(np.uint8), device=DEV)
    widths, labels = np.array([48, 40, 44, 48]), [[1, 2], [3], [4, 4, 5], [6]]
    a = train.Trainer(params, rnn_sizes=(32, 32), learning_rate=1e-3)
    for _ in range(3):
        a.train_step(img, widths, labels)
    path = str(tmp_path / "model.ckpt-3.npz")
    a.save_npz(path)
    with np.load(path) as z:
        assert int(z["global_step"]) == 3 and "rnn/logits/kernel/Adam" in z.files and "convnet/conv1/kernel/Adam_1" in z.files
    # inference model from the checkpoint: the same logits as the trainer's own variables
    m1, m2 = a.to_model(), model.Model.load_npz(path, cell_type="lstm", rnn_sizes=(32, 32))
    f1, s1 = m1.convnet_layers(img, torch.tensor(widths))
    f2, s2 = m2.convnet_layers(img, torch.tensor(widths))
    assert torch.equal(m1.rnn_layers(f1, s1), m2.rnn_layers(f2, s2))
    # resume: same step as the uninterrupted run, bit for bit
    b = train.Trainer.load_npz(path, rnn_sizes=(32, 32), learning_rate=1e-3)
    assert b.global_step == 3 and torch.equal(b.adam_m, a.adam_m) and torch.equal(b.adam_v, a.adam_v) and torch.equal(b.theta, a.theta)
    la, lb = a.train_step(img, widths, labels), b.train_step(img, widths, labels)
    assert torch.equal(la, lb) and torch.equal(a.theta, b.theta)
    # a resume WITHOUT the slots (what round 1 did) takes a different, ~3x larger first step
    c = train.Trainer.load_npz(path, rnn_sizes=(32, 32), learning_rate=1e-3)
    c.adam_m.zero_(); c.adam_v.zero_()
    c.train_step(img, widths, labels)
    assert not torch.equal(c.theta, a.theta)


def test_tune_from_and_tune_scope(tmp_path):
    """--tune_from restores every variable of the graph from a checkpoint; --tune_scope limits the optimiser to the
    variables whose name matches (train.py:105-111,152-165): the rest keep their values, moving averages still update."""
    from cnn_lstm_ctc_ocr_b200 import train
    from oracle import model_oracle as mo
    rng = np.random.default_rng(1)
    pre = mo.init_params(7, "lstm", (32, 32), 19, np.float64, randomize_bn=True)
    path = str(tmp_path / "pretrained.npz")
    np.savez(path, **pre)
    fresh = mo.init_params(8, "lstm", (32, 32), 19, np.float64)
    tr = train.Trainer(fresh, rnn_sizes=(32, 32), tune_from=path, tune_scope="rnn", learning_rate=1e-3)
    for k, v in tr.all_params().items():
        assert np.allclose(v.cpu().numpy(), pre[k], atol=1e-7), k
    assert all(n.startswith("rnn/") for n in tr.tuned) and len(tr.tuned) == 10
    before = {k: v.clone() for k, v in tr.all_params().items()}
    img = torch.tensor(rng.integers(0, 256, (4, 32, 48, 1)).astype(np.uint8), device=DEV)
    tr.train_step(img, np.array([48, 40, 44, 48]), [[1, 2], [3], [4, 4, 5], [6]])
    for k, v in tr.params.items():
        moved = not torch.equal(v, before[k])
        assert moved == k.startswith("rnn/"), k
    assert any(not torch.equal(v, before[k]) for k, v in tr.stats.items())           # UPDATE_OPS run regardless of the scope
    with pytest.raises(ValueError, match="No variables to optimize"):
        train.Trainer(fresh, rnn_sizes=(32, 32), tune_scope="nothing_matches")
    with pytest.raises(KeyError):
        train.Trainer(fresh, rnn_sizes=(32, 32), tune_from={k: v for k, v in pre.items() if "conv3" not in k})
