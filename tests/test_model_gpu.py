"""GPU tier: the recognizer graph (conv stack -> BiLSTM / BiGRU -> logits -> greedy decode) on B200 against the
float64 numpy oracle (oracle/model_oracle.py).

Tolerance.  The contractions run as TF32 tensor-core products (each factor rounded to 10 mantissa bits, sums in
fp32): relative error per product <= 2^-10, accumulated over K <= 2304 terms of mixed sign and eight conv layers +
two recurrent layers it stays at the 1e-3 level relative to the activation scale.  Asserted: features within
5e-3 * max|features|, logits within 1e-2 * max|logits| of the float64 reference.  Decodes are compared on the
kernel's OWN logits (bit-exact against the CTC oracle); decode equality against the float64 graph is reported,
not required, because ReLU logits have near-ties that a 1e-3 perturbation may flip."""
import numpy as np
import pytest
import torch

from oracle import model_oracle as mo

pytestmark = pytest.mark.gpu


def _inputs(B, W, seed):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (B, 32, W, 1)).astype(np.uint8)
    widths = np.full(B, W, np.int32)
    return img, widths


@pytest.mark.parametrize("cell,sizes", [("lstm", (512, 512)), ("gru", (512, 256))])
def test_full_graph_vs_oracle(cell, sizes):
    from cnn_lstm_ctc_ocr_b200 import model
    B, W = 4, 128
    params = mo.init_params(seed=0, cell_type=cell, sizes=sizes, num_classes=95, dtype=np.float64, randomize_bn=True)
    img, widths = _inputs(B, W, 1)
    widths[1], widths[3] = 100, 41                     # ragged true widths inside the padded bucket (server.py:30)
    feats_ref, sl_ref = mo.convnet_layers(mo.preprocess_image(img), widths, params)
    logits_ref = mo.rnn_layers(feats_ref, sl_ref, params, cell, sizes)

    m = model.Model(params, cell_type=cell, rnn_sizes=sizes)
    dev = torch.device("cuda:0")
    feats, sl = m.convnet_layers(torch.tensor(img, device=dev), torch.tensor(widths), model.ModeKeys.INFER)
    assert sl.cpu().tolist() == sl_ref.tolist() == [61, 47, 61, 17]
    f = feats.cpu().numpy()
    assert f.shape == feats_ref.shape == (B, 61, 256)
    assert np.abs(f - feats_ref).max() <= 5e-3 * np.abs(feats_ref).max()
    # float input path (validate._preprocess_image applied by the caller): same features up to TF32 rounding noise
    # (last-bit differences of the preprocessed pixels land on different TF32 roundings downstream)
    feats2, _ = m.convnet_layers(model.preprocess_image(torch.tensor(img, device=dev)), torch.tensor(widths))
    assert np.abs(feats2.cpu().numpy() - f).max() <= 2e-3 * np.abs(f).max()

    logits = m.rnn_layers(feats, sl, 95)
    lg = logits.cpu().numpy()
    assert lg.shape == logits_ref.shape == (61, B, 96) and (lg >= 0).all()
    for b in range(B):   # frames past the sequence length are relu(bias) on both sides; compare the live part
        n = sl_ref[b]
        assert np.abs(lg[:n, b] - logits_ref[:n, b]).max() <= 1e-2 * np.abs(logits_ref).max()
    # decode of the kernel's own logits: bit-exact against the CTC oracle
    from oracle import ctc_oracle
    dense = m.get_output(logits, sl)[0].cpu().numpy()
    od, ol, _ = ctc_oracle.ctc_greedy_decoder(lg, sl_ref)
    assert (dense == ctc_oracle.densify(od, ol)).all()
    texts = m.recognize(torch.tensor(img, device=dev), torch.tensor(widths))
    assert len(texts) == B and all(isinstance(t, str) for t in texts)


@pytest.mark.parametrize("path", [0, 10, 1])
def test_rnn_layer_masks_and_directions(path):
    """Per-example lengths: zeros past the length, backward direction starts at len-1 (bidirectional_dynamic_rnn).
    path 0 = persistent tcgen05 LSTM kernel (binary16 recurrent operands, the default), 10 = the same kernel with TF32
    operands (ocr_debug_lstm_operands(0)), path 1 = frame-by-frame launches."""
    from cnn_lstm_ctc_ocr_b200 import model, _lib
    _lib.check(_lib.load().ocr_debug_lstm_operands(0 if path >= 10 else 1), "ocr_debug_lstm_operands")
    _lib.check(_lib.load().ocr_birnn_set_path(path % 10), "ocr_birnn_set_path")
    rng = np.random.default_rng(5)
    params = mo.init_params(seed=2, cell_type="lstm", sizes=(512, 512), dtype=np.float64)
    for k in list(params):
        if "lstm_cell/kernel" in k:
            params[k] = params[k] * 8      # make the recurrence matter (TruncNormal(0.01) is nearly linear)
    T, B = 12, 5
    feats = rng.standard_normal((B, T, 256))
    sl = np.array([12, 1, 7, 0, 12], np.int32)
    ref = mo.rnn_layer(np.transpose(feats, (1, 0, 2)), sl, params, "bdrnn1", "lstm", 512)
    m = model.Model(params, cell_type="lstm", rnn_sizes=(512, 512))
    dev = torch.device("cuda:0")
    out = m.rnn_layer(torch.tensor(feats, device=dev, dtype=torch.float32).transpose(0, 1).contiguous(), torch.tensor(sl, device=dev), 0)
    o = out.cpu().numpy()
    _lib.load().ocr_birnn_set_path(0)
    _lib.load().ocr_debug_lstm_operands(1)
    assert np.abs(o - ref).max() <= 5e-3 * np.abs(ref).max()
    for b in range(B):
        assert (o[sl[b]:, b] == 0).all()


@pytest.mark.parametrize("path", [0, 10, 1])
@pytest.mark.parametrize("layer,I,H", [(0, 256, 512), (1, 1024, 256)])
def test_gru_layer_masks_and_directions(path, layer, I, H):
    """The GRU model's two layers (model.py:167-199, 213-214; H = 512 and 256) against the numpy oracle: per-example lengths,
    zeros past the length, backward direction from len-1.  path 0 = persistent tcgen05 GRU kernel (two products per frame with
    a grid-wide exchange of r*h between them; binary16 recurrent operands, the default), 10 = the same kernel with TF32
    operands, path 1 = frame-by-frame launches."""
    from cnn_lstm_ctc_ocr_b200 import model, _lib
    _lib.check(_lib.load().ocr_debug_lstm_operands(0 if path >= 10 else 1), "ocr_debug_lstm_operands")
    _lib.check(_lib.load().ocr_birnn_set_path(path % 10), "ocr_birnn_set_path")
    rng = np.random.default_rng(6 + layer)
    params = mo.init_params(seed=3, cell_type="gru", sizes=(512, 256), dtype=np.float64)
    for k in list(params):
        if "gru_cell" in k and k.endswith("kernel"):
            params[k] = params[k] * 8      # make the recurrence matter (TruncNormal(0.01) is nearly linear)
    T, B = 13, 37
    feats = rng.standard_normal((B, T, I))
    sl = rng.integers(0, T + 1, B).astype(np.int32)
    sl[:4] = [13, 1, 7, 0]
    scope = "bdrnn%d" % (layer + 1)
    ref = mo.rnn_layer(np.transpose(feats, (1, 0, 2)), sl, params, scope, "gru", H)
    m = model.Model(params, cell_type="gru", rnn_sizes=(512, 256))
    dev = torch.device("cuda:0")
    out = m.rnn_layer(torch.tensor(feats, device=dev, dtype=torch.float32).transpose(0, 1).contiguous(), torch.tensor(sl, device=dev), layer)
    o = out.cpu().numpy()
    _lib.load().ocr_birnn_set_path(0)
    _lib.load().ocr_debug_lstm_operands(1)
    assert np.abs(o - ref).max() <= 5e-3 * np.abs(ref).max()
    for b in range(B):
        assert (o[sl[b]:, b] == 0).all()


def test_conv_paths_agree():
    """Implicit GEMM (default) and explicit im2col + GEMM are the same contraction in the same order of k."""
    from cnn_lstm_ctc_ocr_b200 import model
    params = mo.init_params(seed=4, dtype=np.float32, randomize_bn=True)
    img, widths = _inputs(3, 75, 2)      # odd width: ragged tiles on every layer
    dev = torch.device("cuda:0")
    f1, _ = model.Model(params, conv_path="igemm").convnet_layers(torch.tensor(img, device=dev), torch.tensor(widths))
    f2, _ = model.Model(params, conv_path="im2col").convnet_layers(torch.tensor(img, device=dev), torch.tensor(widths))
    assert f1.shape == f2.shape
    assert torch.equal(f1, f2)


def test_train_mode_is_the_trainers_job():
    from cnn_lstm_ctc_ocr_b200 import model
    m = model.Model(mo.init_params(0, dtype=np.float32))
    with pytest.raises(NotImplementedError):
        m.convnet_layers(torch.zeros((1, 32, 64, 1), device="cuda"), torch.tensor([64]), model.ModeKeys.TRAIN)


def test_get_testing_metrics():
    """test._get_testing (test.py:75-104): loss, label_error, sequence_error against the oracles."""
    from cnn_lstm_ctc_ocr_b200 import model
    from oracle import ctc_oracle
    from util import cfg2_inputs
    x, labels, seq_len = cfg2_inputs(seed=11, T=40, B=24, C=30, scale=3.0)
    dev = torch.device("cuda:0")
    idx = torch.tensor([[b, i] for b, l in enumerate(labels) for i in range(len(l))], dtype=torch.int64)
    vals = torch.tensor(sum(labels, []), dtype=torch.int32)
    label = (idx, vals, torch.tensor([24, 16]))
    lens = torch.tensor([len(l) for l in labels])
    loss, le, se = model.get_testing(torch.tensor(x, device=dev), torch.tensor(seq_len), label, lens)
    l64, _, _ = ctc_oracle.ctc_loss(x, labels, seq_len, f64=True, want_grad=False)
    od, ol, _ = ctc_oracle.ctc_beam_search_decoder(x, seq_len, 128, 1, True, nthreads=8)
    d = ctc_oracle.edit_distance([od[b, 0, :ol[b, 0]].tolist() for b in range(24)], labels)
    assert abs(float(loss) - l64.mean()) < 1e-4 * l64.mean()
    assert abs(float(le) - d.sum() / sum(len(l) for l in labels)) < 1e-6
    assert abs(float(se) - np.count_nonzero(d) / 24.0) < 1e-6


def test_recognize_graph_equals_eager():
    """Model.recognize replays a CUDA graph recorded per batch shape: same strings as the eager path, for device and
    host (pinned or pageable) uint8 batches, across shapes and repeated calls."""
    from cnn_lstm_ctc_ocr_b200 import model
    params = mo.init_params(seed=3, cell_type="lstm", sizes=(512, 512), num_classes=95, dtype=np.float32)
    m = model.Model(params)
    for (B, W) in ((4, 128), (3, 96), (4, 128)):
        img, widths = _inputs(B, W, 7 + W)
        widths[0] = W - 9
        eager = m.recognize(torch.tensor(img, device="cuda:0"), torch.tensor(widths), use_graph=False)
        assert m.recognize(torch.tensor(img, device="cuda:0"), torch.tensor(widths)) == eager
        assert m.recognize(torch.tensor(img), torch.tensor(widths)) == eager
        assert m.recognize(torch.tensor(img).pin_memory(), torch.tensor(widths)) == eager
    assert len(m._graphs) == 2


@pytest.mark.parametrize("B,H,W,C,Co,relu", [(2, 30, 37, 32, 32, 1), (3, 15, 21, 32, 64, 0), (2, 15, 40, 64, 64, 1), (1, 12, 8, 64, 32, 0),
                                              (2, 30, 254, 32, 32, 1)])
def test_conv_halo_kernel_equals_gather_kernel(B, H, W, C, Co, relu):
    """The halo-tile convolution (TMA-loaded input halo, nine taps as shifted descriptor views) and the gather kernel run the
    same MMAs in the same k order: identical bits, and both within TF32 tolerance of a float64 convolution."""
    import torch.nn.functional as F
    from cnn_lstm_ctc_ocr_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(B * 1000 + W)
    x = rng.standard_normal((B, H, W, C)).astype(np.float32)
    w = (rng.standard_normal((3, 3, C, Co)) * 0.1).astype(np.float32)
    bias = rng.standard_normal(Co).astype(np.float32)
    dx, dbias = torch.tensor(x, device="cuda:0"), torch.tensor(bias, device="cuda:0")
    dw = torch.tensor(np.ascontiguousarray(w.reshape(9 * C, Co).T), device="cuda:0")     # [Co, 9C] K-major
    outs = []
    for path in (1, 2):
        L.check(lib.ocr_conv_set_path(path), "path")
        out = torch.full((B, H, W, Co), 7.0, device="cuda:0")
        L.check(lib.ocr_conv3x3_same(L.ptr(dx), B, H, W, C, L.ptr(dw), L.ptr(dbias), Co, relu, L.ptr(out), L.stream_handle()), "conv")
        outs.append(out.cpu().numpy())
    L.check(lib.ocr_conv_set_path(0), "path")
    ref = F.conv2d(torch.tensor(x, dtype=torch.float64).permute(0, 3, 1, 2), torch.tensor(w, dtype=torch.float64).permute(3, 2, 0, 1),
                   torch.tensor(bias, dtype=torch.float64), padding=1).permute(0, 2, 3, 1).numpy()
    if relu:
        ref = np.maximum(ref, 0)
    assert np.abs(outs[1] - ref).max() <= 3e-3 * np.abs(ref).max()
    assert (outs[0] == outs[1]).all()


@pytest.mark.parametrize("B,H,W,C,Co,relu,stride_w", [(2, 30, 37, 32, 32, 1, 2), (2, 30, 254, 32, 32, 1, 2), (3, 15, 21, 64, 64, 1, 1), (2, 15, 126, 64, 64, 1, 1),
                                                       (1, 12, 8, 64, 32, 0, 2), (2, 13, 15, 32, 64, 0, 1), (1, 30, 1022, 32, 32, 1, 2)])
def test_conv_with_fused_pool_equals_conv_then_pool(B, H, W, C, Co, relu, stride_w):
    """ocr_conv3x3_same_pool (the pool taken in the halo-tile kernel's epilogue; pool2 = stride (2,2), pool4 = stride (2,1)) writes
    the bits of ocr_conv3x3_same followed by ocr_maxpool, touches nothing outside the pooled tensor, and the TMA-store and STG
    epilogues of both convolution kernels write the same bits."""
    from cnn_lstm_ctc_ocr_b200 import _lib as L
    lib, sh = L.load(), L.stream_handle()
    rng = np.random.default_rng(B * 1000 + W + stride_w)
    dx = torch.tensor(rng.standard_normal((B, H, W, C)).astype(np.float32), device="cuda:0")
    dw = torch.tensor(np.ascontiguousarray((rng.standard_normal((3, 3, C, Co)) * 0.1).astype(np.float32).reshape(9 * C, Co).T), device="cuda:0")
    dbias = torch.tensor(rng.standard_normal(Co).astype(np.float32), device="cuda:0")
    assert lib.ocr_conv3x3_pool_fused(B, H, W, C, Co, stride_w) == 1
    plain = {}
    for path in (1, 2):                      # gather kernel, halo-tile kernel
        for store in (1, 0):                 # TMA-store epilogue, STG epilogue
            L.check(lib.ocr_conv_set_path(path), "path")
            L.check(lib.ocr_debug_conv_tma_store(store), "store")
            out = torch.full((B, H, W, Co), float("nan"), device="cuda:0")
            L.check(lib.ocr_conv3x3_same(L.ptr(dx), B, H, W, C, L.ptr(dw), L.ptr(dbias), Co, relu, L.ptr(out), sh), "conv")
            plain[(path, store)] = out
    L.check(lib.ocr_conv_set_path(0), "path")
    L.check(lib.ocr_debug_conv_tma_store(1), "store")
    ref = plain[(1, 0)]
    assert torch.isfinite(ref).all()
    for k, v in plain.items():
        assert torch.equal(v, ref), k
    Hp, Wp = (H - 2) // 2 + 1, (W - 2) // stride_w + 1
    want = torch.empty((B, Hp, Wp, Co), device="cuda:0")
    L.check(lib.ocr_maxpool(L.ptr(ref), B, H, W, Co, 2, 2, 2, stride_w, L.ptr(want), sh), "pool")
    buf = torch.full((B * Hp * Wp * Co + 64,), float("nan"), device="cuda:0")
    got = buf[:B * Hp * Wp * Co].view(B, Hp, Wp, Co)
    L.check(lib.ocr_conv3x3_same_pool(L.ptr(dx), B, H, W, C, L.ptr(dw), L.ptr(dbias), Co, relu, stride_w, L.ptr(got), sh), "conv+pool")
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    assert torch.isnan(buf[B * Hp * Wp * Co:]).all()


@pytest.mark.parametrize("cell,sizes", [("lstm", (512, 512)), ("gru", (512, 256))])
def test_fused_pool_graph_equals_unfused_graph(cell, sizes):
    """Model(fuse_pool=True) (conv2 + pool2 and conv4 + pool4 as one launch each) == Model(fuse_pool=False), bit for bit."""
    from cnn_lstm_ctc_ocr_b200 import model, _lib
    params = model.init_params(0, cell, sizes)
    img, widths = _inputs(5, 150, 9)
    widths[1], widths[3] = 97, 64
    res = []
    for fuse in (True, False):
        m = model.Model(params, cell_type=cell, rnn_sizes=sizes, fuse_pool=fuse)
        n0 = _lib.launch_count()
        f, sl = m.convnet_layers(torch.tensor(img, device="cuda:0"), torch.tensor(widths))
        res.append((f.clone(), _lib.launch_count() - n0))
    assert torch.equal(res[0][0], res[1][0])
    assert res[0][1] == res[1][1] - 2        # two launches fewer
