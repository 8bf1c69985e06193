"""CPU tier: cross-checks oracle/model_oracle.py (the numpy restatement of the reference's TensorFlow graph)
against an independent implementation of the same mathematics (torch CPU ops) and hand-computable cases."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import model_oracle as mo


def test_preprocess_and_seq_len():
    x = np.array([0, 127, 255], np.uint8)
    np.testing.assert_allclose(mo.preprocess_image(x), [-0.5, 127 / 255 - 0.5, 0.5])
    p = mo.init_params(0, dtype=np.float64)
    feats, sl = mo.convnet_layers(np.zeros((3, 32, 64, 1)), [64, 33, 1000], p)
    assert feats.shape == (3, 29, 256)           # T = (W-2)//2 - 2   
    assert sl.tolist() == [29, 13, 497]


def test_convnet_vs_torch():
    rng = np.random.default_rng(1)
    p = mo.init_params(3, dtype=np.float64, randomize_bn=True)
    x = rng.uniform(-0.5, 0.5, (2, 32, 70, 1))
    feats, _ = mo.convnet_layers(x, [70, 70], p)
    t = torch.tensor(x).permute(0, 3, 1, 2)
    for (filters, k, padding, name, bn) in mo.LAYER_PARAMS:
        w = torch.tensor(p["convnet/%s/kernel" % name]).permute(3, 2, 0, 1)
        t = F.conv2d(t, w, torch.tensor(p["convnet/%s/bias" % name]), padding=0 if padding == "valid" else 1)
        if bn:
            q = "convnet/%s/batch_norm/" % name
            t = F.batch_norm(t, torch.tensor(p[q + "moving_mean"]), torch.tensor(p[q + "moving_variance"]),
                             torch.tensor(p[q + "gamma"]), torch.tensor(p[q + "beta"]), False, 0.0, mo.BN_EPS)
        t = F.relu(t)
        if name == "conv2":
            t = F.max_pool2d(t, 2, (2, 2))
        elif name in ("conv4", "conv6"):
            t = F.max_pool2d(t, 2, (2, 1))
        elif name == "conv8":
            t = F.max_pool2d(t, (3, 1), (3, 1))
    ref = t[:, :, 0, :].permute(0, 2, 1).numpy()
    assert ref.shape == feats.shape == (2, 32, 256)
    np.testing.assert_allclose(feats, ref, rtol=1e-9, atol=1e-10)


def test_bilstm_vs_torch():
    """TF LSTMCell (gate order i,j,f,o; forget_bias 1) == torch.nn.LSTM (i,f,g,o) after re-laying the weights;
    per-example lengths follow bidirectional_dynamic_rnn: zeros past the length, backward pass starts at len-1."""
    rng = np.random.default_rng(2)
    T, B, I, H = 9, 4, 6, 5
    seq = rng.standard_normal((T, B, I))
    seq_len = np.array([9, 4, 1, 7])
    params = {}
    lstm = torch.nn.LSTM(I, H, bidirectional=True).double()
    for d, sfx in (("fw", ""), ("bw", "_reverse")):
        k = rng.standard_normal((I + H, 4 * H)) * 0.3
        b = rng.standard_normal(4 * H) * 0.1
        params["rnn/l/%s/lstm_cell/kernel" % d] = k
        params["rnn/l/%s/lstm_cell/bias" % d] = b
        def relay(m):  # TF columns (i, j, f, o) -> torch rows (i, f, g, o)
            i_, j_, f_, o_ = np.split(m, 4, axis=-1)
            return np.concatenate([i_, f_, j_, o_], axis=-1)
        bb = b.copy()
        bb[2 * H:3 * H] += 1.0  # forget_bias
        with torch.no_grad():
            getattr(lstm, "weight_ih_l0" + sfx).copy_(torch.tensor(relay(k[:I]).T))
            getattr(lstm, "weight_hh_l0" + sfx).copy_(torch.tensor(relay(k[I:]).T))
            getattr(lstm, "bias_ih_l0" + sfx).copy_(torch.tensor(relay(bb)))
            getattr(lstm, "bias_hh_l0" + sfx).zero_()
    out = mo.rnn_layer(seq, seq_len, params, "l", "lstm", H)
    packed = torch.nn.utils.rnn.pack_padded_sequence(torch.tensor(seq), torch.tensor(seq_len), enforce_sorted=False)
    ref, _ = lstm(packed)
    ref, _ = torch.nn.utils.rnn.pad_packed_sequence(ref, total_length=T)
    np.testing.assert_allclose(out, ref.detach().numpy(), rtol=1e-9, atol=1e-10)
    for b in range(B):
        assert (out[seq_len[b]:, b] == 0).all()


def test_gru_cell_hand_case():
    """TF GRUCell applies the reset gate BEFORE the candidate matmul (unlike torch.nn.GRU)."""
    x = np.array([[1.0]]); h = np.array([[2.0]])
    gk = np.array([[0.5, -0.5], [0.25, 0.75]]); gb = np.array([0.1, -0.2])
    ck = np.array([[0.3], [-0.4]]); cb = np.array([0.05])
    r = 1 / (1 + np.exp(-(0.5 * 1 + 0.25 * 2 + 0.1)))
    u = 1 / (1 + np.exp(-(-0.5 * 1 + 0.75 * 2 - 0.2)))
    c = np.tanh(0.3 * 1 - 0.4 * (r * 2) + 0.05)
    np.testing.assert_allclose(mo.gru_cell(x, h, gk, gb, ck, cb), [[u * 2 + (1 - u) * c]], rtol=1e-12)


def test_full_graph_shapes_and_relu_logits():
    p = mo.init_params(0, cell_type="gru", sizes=(512, 256), dtype=np.float64)
    x = np.random.default_rng(0).uniform(-0.5, 0.5, (2, 32, 40, 1))
    feats, sl = mo.convnet_layers(x, [40, 36], p)
    logits = mo.rnn_layers(feats, sl, p, "gru", (512, 256))
    assert logits.shape == (17, 2, 96) and (logits >= 0).all()
    assert sl.tolist() == [17, 15]


def test_product_initialiser_matches_the_oracles():
    """cnn_lstm_ctc_ocr_b200.model.init_params (what bench.py and a training run start from) draws the same variables as the
    oracle's restatement of the reference's initialisers, for both cells, under TensorFlow's variable names."""
    from cnn_lstm_ctc_ocr_b200 import model
    for cell, sizes in (("lstm", (512, 512)), ("gru", (512, 256))):
        a = model.init_params(3, cell, sizes)
        b = mo.init_params(3, cell, sizes, 95, np.float32)
        assert sorted(a) == sorted(b)
        assert all(np.array_equal(a[k], b[k]) for k in a)
    assert model.init_params(0, num_classes=62)["rnn/logits/bias"].shape == (63,)
