"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, float64 or float32) of the reference's recognizer graph.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  The product package (cnn_lstm_ctc_ocr_b200) never does.

Restates, layer for layer, what the reference builds with TensorFlow 1.x (the arithmetic lives in the
un-vendored, un-pinned TensorFlow dependency, README.md:62; semantics per SURVEY.md Appendix A.1-A.3):

  preprocess_image   <- validate._preprocess_image      /root/reference/src/weinman/validate.py:56-68
  conv_layer         <- model.conv_layer                src/weinman/model.py:84-109   (tf.layers.conv2d, NHWC/HWIO)
  norm_layer         <- model.norm_layer                src/weinman/model.py:118-123  (tf.layers.batch_normalization, INFER)
  pool_layer         <- model.pool_layer / pool8        src/weinman/model.py:111-116, 145-146
  convnet_layers     <- model.convnet_layers            src/weinman/model.py:126-165
  rnn_layer (LSTM)   <- model_bu.rnn_layer              src/weinman/model_bu.py:167-199 (LSTMCell, bidirectional_dynamic_rnn)
  rnn_layer (GRU)    <- model.rnn_layer                 src/weinman/model.py:167-199    (GRUCell)
  rnn_layers         <- model.rnn_layers                src/weinman/model.py:202-221

PARITY PIN: the reference holds no golden vectors for this path and TensorFlow cannot run here, so parity
with TensorFlow itself is "unpinned".  The restatement is cross-checked in tests/test_model_oracle.py against
an independent implementation of the same mathematics (torch.nn.functional.conv2d / max_pool2d / batch_norm,
torch.nn.LSTM with re-laid-out weights) and on hand-computable cases.
"""
import numpy as np

# (filters, kernel, padding, name, batch_norm)  -- model.py:47-54 ("model_version4")
LAYER_PARAMS = [(32, 3, "valid", "conv1", False), (32, 3, "same", "conv2", True),
                (64, 3, "same", "conv3", False), (64, 3, "same", "conv4", True),
                (128, 3, "same", "conv5", False), (128, 3, "same", "conv6", True),
                (256, 3, "same", "conv7", False), (256, 3, "same", "conv8", True)]
BN_EPS = 1e-3  # tf.layers.batch_normalization default epsilon


def preprocess_image(image_u8):
    """uint8 [..] -> float: convert_image_dtype (x/255) then subtract 0.5 (validate.py:61-62)."""
    return image_u8.astype(np.float64) / 255.0 - 0.5


def conv2d(x, kernel, bias, padding):
    """tf.layers.conv2d, stride 1, NHWC input [B,H,W,Cin], HWIO kernel [3,3,Cin,Cout], cross-correlation."""
    B, H, W, Cin = x.shape
    kh, kw, _, Cout = kernel.shape
    if padding == "same":
        x = np.pad(x, ((0, 0), (kh // 2, kh // 2), (kw // 2, kw // 2), (0, 0)))
    Ho, Wo = x.shape[1] - kh + 1, x.shape[2] - kw + 1
    out = np.zeros((B, Ho, Wo, Cout), x.dtype)
    for i in range(kh):
        for j in range(kw):
            out += x[:, i:i + Ho, j:j + Wo, :] @ kernel[i, j]
    return out + bias


def batch_norm_infer(x, gamma, beta, mean, var):
    return gamma * (x - mean) / np.sqrt(var + BN_EPS) + beta


def max_pool(x, window, strides):
    """tf.layers.max_pooling2d, padding 'valid': out = floor((in - win)/stride) + 1."""
    B, H, W, C = x.shape
    ph, pw = window
    sh, sw = strides
    Ho, Wo = (H - ph) // sh + 1, (W - pw) // sw + 1
    out = np.full((B, Ho, Wo, C), -np.inf, x.dtype)
    for i in range(ph):
        for j in range(pw):
            out = np.maximum(out, x[:, i:i + sh * Ho:sh, j:j + sw * Wo:sw, :][:, :Ho, :Wo, :])
    return out


def convnet_layers(inputs, widths, params):
    """inputs [B,32,W,1] float (already preprocessed), widths [B] -> (features [B,T,256], sequence_length [B]).
    INFER mode (moving statistics).  params: dict keyed by TF variable names."""
    x = inputs
    feats = {}
    for (filters, k, padding, name, bn) in LAYER_PARAMS:
        x = conv2d(x, params["convnet/%s/kernel" % name], params["convnet/%s/bias" % name], padding)
        if bn:
            p = "convnet/%s/batch_norm/" % name
            x = batch_norm_infer(x, params[p + "gamma"], params[p + "beta"], params[p + "moving_mean"], params[p + "moving_variance"])
        x = np.maximum(x, 0)  # ReLU: the conv's own activation (no BN) or the explicit relu after BN (model.py:105-107)
        feats[name] = x
        if name == "conv2":
            x = max_pool(x, (2, 2), (2, 2))
        elif name in ("conv4", "conv6"):
            x = max_pool(x, (2, 2), (2, 1))
        elif name == "conv8":
            x = max_pool(x, (3, 1), (3, 1))
    features = x[:, 0]  # squeeze the row dimension (model.py:147)
    widths = np.asarray(widths, np.int64)
    seq_len = ((widths - 2) // 2 - 1 - 1).astype(np.int32)  # model.py:152-163
    return features, seq_len


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def lstm_cell(x, h, c, kernel, bias, forget_bias=1.0):
    """tf.contrib.rnn.LSTMCell (no peepholes / projection): [x,h] @ kernel + bias, split i, j, f, o."""
    z = np.concatenate([x, h], axis=1) @ kernel + bias
    H = h.shape[1]
    i, j, f, o = z[:, :H], z[:, H:2 * H], z[:, 2 * H:3 * H], z[:, 3 * H:]
    c_new = _sigmoid(f + forget_bias) * c + _sigmoid(i) * np.tanh(j)
    h_new = _sigmoid(o) * np.tanh(c_new)
    return h_new, c_new


def gru_cell(x, h, gate_kernel, gate_bias, cand_kernel, cand_bias):
    """tf.contrib.rnn.GRUCell: r,u = sigmoid([x,h] Wg + bg); c = tanh([x, r*h] Wc + bc); h' = u*h + (1-u)*c."""
    H = h.shape[1]
    g = _sigmoid(np.concatenate([x, h], axis=1) @ gate_kernel + gate_bias)
    r, u = g[:, :H], g[:, H:]
    cand = np.tanh(np.concatenate([x, r * h], axis=1) @ cand_kernel + cand_bias)
    return u * h + (1.0 - u) * cand


def _run_direction(seq, seq_len, cell, H, reverse):
    """dynamic_rnn over time-major seq [T,B,I] with per-example lengths: outputs are zero past the length and the
    state is carried through; the backward direction runs on reverse_sequence(seq) and is reversed back."""
    T, B, _ = seq.shape
    out = np.zeros((T, B, H), seq.dtype)
    h = np.zeros((B, H), seq.dtype)
    c = np.zeros((B, H), seq.dtype)
    idx = np.arange(B)
    for s in range(T):
        t = (seq_len - 1 - s) if reverse else np.full(B, s)
        live = s < seq_len
        tt = np.where(live, t, 0)
        h_new, c_new = cell(seq[tt, idx], h, c)
        h = np.where(live[:, None], h_new, h)
        c = np.where(live[:, None], c_new, c)
        out[tt[live], idx[live]] = h[live]
    return out


def rnn_layer(seq, seq_len, params, scope, cell_type, H):
    """Bidirectional layer, outputs [T,B,2H] = concat(fw, bw) (model.py:187-197)."""
    outs = []
    for d, reverse in (("fw", False), ("bw", True)):
        p = "rnn/%s/%s/" % (scope, d)
        if cell_type == "lstm":
            k, b = params[p + "lstm_cell/kernel"], params[p + "lstm_cell/bias"]
            cell = lambda x, h, c, k=k, b=b: lstm_cell(x, h, c, k, b)
        else:
            gk, gb = params[p + "gru_cell/gates/kernel"], params[p + "gru_cell/gates/bias"]
            ck, cb = params[p + "gru_cell/candidate/kernel"], params[p + "gru_cell/candidate/bias"]
            cell = lambda x, h, c, gk=gk, gb=gb, ck=ck, cb=cb: (gru_cell(x, h, gk, gb, ck, cb), c)
        outs.append(_run_direction(seq, np.asarray(seq_len), cell, H, reverse))
    return np.concatenate(outs, axis=2)


def rnn_layers(features, seq_len, params, cell_type="lstm", sizes=(512, 512)):
    """features [B,T,256] -> logits [T,B,num_classes+1] (dense + ReLU, model.py:216-220)."""
    seq = np.transpose(features, (1, 0, 2))
    r1 = rnn_layer(seq, seq_len, params, "bdrnn1", cell_type, sizes[0])
    r2 = rnn_layer(r1, seq_len, params, "bdrnn2", cell_type, sizes[1])
    return np.maximum(r2 @ params["rnn/logits/kernel"] + params["rnn/logits/bias"], 0)


def truncated_normal(rng, shape, std):
    """tf truncated normal: resample beyond two standard deviations."""
    x = rng.standard_normal(shape)
    bad = np.abs(x) > 2
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2
    return x * std


def init_params(seed=0, cell_type="lstm", sizes=(512, 512), num_classes=95, dtype=np.float32, randomize_bn=False):
    """Random-init parameters with the reference's initialisers (model.py:94-95,170,207-208; App. A.1-A.3)."""
    rng = np.random.default_rng(seed)
    p = {}
    cin = 1
    for (filters, k, padding, name, bn) in LAYER_PARAMS:
        std = np.sqrt(1.3 * 2.0 / (k * k * cin))  # variance_scaling_initializer(factor=2, FAN_IN, truncated normal)
        p["convnet/%s/kernel" % name] = truncated_normal(rng, (k, k, cin, filters), std)
        p["convnet/%s/bias" % name] = np.zeros(filters)
        if bn:
            q = "convnet/%s/batch_norm/" % name
            p[q + "gamma"] = np.ones(filters)
            p[q + "beta"] = np.zeros(filters)
            p[q + "moving_mean"] = np.zeros(filters)
            p[q + "moving_variance"] = np.ones(filters)
            if randomize_bn:  # exercise the folding with non-trivial statistics
                p[q + "gamma"] = rng.uniform(0.5, 1.5, filters)
                p[q + "beta"] = rng.normal(0, 0.1, filters)
                p[q + "moving_mean"] = rng.normal(0, 0.2, filters)
                p[q + "moving_variance"] = rng.uniform(0.5, 2.0, filters)
        cin = filters
    I = 256
    for scope, H in (("bdrnn1", sizes[0]), ("bdrnn2", sizes[1])):
        for d in ("fw", "bw"):
            q = "rnn/%s/%s/" % (scope, d)
            if cell_type == "lstm":
                p[q + "lstm_cell/kernel"] = truncated_normal(rng, (I + H, 4 * H), 0.01)
                p[q + "lstm_cell/bias"] = np.zeros(4 * H)
            else:
                p[q + "gru_cell/gates/kernel"] = truncated_normal(rng, (I + H, 2 * H), 0.01)
                p[q + "gru_cell/gates/bias"] = truncated_normal(rng, (2 * H,), 0.01)
                p[q + "gru_cell/candidate/kernel"] = truncated_normal(rng, (I + H, H), 0.01)
                p[q + "gru_cell/candidate/bias"] = truncated_normal(rng, (H,), 0.01)
        I = 2 * H
    p["rnn/logits/kernel"] = truncated_normal(rng, (I, num_classes + 1), np.sqrt(1.3 * 2.0 / I))
    p["rnn/logits/bias"] = np.zeros(num_classes + 1)
    return {k: v.astype(dtype) for k, v in p.items()}
