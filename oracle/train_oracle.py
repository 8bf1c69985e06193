"""TEST INFRASTRUCTURE ONLY -- float64 restatement of the reference's TRAINING step with torch autograd on the CPU.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.

Restates train.main's step (/root/reference/src/weinman/train.py:168-199):
  convnet_layers(mode=TRAIN)  model.py:126-165   batch-norm with BATCH statistics + moving-average update (train.py:116-118)
  rnn_layers                  model.py:202-221 / model_bu.py:202-221
  ctc_loss_layer              model.py:224-229   mean over the batch of tf.nn.ctc_loss
  _get_training               train.py:101-141   exponential_decay (non-staircase) + AdamOptimizer(beta1=momentum) via
                                                 optimize_loss over all variables of scope "convnet|rnn", no clipping
TensorFlow semantics per SURVEY.md App. A.2 (batch-norm), A.3 (cells), A.4 (CTC), A.7 (optimiser).  The forward
is the same mathematics as oracle/model_oracle.py written with differentiable torch ops, so autograd supplies the
reference gradients of every variable; PARITY against live TensorFlow is "unpinned" (TF cannot run here).
Cross-check: tests/test_train_oracle.py compares its INFER forward and its CTC gradient with the numpy / C oracles.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import model_oracle as mo

BN_MOMENTUM = 0.99  # tf.layers.batch_normalization default


def _conv(x, kernel, bias, padding):
    # x [B,H,W,C] -> NCHW for torch, HWIO -> OIHW
    y = F.conv2d(x.permute(0, 3, 1, 2), kernel.permute(3, 2, 0, 1), bias, padding=0 if padding == "valid" else 1)
    return y.permute(0, 2, 3, 1)


def _pool(x, window, strides):
    return F.max_pool2d(x.permute(0, 3, 1, 2), window, strides).permute(0, 2, 3, 1)


def forward_train(params, images, widths, cell_type="lstm", sizes=(512, 512), fused_unbiased_moving_var=True):
    """params: dict name -> torch float64 tensor (requires_grad on the trainable ones); images [B,32,W,1] float64
    (preprocessed).  Returns (logits [T,B,C], seq_len, new_moving_stats dict)."""
    x = images
    new_stats = {}
    for (filters, k, padding, name, bn) in mo.LAYER_PARAMS:
        x = _conv(x, params["convnet/%s/kernel" % name], params["convnet/%s/bias" % name], padding)
        if bn:
            q = "convnet/%s/batch_norm/" % name
            mean = x.mean(dim=(0, 1, 2))
            var = x.var(dim=(0, 1, 2), unbiased=False)
            n = x.numel() // x.shape[-1]
            x = params[q + "gamma"] * (x - mean) / torch.sqrt(var + mo.BN_EPS) + params[q + "beta"]
            mv = var * (n / (n - 1.0)) if fused_unbiased_moving_var else var   # fused batch-norm feeds the unbiased variance
            new_stats[q + "moving_mean"] = BN_MOMENTUM * params[q + "moving_mean"] + (1 - BN_MOMENTUM) * mean.detach()
            new_stats[q + "moving_variance"] = BN_MOMENTUM * params[q + "moving_variance"] + (1 - BN_MOMENTUM) * mv.detach()
        x = torch.relu(x)
        if name == "conv2":
            x = _pool(x, (2, 2), (2, 2))
        elif name in ("conv4", "conv6"):
            x = _pool(x, (2, 2), (2, 1))
        elif name == "conv8":
            x = _pool(x, (3, 1), (3, 1))
    feats = x[:, 0]                                  # [B,T,256]
    seq_len = ((np.asarray(widths, np.int64) - 2) // 2 - 2).astype(np.int32)
    seq = feats.permute(1, 0, 2)                     # time-major
    I = 256
    for scope, H in (("bdrnn1", sizes[0]), ("bdrnn2", sizes[1])):
        outs = []
        for d, reverse in (("fw", False), ("bw", True)):
            outs.append(_run_direction(seq, seq_len, params, "rnn/%s/%s/" % (scope, d), cell_type, H, reverse))
        seq = torch.cat(outs, dim=2)
        I = 2 * H
    logits = torch.relu(seq @ params["rnn/logits/kernel"] + params["rnn/logits/bias"])
    return logits, seq_len, new_stats


def _run_direction(seq, seq_len, params, p, cell_type, H, reverse, probe=None):
    """One direction of tf.nn.bidirectional_dynamic_rnn.  probe (LSTM only): a zero tensor [T,B,4H] with
    requires_grad, added to the gate pre-activations of the step that visits frame t -- after backward() its .grad is the
    gradient of the pre-activations frame by frame (what ocr_birnn_lstm_bwd leaves in `gates`)."""
    T, B, _ = seq.shape
    h = torch.zeros((B, H), dtype=seq.dtype)
    c = torch.zeros((B, H), dtype=seq.dtype)
    rows = [[None] * B for _ in range(T)]
    idx = torch.arange(B)
    sl = torch.as_tensor(np.asarray(seq_len, np.int64))
    outs = torch.zeros((T, B, H), dtype=seq.dtype)
    out_list = []
    for s in range(T):
        live = s < sl
        t = torch.where(live, (sl - 1 - s) if reverse else torch.full_like(sl, s), torch.zeros_like(sl))
        x = seq[t, idx]
        if cell_type == "lstm":
            z = torch.cat([x, h], dim=1) @ params[p + "lstm_cell/kernel"] + params[p + "lstm_cell/bias"]
            if probe is not None:
                z = z + probe[t, idx] * live[:, None].to(seq.dtype)
            i, j, f, o = z[:, :H], z[:, H:2 * H], z[:, 2 * H:3 * H], z[:, 3 * H:]
            c_new = torch.sigmoid(f + 1.0) * c + torch.sigmoid(i) * torch.tanh(j)
            h_new = torch.sigmoid(o) * torch.tanh(c_new)
        else:
            g = torch.sigmoid(torch.cat([x, h], dim=1) @ params[p + "gru_cell/gates/kernel"] + params[p + "gru_cell/gates/bias"])
            r, u = g[:, :H], g[:, H:]
            cand = torch.tanh(torch.cat([x, r * h], dim=1) @ params[p + "gru_cell/candidate/kernel"] + params[p + "gru_cell/candidate/bias"])
            h_new = u * h + (1 - u) * cand
            c_new = c
        m = live[:, None].to(seq.dtype)
        h = m * h_new + (1 - m) * h
        c = m * c_new + (1 - m) * c
        out_list.append((t, live, h))
    # scatter: out[t_b, b] = h at the step that visited frame t_b (index_put keeps autograd)
    for (t, live, hh) in out_list:
        mask = live[:, None].to(seq.dtype)
        upd = torch.zeros((T, B, H), dtype=seq.dtype).index_put((t, idx), hh * mask, accumulate=True)
        outs = outs + upd
    return outs


def ctc_mean_loss(logits, labels, seq_len):
    """model.ctc_loss_layer: mean over the batch of the per-example CTC loss (blank = C-1)."""
    T, B, C = logits.shape
    lp = torch.log_softmax(logits, dim=-1)
    tgt = torch.tensor([v for l in labels for v in l], dtype=torch.long)
    losses = F.ctc_loss(lp, tgt, torch.as_tensor(np.asarray(seq_len, np.int64)), torch.tensor([len(l) for l in labels]),
                        blank=C - 1, reduction="none", zero_infinity=False)
    return losses.mean(), losses


def learning_rate(step, lr0=1e-4, decay_steps=65536.0, decay_rate=0.9):
    """tf.train.exponential_decay, staircase=False (train.py:120-126)."""
    return lr0 * decay_rate ** (step / decay_steps)


def adam_step(p, g, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer update ("epsilon hat" form, App. A.7); t = 1 for the first step."""
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    lr_t = lr * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    return p - lr_t * m / (np.sqrt(v) + eps), m, v


TRAINABLE = lambda name: not ("moving_mean" in name or "moving_variance" in name)


def train_step_reference(params_np, images_u8, widths, labels, step, cell_type="lstm", sizes=(512, 512), adam_state=None, dtype=np.float64):
    """One reference training step, float64 by default (the parity oracle); dtype=np.float32 is the reference's own arithmetic type
    (bench.py's CPU baseline leg).  Returns dict(loss, grads, new_params, new_stats, adam_state)."""
    params = {k: torch.tensor(np.asarray(v, dtype), requires_grad=TRAINABLE(k)) for k, v in params_np.items()}
    x = torch.tensor(mo.preprocess_image(images_u8).astype(dtype))
    logits, seq_len, new_stats = forward_train(params, x, widths, cell_type, sizes)
    loss, losses = ctc_mean_loss(logits, labels, seq_len)
    loss.backward()
    grads = {k: p.grad.numpy().copy() for k, p in params.items() if p.requires_grad}
    lr = learning_rate(step)
    adam_state = adam_state or {k: (np.zeros_like(g), np.zeros_like(g)) for k, g in grads.items()}
    new_params, new_adam = {}, {}
    for k, v in params_np.items():
        if k in grads:
            pnew, m, vv = adam_step(np.asarray(v, dtype), grads[k], adam_state[k][0], adam_state[k][1], step + 1, lr)
            new_params[k], new_adam[k] = pnew, (m, vv)
        else:
            new_params[k] = new_stats[k].numpy()
    return dict(loss=float(loss.detach()), losses=losses.detach().numpy(), logits=logits.detach().numpy(), grads=grads, new_params=new_params,
                adam_state=new_adam, lr=lr, seq_len=seq_len)
