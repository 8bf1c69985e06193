"""The weinman CNN -> bidirectional LSTM/GRU -> logits graph of the reference on B200 (INFER mode).

Mirrors the reference's Python-level interface (src/weinman/model.py, model_bu.py, validate.py):

  convnet_layers(inputs, widths, mode)                    model.py:126-165
  rnn_layers(features, sequence_length, num_classes)      model.py:202-221  (GRU 512/256)  /  model_bu.py:202-221 (LSTM 512/512)
  preprocess_image(image)                                 validate.py:56-68
  get_output(rnn_logits, sequence_length)                 validate.py:81-92 (greedy decode, dense, -1 padded)
  get_string(labels), out_charset, num_classes()          validate.py:126-129, mjsynth.py:23-26

TensorFlow keeps the weights in variable scopes; here a `Model` object holds them under the same variable
names ("convnet/conv1/kernel", "rnn/bdrnn1/fw/lstm_cell/kernel", ... SURVEY.md App. A.8), importable from /
exportable to an .npz keyed by those names.  All arithmetic runs in libocr_b200.so: the convolutions, the RNN
input / recurrent projections and the logits layer on the tcgen05 tensor cores (TF32 products, fp32 sums),
the rest as coalesced CUDA kernels.  There is no CPU path.
"""
import ctypes

import numpy as np
import torch

from . import _lib, ctc

# mjsynth.py:23-26
out_charset = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789 `~!@#$%^&*()-=_+[]{};'\\:\"|,./<>?"


def num_classes():
    return len(out_charset)


class ModeKeys:  # tf.contrib.learn.ModeKeys
    TRAIN = "train"
    EVAL = "eval"
    INFER = "infer"


# (filters, kernel, padding, name, batch_norm)  -- model.py:47-54
LAYER_PARAMS = [(32, 3, "valid", "conv1", False), (32, 3, "same", "conv2", True),
                (64, 3, "same", "conv3", False), (64, 3, "same", "conv4", True),
                (128, 3, "same", "conv5", False), (128, 3, "same", "conv6", True),
                (256, 3, "same", "conv7", False), (256, 3, "same", "conv8", True)]
# max-pool applied to the INPUT of each 'same' convolution (model.py:134-144): (window_h, window_w, stride_h, stride_w)
_POOL_BEFORE = {"conv2": (1, 1, 1, 1), "conv3": (2, 2, 2, 2), "conv4": (1, 1, 1, 1), "conv5": (2, 2, 2, 1),
                "conv6": (1, 1, 1, 1), "conv7": (2, 2, 2, 1), "conv8": (1, 1, 1, 1)}
BN_EPS = 1e-3


def _truncated_normal(rng, shape, std):
    x = rng.standard_normal(shape)
    bad = np.abs(x) > 2
    while bad.any():           # tf.truncated_normal: resample beyond two standard deviations
        x[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2
    return x * std


def init_params(seed=0, cell_type="lstm", rnn_sizes=(512, 512), num_classes=None, dtype=np.float32):
    """Fresh variables with the reference's initialisers, keyed by TensorFlow variable names: what
    tf.global_variables_initializer() gives train.py's graph.  Convolutions / logits: variance_scaling_initializer
    (factor 2, FAN_IN, truncated normal; model.py:94,207), zero biases (model.py:95,208); RNN cell kernels
    truncated_normal(stddev 0.01) (model.py:170), LSTM biases zero, GRU biases truncated normal; batch-norm gamma 1, beta 0,
    moving mean 0, moving variance 1."""
    n_out = (len(out_charset) if num_classes is None else num_classes) + 1
    rng = np.random.default_rng(seed)
    p = {}
    cin = 1
    for (filters, k, padding, name, bn) in LAYER_PARAMS:
        p["convnet/%s/kernel" % name] = _truncated_normal(rng, (k, k, cin, filters), np.sqrt(1.3 * 2.0 / (k * k * cin)))
        p["convnet/%s/bias" % name] = np.zeros(filters)
        if bn:
            q = "convnet/%s/batch_norm/" % name
            p[q + "gamma"], p[q + "beta"] = np.ones(filters), np.zeros(filters)
            p[q + "moving_mean"], p[q + "moving_variance"] = np.zeros(filters), np.ones(filters)
        cin = filters
    I = 256
    for scope, H in (("bdrnn1", rnn_sizes[0]), ("bdrnn2", rnn_sizes[1])):
        for d in ("fw", "bw"):
            q = "rnn/%s/%s/" % (scope, d)
            if cell_type == "lstm":
                p[q + "lstm_cell/kernel"] = _truncated_normal(rng, (I + H, 4 * H), 0.01)
                p[q + "lstm_cell/bias"] = np.zeros(4 * H)
            else:
                p[q + "gru_cell/gates/kernel"] = _truncated_normal(rng, (I + H, 2 * H), 0.01)
                p[q + "gru_cell/gates/bias"] = _truncated_normal(rng, (2 * H,), 0.01)
                p[q + "gru_cell/candidate/kernel"] = _truncated_normal(rng, (I + H, H), 0.01)
                p[q + "gru_cell/candidate/bias"] = _truncated_normal(rng, (H,), 0.01)
        I = 2 * H
    p["rnn/logits/kernel"] = _truncated_normal(rng, (I, n_out), np.sqrt(1.3 * 2.0 / I))
    p["rnn/logits/bias"] = np.zeros(n_out)
    return {k: v.astype(dtype) for k, v in p.items()}


def preprocess_image(image):
    """validate._preprocess_image: uint8 -> float32 in [-0.5, 0.5].  (Model.convnet_layers also accepts the uint8
    tensor directly and fuses this into conv1.)"""
    return image.to(torch.float32) * float(np.float32(1.0 / 255.0)) - 0.5     # convert_image_dtype multiplies by float32(1/255)


class Placeholders(object):
    """validate._get_input (validate.py:71-78): the pair of placeholders the reference's graph is fed through --
    image uint8 [bucket_size, 32, None, 1], width int32 [bucket_size].  TensorFlow checks a feed against the placeholder's
    dtype and static shape and raises ValueError; `feed` applies the same checks (the width axis is free) and hands back the
    pair as tensors on `device`, ready for Model.convnet_layers / Model.recognize."""

    def __init__(self, bucket_size):
        self.bucket_size = int(bucket_size)
        self.image_shape = (self.bucket_size, 32, None, 1)
        self.width_shape = (self.bucket_size,)

    def feed(self, image, width, device=None):
        image = torch.as_tensor(image)
        width = torch.as_tensor(width)
        if image.dtype != torch.uint8:
            raise ValueError("image: placeholder is uint8, got %s" % image.dtype)
        if image.dim() != 4 or image.shape[0] != self.bucket_size or image.shape[1] != 32 or image.shape[3] != 1:
            raise ValueError("Cannot feed value of shape %s for a placeholder of shape %s" % (tuple(image.shape), self.image_shape))
        if width.dtype not in (torch.int32, torch.int64) or tuple(width.shape) != self.width_shape:
            raise ValueError("Cannot feed value of shape %s / dtype %s for the int32 placeholder of shape %s"
                             % (tuple(width.shape), width.dtype, self.width_shape))
        width = width.to(torch.int32)
        if device is not None:
            image, width = image.to(device), width.to(device)
        return image, width


def get_input(bucket_size):
    """validate._get_input(bucket_size) -> the (image, width) placeholder pair, here one object with a checked `feed`."""
    return Placeholders(bucket_size)


def get_string(labels):
    """validate._get_string: label ids -> text."""
    return "".join(out_charset[int(c)] for c in labels)


def ctc_loss_layer(rnn_logits, sequence_labels, sequence_length):
    """model.ctc_loss_layer (model.py:224-229): mean over the batch of tf.nn.ctc_loss(time_major=True)."""
    return ctc.ctc_loss(sequence_labels, rnn_logits, sequence_length, time_major=True).mean()


def get_testing(rnn_logits, sequence_length, label, label_length):
    """test._get_testing (test.py:75-104) -> (loss, label_error, sequence_error), all scalars:
    CTC loss; beam search (width 128, top path, merge_repeated=True); label_error = sum of edit distances / sum of
    label lengths; sequence_error = fraction of sequences with a non-zero edit distance.
    label: SparseTensor-like triple (indices [N,2], values [N], dense_shape [2]); label_length [B]."""
    loss = ctc_loss_layer(rnn_logits, label, sequence_length)
    predictions, _ = ctc.ctc_beam_search_decoder(rnn_logits, sequence_length, beam_width=128, top_paths=1, merge_repeated=True)
    truth = ctc.SparseTensor(label[0].to(rnn_logits.device), label[1].to(rnn_logits.device), label[2])
    label_errors = ctc.edit_distance(predictions[0], truth, normalize=False)
    total_labels = torch.as_tensor(label_length).to(label_errors.device).sum().to(torch.float32)
    label_error = label_errors.sum() / total_labels
    sequence_error = (label_errors != 0).sum().to(torch.float32) / label_errors.numel()
    return loss, label_error, sequence_error


class _Pending:
    """A batch in flight (Model.recognize_async)."""

    def __init__(self, model, g, host, event):
        self.model, self.g, self.host, self.event = model, g, host, event

    def result(self):
        self.event.synchronize()
        dense = self.host.numpy()
        chars = self.model.__dict__.setdefault("_charset_arr", np.array(list(out_charset)))
        texts = ["".join(chars[row[row >= 0]]) for row in dense]
        self.g["free"].append(self.host)
        return texts


class Model:
    """Weights of the recognizer + the reference's graph-building functions as methods."""

    def __init__(self, params, cell_type="lstm", rnn_sizes=(512, 512), device="cuda", conv_path="igemm", fuse_pool=True):
        # conv_path: "igemm" = implicit GEMM (patches gathered inside the kernel); "im2col" = explicit patch matrix + GEMM
        # fuse_pool: pool2 / pool4 applied in the epilogue of conv2 / conv4 (ocr_conv3x3_same_pool); False = separate ocr_maxpool launches
        self.conv_path = conv_path
        self.fuse_pool = fuse_pool
        if cell_type not in ("lstm", "gru"):
            raise ValueError("cell_type must be 'lstm' (model_bu.py) or 'gru' (model.py)")
        self.cell_type = cell_type
        self.rnn_sizes = tuple(rnn_sizes)
        self.device = torch.device(device)
        self.params = {k: torch.as_tensor(np.asarray(v), dtype=torch.float32).to(self.device) for k, v in params.items()}
        self._prepare()

    # ------------------------------------------------------------------ weights
    @classmethod
    def load_npz(cls, path, **kw):
        """Variables from an .npz keyed by TensorFlow variable names (validate._get_init_trained, validate.py:116-124: a
        Saver over the graph's variables).  Training-only entries of a Trainer checkpoint (Adam slots, beta powers,
        global_step) are ignored, as a Saver built from the inference graph ignores them."""
        skip = ("global_step", "beta1_power", "beta2_power")
        with np.load(path) as z:
            return cls({k: z[k] for k in z.files if k not in skip and not k.endswith(("/Adam", "/Adam_1"))}, **kw)

    def save_npz(self, path):
        np.savez(path, **{k: v.cpu().numpy() for k, v in self.params.items()})

    def _prepare(self):
        """Kernel-side layouts: BN folded into filters/biases, filters as K-major [Cout, 9*Cin], RNN kernels split
        into x / h parts, transposed to [N, K] and stacked fw | bw."""
        p = self.params
        self.conv = {}
        for (filters, k, padding, name, bn) in LAYER_PARAMS:
            w = p["convnet/%s/kernel" % name].double()   # [3,3,Cin,Cout]
            b = p["convnet/%s/bias" % name].double()
            if bn:
                q = "convnet/%s/batch_norm/" % name
                scale = p[q + "gamma"].double() / torch.sqrt(p[q + "moving_variance"].double() + BN_EPS)
                w = w * scale
                b = (b - p[q + "moving_mean"].double()) * scale + p[q + "beta"].double()
            if name == "conv1":
                self.conv[name] = (w.float().contiguous(), b.float().contiguous())
            else:
                self.conv[name] = (w.reshape(-1, filters).t().float().contiguous(), b.float().contiguous())
        self.rnn = []
        I = 256
        for scope, H in (("bdrnn1", self.rnn_sizes[0]), ("bdrnn2", self.rnn_sizes[1])):
            if self.cell_type == "lstm":
                ks = [p["rnn/%s/%s/lstm_cell/kernel" % (scope, d)] for d in ("fw", "bw")]
                wx = torch.cat([k[:I].t() for k in ks], 0).contiguous()          # [8H, I]
                wh = torch.cat([k[I:].t() for k in ks], 0).contiguous()          # [8H, H]
                bias = torch.cat([p["rnn/%s/%s/lstm_cell/bias" % (scope, d)] for d in ("fw", "bw")]).contiguous()
                wh2 = None
                if self.device.type == "cuda" and H % 16 == 0:
                    # recurrent weights pre-arranged (gate-major hidden slices) for the persistent kernel, once
                    wh2 = torch.empty_like(wh)
                    _lib.check(_lib.load().ocr_lstm_prepare_wh(_lib.ptr(wh), H, _lib.ptr(wh2), _lib.stream_handle()), "ocr_lstm_prepare_wh")
                self.rnn.append(dict(I=I, H=H, wx=wx, wh=wh, wh2=wh2, bias=bias))
            else:
                gk = [p["rnn/%s/%s/gru_cell/gates/kernel" % (scope, d)] for d in ("fw", "bw")]
                ck = [p["rnn/%s/%s/gru_cell/candidate/kernel" % (scope, d)] for d in ("fw", "bw")]
                wx = torch.cat([torch.cat([g[:I].t(), c[:I].t()], 0) for g, c in zip(gk, ck)], 0).contiguous()   # [6H, I]
                wh = torch.cat([g[I:].t() for g in gk], 0).contiguous()          # [4H, H]
                wh2 = torch.cat([c[I:].t() for c in ck], 0).contiguous()         # [2H, H]
                bias = torch.cat([torch.cat([p["rnn/%s/%s/gru_cell/gates/bias" % (scope, d)],
                                             p["rnn/%s/%s/gru_cell/candidate/bias" % (scope, d)]]) for d in ("fw", "bw")]).contiguous()
                self.rnn.append(dict(I=I, H=H, wx=wx, wh=wh, wh2=wh2, bias=bias))
            I = 2 * H
        self.logits_w = p["rnn/logits/kernel"].t().contiguous()   # [C, 2H]
        self.logits_b = p["rnn/logits/bias"].contiguous()

    # ------------------------------------------------------------------ graph
    def convnet_layers(self, inputs, widths, mode=ModeKeys.INFER):
        """inputs [B,32,W,1] float32 NHWC (preprocessed) or uint8 (preprocessing fused), widths [B] int32
        -> (features [B,T,256], sequence_length [B] int32).   model.py:126-165"""
        if mode == ModeKeys.TRAIN:
            raise NotImplementedError("Model is the inference graph; TRAIN mode (batch statistics + backward pass) lives in cnn_lstm_ctc_ocr_b200.train.Trainer")
        _lib.require_cuda(inputs)
        lib = _lib.load()
        sh = _lib.stream_handle()
        B, H, W, one = inputs.shape
        if one != 1:
            raise ValueError("inputs must be [B, H, W, 1]")
        x = inputs.contiguous()
        is_u8 = x.dtype == torch.uint8
        if not is_u8:
            x = x.float()
        dev = x.device
        w1, b1 = self.conv["conv1"]
        a = torch.empty((B, H - 2, W - 2, w1.shape[-1]), dtype=torch.float32, device=dev)
        _lib.check(lib.ocr_conv1_3x3_valid(_lib.ptr(x), int(is_u8), B, H, W, _lib.ptr(w1), _lib.ptr(b1), w1.shape[-1], _lib.ptr(a), sh),
                   "ocr_conv1_3x3_valid")
        names = [lp[3] for lp in LAYER_PARAMS]
        pooled_already = False      # the previous layer's launch applied this layer's pool in its epilogue
        for (filters, k, padding, name, bn) in LAYER_PARAMS[1:]:
            ph, pw, s_h, s_w = (1, 1, 1, 1) if pooled_already else _POOL_BEFORE[name]
            pooled_already = False
            Bn, Hn, Wn, Cn = a.shape
            Hp, Wp = (Hn - ph) // s_h + 1, (Wn - pw) // s_w + 1
            if Hp < 1 or Wp < 1:
                raise ValueError("image too small for the convolutional stack (need height 32, width >= 8)")
            wk, bk = self.conv[name]
            if self.conv_path == "igemm":
                if (ph, pw, s_h, s_w) != (1, 1, 1, 1):
                    pooled = torch.empty((Bn, Hp, Wp, Cn), dtype=torch.float32, device=dev)
                    _lib.check(lib.ocr_maxpool(_lib.ptr(a), Bn, Hn, Wn, Cn, ph, pw, s_h, s_w, _lib.ptr(pooled), sh), "ocr_maxpool")
                    a = pooled
                # conv -> (folded batch-norm) -> ReLU -> the pool in front of the next layer as ONE launch where the halo-tile
                # kernel takes the shape (conv2 + pool2, conv4 + pool4): model.py:105-116
                nxt = names.index(name) + 1
                npool = _POOL_BEFORE[names[nxt]] if nxt < len(names) else (1, 1, 1, 1)
                if self.fuse_pool and npool[:3] == (2, 2, 2) and Hp >= 2 and Wp >= 2 and lib.ocr_conv3x3_pool_fused(Bn, Hp, Wp, Cn, filters, npool[3]):
                    Ho, Wo = (Hp - 2) // 2 + 1, (Wp - 2) // npool[3] + 1
                    out = torch.empty((Bn, Ho, Wo, filters), dtype=torch.float32, device=dev)
                    _lib.check(lib.ocr_conv3x3_same_pool(_lib.ptr(a), Bn, Hp, Wp, Cn, _lib.ptr(wk), _lib.ptr(bk), filters, 1, npool[3],
                                                         _lib.ptr(out), sh), "ocr_conv3x3_same_pool")
                    pooled_already = True
                else:
                    out = torch.empty((Bn, Hp, Wp, filters), dtype=torch.float32, device=dev)
                    _lib.check(lib.ocr_conv3x3_same(_lib.ptr(a), Bn, Hp, Wp, Cn, _lib.ptr(wk), _lib.ptr(bk), filters, 1, _lib.ptr(out), sh),
                               "ocr_conv3x3_same")
                a = out
            else:
                patches = torch.empty((Bn * Hp * Wp, 9 * Cn), dtype=torch.float32, device=dev)
                _lib.check(lib.ocr_im2col3x3_same(_lib.ptr(a), Bn, Hn, Wn, Cn, ph, pw, s_h, s_w, _lib.ptr(patches), sh), "ocr_im2col3x3_same")
                a = torch.empty((Bn, Hp, Wp, filters), dtype=torch.float32, device=dev)
                _lib.check(lib.ocr_gemm_tf32(_lib.ptr(patches), 9 * Cn, _lib.ptr(wk), 9 * Cn, _lib.ptr(bk), _lib.ptr(a), filters,
                                             Bn * Hp * Wp, filters, 9 * Cn, 1, sh), "ocr_gemm_tf32")
                del patches
        Bn, Hn, Wn, Cn = a.shape
        seq = torch.empty((Wn, Bn, Cn), dtype=torch.float32, device=dev)   # pool8 (all Hn = 3 rows) + squeeze, time-major
        _lib.check(lib.ocr_rows_max_to_seq(_lib.ptr(a), Bn, Hn, Wn, Cn, _lib.ptr(seq), sh), "ocr_rows_max_to_seq")
        widths = torch.as_tensor(widths).to(device=dev, dtype=torch.int32)
        sequence_length = torch.div(widths - 2, 2, rounding_mode="floor") - 2   # model.py:152-163
        return seq.transpose(0, 1), sequence_length.to(torch.int32)

    def rnn_layer(self, seq, sequence_length, layer):
        """seq [T,B,I] time-major -> [T,B,2H]   (model.py:167-199)"""
        lib = _lib.load()
        L = self.rnn[layer]
        T, B, I = seq.shape
        H = L["H"]
        cell = 0 if self.cell_type == "lstm" else 1
        need = ctypes.c_size_t(0)
        _lib.check(lib.ocr_birnn_workspace_bytes(cell, T, B, H, ctypes.byref(need)), "ocr_birnn_workspace_bytes")
        ws = torch.empty(need.value, dtype=torch.uint8, device=seq.device)
        out = torch.empty((T, B, 2 * H), dtype=torch.float32, device=seq.device)
        _lib.check(lib.ocr_birnn_layer(cell, _lib.ptr(seq), T, B, I, H, _lib.ptr(sequence_length), _lib.ptr(L["wx"]), _lib.ptr(L["wh"]),
                                       _lib.ptr(L["wh2"]), _lib.ptr(L["bias"]), _lib.ptr(out), _lib.ptr(ws), need.value,
                                       _lib.stream_handle()), "ocr_birnn_layer")
        return out

    def rnn_layers(self, features, sequence_length, num_classes=None):
        """features [B,T,256], sequence_length [B] -> logits [T,B,num_classes+1] (dense + ReLU).   model.py:202-221"""
        _lib.require_cuda(features)
        lib = _lib.load()
        C = self.logits_w.shape[0]
        if num_classes is not None and num_classes + 1 != C:
            raise ValueError("this model's logits layer has %d outputs, not num_classes+1 = %d" % (C, num_classes + 1))
        seq = features.transpose(0, 1).contiguous().float()   # time-major (model.py:212)
        sl = sequence_length.to(device=seq.device, dtype=torch.int32).contiguous()
        r1 = self.rnn_layer(seq, sl, 0)
        r2 = self.rnn_layer(r1, sl, 1)
        T, B, F = r2.shape
        logits = torch.empty((T, B, C), dtype=torch.float32, device=seq.device)
        _lib.check(lib.ocr_gemm_tf32(_lib.ptr(r2), F, _lib.ptr(self.logits_w), F, _lib.ptr(self.logits_b), _lib.ptr(logits), C,
                                     T * B, C, F, 1, _lib.stream_handle()), "ocr_gemm_tf32")
        return logits

    def get_output(self, rnn_logits, sequence_length):
        """validate._get_output: greedy CTC decode -> [dense int64 [B,Lmax], -1 padded]."""
        predictions, _ = ctc.ctc_greedy_decoder(rnn_logits, sequence_length, merge_repeated=True)
        return [ctc.sparse_tensor_to_dense(predictions[0], default_value=-1)]

    def recognize(self, images, widths, use_graph=True):
        """images [B,32,W,1] uint8 or float -> list of strings: the graph LocalServer.run builds (server.py:80-89)
        plus its post-processing (server.py:134-138).

        uint8 batches (what the server feeds, server.py:71-78) replay a CUDA graph recorded per batch shape -- the
        reference builds its TensorFlow graph once per process for the same reason: the ~20 launches of a batch cost more
        host time than device time.  All recorded shapes share one memory pool (they never run concurrently)."""
        if not (use_graph and images.dtype == torch.uint8 and self.device.type == "cuda"):
            features, sl = self.convnet_layers(images.to(self.device), widths, ModeKeys.INFER)
            logits = self.rnn_layers(features, sl)
            dense = self.get_output(logits, sl)[0].cpu().numpy()
            return [get_string([c for c in row if c >= 0]) for row in dense]
        return self.recognize_async(images, widths).result()

    def recognize_async(self, images, widths):
        """Enqueue one uint8 batch (copy in, graph replay, copy out) and return a handle; handle.result() waits for it and
        gives the strings.  Lets a caller prepare the next batch on the host while this one runs (server.LocalServer.flush)."""
        key = tuple(images.shape)
        cache = self.__dict__.setdefault("_graphs", {})
        g = cache.get(key)
        if g is None:
            g = self._record(key)
            cache[key] = g
        g["img"].copy_(images, non_blocking=True)
        g["widths"].copy_(torch.as_tensor(widths).to(torch.int32), non_blocking=True)
        g["graph"].replay()
        host = g["free"].pop() if g["free"] else torch.empty(tuple(g["dec"].shape), dtype=torch.int64).pin_memory()
        host.copy_(g["dec"], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        return _Pending(self, g, host, ev)

    def _record(self, shape):
        from . import ctc as _ctc
        dev = self.device
        g = dict(img=torch.zeros(shape, dtype=torch.uint8, device=dev), widths=torch.full((shape[0],), shape[2], dtype=torch.int32, device=dev))

        def run():
            features, sl = self.convnet_layers(g["img"], g["widths"], ModeKeys.INFER)
            logits = self.rnn_layers(features, sl)
            dec, ln, ns = _ctc.ctc_greedy_decode_raw(logits, sl)
            return dec
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            run()                                   # warm-up outside the capture (lazy one-time initialisation)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        if "_graph_pool" not in self.__dict__:
            self._graph_pool = torch.cuda.graph_pool_handle()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, pool=self._graph_pool):
            g["dec"] = run()
        g["graph"] = gr
        g["free"] = []          # pinned result buffers, recycled by _Pending.result
        return g
