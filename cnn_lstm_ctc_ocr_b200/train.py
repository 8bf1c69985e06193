"""The reference's training step on B200 (src/weinman/train.py), replayed on libocr_b200.so.

  Trainer(params, ...)                       the variables of scope "convnet|rnn" (train.py:105-111) + Adam slots
  Trainer.train_step(image, width, label)    one `sess.run(train_op)` (train.py:175-199):
        convnet_layers(mode=TRAIN)  model.py:126-165   batch-norm with batch statistics, moving averages updated
        rnn_layers                  model_bu.py:202-221 (bidirectional LSTM 512/512)
        ctc_loss_layer              model.py:224-229
        _get_training               train.py:101-141   exponential_decay (non-staircase) + AdamOptimizer(beta1=momentum)
  learning_rate(step)                        tf.train.exponential_decay as configured by train.py:36-45,120-126

TensorFlow derives the backward graph by itself; here it is written out: every line below is one call into the C ABI
(include/ocr_b200.h, "Training step").  All trainable variables live in ONE flat float32 buffer in TensorFlow layout
(ordered logits, bdrnn2, bdrnn1, conv8 .. conv1 = the order their gradients become available), with gradients and the
Adam slots in matching flat buffers: Adam is one kernel launch, and data-parallel training all-reduces two contiguous
buckets over NCCL (RNN + logits as soon as the recurrent layers are done, the convolutional stack at the end) while the
remaining backward pass is still running.  Per-replica batch-norm statistics by default (as TensorFlow towers would);
sync_bn=True all-reduces the per-channel sums so that N replicas reproduce the single-GPU step exactly.
PyTorch supplies device memory, streams and torch.distributed only.  There is no CPU path.
"""
import contextlib
import ctypes
import math

import numpy as np
import torch

from . import _lib, ctc
from .model import BN_EPS, LAYER_PARAMS, _POOL_BEFORE, Model, ModeKeys

BN_MOMENTUM = 0.99   # tf.layers.batch_normalization default (model.py:120)


def learning_rate(step, learning_rate=1e-4, decay_steps=2 ** 16, decay_rate=0.9, staircase=False):
    """tf.train.exponential_decay (train.py:120-126; flags train.py:36-45)."""
    e = step / float(decay_steps)
    if staircase:
        e = math.floor(e)
    return learning_rate * decay_rate ** e


def flat_layout(shapes, cell_type="lstm"):
    """Offsets of the trainable variables in the flat parameter / gradient / Adam buffers and the boundary between the two
    all-reduce buckets.  shapes: name -> shape.  Returns (names in buffer order, offsets, floats in bucket 1 (logits + RNN),
    total floats).  Host logic only (shared by Trainer and the CPU tests of the data-parallel exchange)."""
    names, n_rnn = _param_order(cell_type)
    offsets, off, first = {}, 0, 0
    for i, n in enumerate(names):
        if i == n_rnn:
            first = off
        offsets[n] = off
        off += (int(np.prod(shapes[n])) + 63) // 64 * 64     # 256-byte aligned slots (TMA operands, float4)
    return names, offsets, first, off


def allreduce_buckets(flat_grad, n_first, world, all_reduce, async_first=True):
    """The data-parallel exchange of a step: SUM all-reduce of bucket 1 (issued when the recurrent layers' gradients are
    complete) and bucket 2 (the convolutional stack), then the 1/world scale the optimiser applies.  `all_reduce(tensor)` is
    the collective (NCCL on GPU; gloo in the CPU tests).  Returns the scale to apply."""
    all_reduce(flat_grad[:n_first])
    all_reduce(flat_grad[n_first:])
    return 1.0 / world


def _is_trainable(name):
    return "moving_mean" not in name and "moving_variance" not in name


def _param_order(cell_type):
    names = ["rnn/logits/kernel", "rnn/logits/bias"]
    for scope in ("bdrnn2", "bdrnn1"):
        if cell_type == "lstm":
            names += ["rnn/%s/fw/lstm_cell/kernel" % scope, "rnn/%s/bw/lstm_cell/kernel" % scope,
                      "rnn/%s/fw/lstm_cell/bias" % scope, "rnn/%s/bw/lstm_cell/bias" % scope]
        else:
            for d in ("fw", "bw"):
                names += ["rnn/%s/%s/gru_cell/%s" % (scope, d, v) for v in ("gates/kernel", "gates/bias", "candidate/kernel", "candidate/bias")]
    n_rnn = len(names)
    for (filters, k, padding, name, bn) in reversed(LAYER_PARAMS):
        names += ["convnet/%s/kernel" % name, "convnet/%s/bias" % name]
        if bn:
            names += ["convnet/%s/batch_norm/gamma" % name, "convnet/%s/batch_norm/beta" % name]
    return names, n_rnn


class Trainer:
    def __init__(self, params, cell_type="lstm", rnn_sizes=(512, 512), device="cuda", learning_rate=1e-4, momentum=0.9,
                 decay_rate=0.9, decay_steps=2 ** 16, decay_staircase=False, beta2=0.999, epsilon=1e-8,
                 process_group=None, sync_bn=False, global_step=0, overlap_weight_gradients=True, tune_scope="", tune_from=None):
        if cell_type not in ("lstm", "gru"):
            raise ValueError("cell_type must be 'lstm' (model_bu.py) or 'gru' (model.py)")
        self.cell_type = cell_type
        self.rnn_sizes = tuple(rnn_sizes)
        if cell_type == "lstm" and any(h % 16 for h in self.rnn_sizes):
            # the fw | bw LSTM biases are read (and their gradient written) as ONE [8H] slice of the flat buffer, whose
            # slots are padded to 64 floats: adjacent only when 4H % 64 == 0
            raise ValueError("LSTM sizes must be multiples of 16, got %s" % (self.rnn_sizes,))
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.OcrLibraryError("Trainer needs a CUDA device; there is no CPU path")
        self.lib = _lib.load()
        self.hp = dict(learning_rate=learning_rate, decay_steps=decay_steps, decay_rate=decay_rate, staircase=decay_staircase)
        self.beta1, self.beta2, self.epsilon = momentum, beta2, epsilon
        self.global_step = int(global_step)
        self.pg = process_group
        self.world = 1
        if process_group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(process_group if process_group is not True else None)
        self.sync_bn = bool(sync_bn) and self.world > 1

        names, n_rnn = _param_order(cell_type)
        missing = [n for n in names if n not in params]
        if missing:
            raise KeyError("parameters missing: %s" % missing[:3])
        self.names = names
        self.shapes = {n: tuple(np.asarray(params[n]).shape) for n in names}
        _, self.offsets, self.n_rnn_floats, off = flat_layout(self.shapes, cell_type)
        self.n_floats = off
        dev = self.device
        self.theta = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        self.adam_m = torch.zeros(off, dtype=torch.float32, device=dev)
        self.adam_v = torch.zeros(off, dtype=torch.float32, device=dev)
        self.params, self.grads = {}, {}
        for n in names:
            o, sz = self.offsets[n], int(np.prod(self.shapes[n]))
            self.params[n] = self.theta[o:o + sz].view(self.shapes[n])
            self.grads[n] = self.grad[o:o + sz].view(self.shapes[n])
            self.params[n].copy_(torch.as_tensor(np.asarray(params[n]), dtype=torch.float32))
        self.stats = {n: torch.as_tensor(np.asarray(v), dtype=torch.float32).to(dev).contiguous() for n, v in params.items() if not _is_trainable(n)}
        self.scratch = torch.zeros(16 * 9 * 4096, dtype=torch.uint8, device=dev)   # double sums: up to 8H = 4096 columns x 2
        self.zero_bias = torch.zeros(1024, dtype=torch.float32, device=dev)
        # Weight gradients are off the critical path of the backward pass (nothing downstream reads them before the
        # all-reduce / Adam), while the chain of input gradients -- above all the frame-by-frame BPTT loops -- leaves most
        # SMs idle: the transposes and split-K contractions of every layer's weight gradient run on a side stream.
        self.overlap = bool(overlap_weight_gradients)
        # pooled batch-norm layers keep the arguments of the pool's maxima instead of the full-size activation (tests switch it
        # off to compare with the separate max-pool gradient)
        self.pool_arg = True
        self.blocked_planar = True   # K-blocked planar operands for the conv filter gradients (ocr_gemm_tf32_wgrad_blocked)
        self.early_planar = True     # planar copies of the conv inputs (weight-gradient operands) made during the recurrent layers
        self.side = torch.cuda.Stream(device=dev) if self.overlap else None
        self.scratch_side = torch.zeros_like(self.scratch)
        self._side_keep = []
        self.wscratch = None
        self._set_tune_scope(tune_scope)
        self._alloc_derived()
        self.derive_layouts()
        if tune_from:
            self.restore(tune_from)

    # ------------------------------------------------------------------ helpers
    def _c(self, rc, what):
        _lib.check(rc, what)

    def _sh(self):
        return _lib.stream_handle()

    def _new(self, *shape):
        return torch.empty(shape, dtype=torch.float32, device=self.device)

    @contextlib.contextmanager
    def _on_side(self, *keep):
        """Enqueue the enclosed work on the side stream, after everything issued so far on the current stream.  `keep`:
        tensors the side work reads; they stay referenced until _join_side."""
        if not self.overlap:
            yield
            return
        self.side.wait_stream(torch.cuda.current_stream(self.device))
        self._side_keep.extend(keep)
        with torch.cuda.stream(self.side):
            yield

    def _join_side(self):
        if self.overlap:
            torch.cuda.current_stream(self.device).wait_stream(self.side)
        self._side_keep.clear()

    def all_params(self):
        """name -> tensor in TensorFlow variable naming (trainable variables + batch-norm moving statistics)."""
        d = dict(self.params)
        d.update(self.stats)
        return d

    def to_model(self, **kw):
        """An inference Model over (a copy of) the current variables."""
        return Model({k: v.detach().cpu().numpy() for k, v in self.all_params().items()}, cell_type=self.cell_type,
                     rnn_sizes=self.rnn_sizes, device=self.device, **kw)

    def save_npz(self, path, with_optimizer=True):
        """tf.train.Saver over the training graph (train.py:185-201): the variables, global_step and -- with_optimizer -- the
        Adam slots under TensorFlow's slot names ("<variable>/Adam" = m, "<variable>/Adam_1" = v).  The beta powers TF also
        saves (beta1_power, beta2_power) are functions of global_step here (_lr_t) and are written for completeness."""
        d = {k: v.detach().cpu().numpy() for k, v in self.all_params().items()}
        d["global_step"] = np.int64(self.global_step)
        if with_optimizer:
            m, v = self.adam_m.cpu().numpy(), self.adam_v.cpu().numpy()
            for n in self.names:
                o, sz = self.offsets[n], int(np.prod(self.shapes[n]))
                d[n + "/Adam"] = m[o:o + sz].reshape(self.shapes[n])
                d[n + "/Adam_1"] = v[o:o + sz].reshape(self.shapes[n])
            d["beta1_power"] = np.float32(self.beta1 ** (self.global_step + 1))
            d["beta2_power"] = np.float32(self.beta2 ** (self.global_step + 1))
        np.savez(path, **d)

    def load_optimizer_state(self, state):
        """Restore the Adam slots saved by save_npz (name -> array with "<variable>/Adam", "<variable>/Adam_1").  Variables
        without saved slots keep zeros (what TensorFlow gives freshly created slots)."""
        for n in self.names:
            o, sz = self.offsets[n], int(np.prod(self.shapes[n]))
            for suffix, buf in (("/Adam", self.adam_m), ("/Adam_1", self.adam_v)):
                if n + suffix in state:
                    a = np.asarray(state[n + suffix], dtype=np.float32)
                    if a.shape != self.shapes[n]:
                        raise ValueError("optimizer slot %s has shape %s, variable has %s" % (n + suffix, a.shape, self.shapes[n]))
                    buf[o:o + sz].copy_(torch.from_numpy(a.reshape(-1)))

    @classmethod
    def load_npz(cls, path, **kw):
        """A Trainer resumed from a checkpoint written by save_npz -- the Supervisor's restore from FLAGS.output
        (train.py:185-201): variables, batch-norm moving statistics, Adam slots and global_step."""
        with np.load(path) as z:
            state = {k: z[k] for k in z.files}
        variables = {k: v for k, v in state.items() if not k.endswith(("/Adam", "/Adam_1")) and k not in ("beta1_power", "beta2_power", "global_step")}
        kw.setdefault("global_step", int(state.get("global_step", 0)))
        t = cls(variables, **kw)
        t.load_optimizer_state(state)
        return t

    def restore(self, tune_from):
        """train._get_init_pretrained (train.py:152-165): `--tune_from` restores a Saver over ALL global variables of the
        training graph -- the model variables, the batch-norm moving statistics, the Adam slots and global_step -- from a
        checkpoint (.npz keyed by TensorFlow names as written by save_npz, or a dict).  Like tf.train.Saver.restore, a
        model variable missing from the checkpoint is an error; a checkpoint without optimizer slots (exported from an
        inference Model) leaves the slots and global_step as they are."""
        if isinstance(tune_from, dict):
            src = tune_from
        else:
            with np.load(tune_from) as z:
                src = {k: z[k] for k in z.files}
        names = list(self.params) + list(self.stats)
        missing = [n for n in names if n not in src]
        if missing:
            raise KeyError("checkpoint lacks %d variable(s), e.g. %s" % (len(missing), missing[0]))
        for n in names:
            dst = self.params[n] if n in self.params else self.stats[n]
            a = torch.as_tensor(np.asarray(src[n]), dtype=torch.float32)
            if tuple(a.shape) != tuple(dst.shape):
                raise ValueError("checkpoint variable %s has shape %s, graph has %s" % (n, tuple(a.shape), tuple(dst.shape)))
            dst.copy_(a)
        self.load_optimizer_state(src)
        if "global_step" in src:
            self.global_step = int(src["global_step"])
        self.derive_layouts()

    def _set_tune_scope(self, tune_scope):
        """`--tune_scope` (train.py:105-111,132-137): only the TRAINABLE variables whose name matches the scope are handed to
        the optimiser (tf.get_collection(..., scope=s) keeps names for which re.match(s, name) succeeds; default
        "convnet|rnn" = all of them).  The others keep their values; batch-norm moving averages update regardless
        (UPDATE_OPS, train.py:116-118).  Stored as contiguous runs of the flat buffer: one Adam launch per run."""
        import re
        self.tune_scope = tune_scope or "convnet|rnn"
        pat = re.compile(self.tune_scope)
        self.tuned = [n for n in self.names if pat.match(n)]
        if not self.tuned:
            raise ValueError("No variables to optimize.")     # tf.contrib.layers.optimize_loss would have nothing to minimise
        runs = []
        for n in self.names:
            if not pat.match(n):
                continue
            o = self.offsets[n]
            e = o + (int(np.prod(self.shapes[n])) + 63) // 64 * 64
            if runs and runs[-1][1] == o:
                runs[-1][1] = e
            else:
                runs.append([o, e])
        self.adam_runs = [tuple(r) for r in runs]

    # ------------------------------------------------------------------ kernel-side weight layouts
    def _alloc_derived(self):
        self.conv_w = {}
        cin = 1
        for (filters, k, padding, name, bn) in LAYER_PARAMS:
            if name != "conv1":
                self.conv_w[name] = (self._new(filters, 9 * cin), self._new(cin, 9 * filters))
            cin = filters
        self.rnn_w = []
        I = 256
        for H in self.rnn_sizes:
            if self.cell_type == "lstm":
                self.rnn_w.append(dict(I=I, H=H, wx=self._new(8 * H, I), wh=self._new(8 * H, H), wxcat=self._new(I, 8 * H), wh_rows=self._new(2 * H, 4 * H)))
            else:
                self.rnn_w.append(dict(I=I, H=H, wx=self._new(6 * H, I), whg=self._new(4 * H, H), whc=self._new(2 * H, H), bias=self._new(6 * H),
                                       wxcat=self._new(I, 6 * H), wg_rows=self._new(2 * H, 2 * H), wc_rows=self._new(2 * H, H)))
            I = 2 * H
        C = self.shapes["rnn/logits/kernel"][1]
        self.logits_w = self._new(C, I)

    def derive_layouts(self):
        """Re-derive the operand layouts of the kernels from the TensorFlow-layout variables (after every update)."""
        lib, sh = self.lib, self._sh()
        cin = 1
        for (filters, k, padding, name, bn) in LAYER_PARAMS:
            if name != "conv1":
                wf, wd = self.conv_w[name]
                self._c(lib.ocr_conv_filter_layouts(_lib.ptr(self.params["convnet/%s/kernel" % name]), cin, filters, _lib.ptr(wf), _lib.ptr(wd), sh),
                        "ocr_conv_filter_layouts")
            cin = filters
        for scope, L in zip(("bdrnn1", "bdrnn2"), self.rnn_w):
            I, H = L["I"], L["H"]
            if self.cell_type == "gru":
                self._derive_gru(scope, L)
                continue
            for d, dn in enumerate(("fw", "bw")):
                kern = self.params["rnn/%s/%s/lstm_cell/kernel" % (scope, dn)]      # [I+H, 4H]
                kp = kern.data_ptr()
                # x part [I,4H] -> wx rows d*4H.. ([4H, I]);  h part [H,4H] -> wh rows d*4H.. ([4H, H])
                self._c(lib.ocr_transpose(ctypes.c_void_p(kp), I, 4 * H, 4 * H, ctypes.c_void_p(L["wx"].data_ptr() + d * 4 * H * I * 4), I, 0, sh), "ocr_transpose")
                self._c(lib.ocr_transpose(ctypes.c_void_p(kp + I * 4 * H * 4), H, 4 * H, 4 * H, ctypes.c_void_p(L["wh"].data_ptr() + d * 4 * H * H * 4), H, 0, sh), "ocr_transpose")
                self._c(lib.ocr_copy_2d(ctypes.c_void_p(kp), 4 * H, ctypes.c_void_p(L["wxcat"].data_ptr() + d * 4 * H * 4), 8 * H, I, 4 * H, sh), "ocr_copy_2d")
                self._c(lib.ocr_copy_2d(ctypes.c_void_p(kp + I * 4 * H * 4), 4 * H, ctypes.c_void_p(L["wh_rows"].data_ptr() + d * H * 4 * H * 4), 4 * H, H, 4 * H, sh), "ocr_copy_2d")
        kl = self.params["rnn/logits/kernel"]                                        # [2H, C]
        self._c(lib.ocr_transpose(_lib.ptr(kl), kl.shape[0], kl.shape[1], kl.shape[1], _lib.ptr(self.logits_w), kl.shape[0], 0, sh), "ocr_transpose")

    def _derive_gru(self, scope, L):
        """GRUCell variables (gates/kernel [I+H,2H], candidate/kernel [I+H,H] + biases, per direction) -> kernel operands."""
        lib, sh = self.lib, self._sh()
        I, H = L["I"], L["H"]
        vp = ctypes.c_void_p
        for d, dn in enumerate(("fw", "bw")):
            q = "rnn/%s/%s/gru_cell/" % (scope, dn)
            gk, ck = self.params[q + "gates/kernel"].data_ptr(), self.params[q + "candidate/kernel"].data_ptr()
            gb, cb = self.params[q + "gates/bias"].data_ptr(), self.params[q + "candidate/bias"].data_ptr()
            f = 4  # bytes
            # forward operands: wx rows d*3H.. = [gates x-part^T (2H) ; candidate x-part^T (H)], whg rows d*2H.., whc rows d*H..
            self._c(lib.ocr_transpose(vp(gk), I, 2 * H, 2 * H, vp(L["wx"].data_ptr() + d * 3 * H * I * f), I, 0, sh), "ocr_transpose")
            self._c(lib.ocr_transpose(vp(ck), I, H, H, vp(L["wx"].data_ptr() + (d * 3 * H + 2 * H) * I * f), I, 0, sh), "ocr_transpose")
            self._c(lib.ocr_transpose(vp(gk + I * 2 * H * f), H, 2 * H, 2 * H, vp(L["whg"].data_ptr() + d * 2 * H * H * f), H, 0, sh), "ocr_transpose")
            self._c(lib.ocr_transpose(vp(ck + I * H * f), H, H, H, vp(L["whc"].data_ptr() + d * H * H * f), H, 0, sh), "ocr_transpose")
            self._c(lib.ocr_copy_2d(vp(gb), 2 * H, vp(L["bias"].data_ptr() + d * 3 * H * f), 2 * H, 1, 2 * H, sh), "ocr_copy_2d")
            self._c(lib.ocr_copy_2d(vp(cb), H, vp(L["bias"].data_ptr() + (d * 3 * H + 2 * H) * f), H, 1, H, sh), "ocr_copy_2d")
            # backward operands: wxcat [I, 6H] columns d*3H.. = [gates x-part | candidate x-part]; h-part rows as they are
            self._c(lib.ocr_copy_2d(vp(gk), 2 * H, vp(L["wxcat"].data_ptr() + d * 3 * H * f), 6 * H, I, 2 * H, sh), "ocr_copy_2d")
            self._c(lib.ocr_copy_2d(vp(ck), H, vp(L["wxcat"].data_ptr() + (d * 3 * H + 2 * H) * f), 6 * H, I, H, sh), "ocr_copy_2d")
            self._c(lib.ocr_copy_2d(vp(gk + I * 2 * H * f), 2 * H, vp(L["wg_rows"].data_ptr() + d * H * 2 * H * f), 2 * H, H, 2 * H, sh), "ocr_copy_2d")
            self._c(lib.ocr_copy_2d(vp(ck + I * H * f), H, vp(L["wc_rows"].data_ptr() + d * H * H * f), H, H, H, sh), "ocr_copy_2d")

    # ------------------------------------------------------------------ building blocks
    def _conv(self, x, w, bias, cout, relu):
        B, H, W, C = x.shape
        out = self._new(B, H, W, cout)
        self._c(self.lib.ocr_conv3x3_same(_lib.ptr(x), B, H, W, C, _lib.ptr(w), _lib.ptr(bias), cout, int(relu), _lib.ptr(out), self._sh()), "ocr_conv3x3_same")
        return out

    def _pool(self, x, ph, pw, s_h, s_w):
        B, H, W, C = x.shape
        Hp, Wp = (H - ph) // s_h + 1, (W - pw) // s_w + 1
        if Hp < 1 or Wp < 1:
            raise ValueError("image too small for the convolutional stack (need height 32, width >= 8)")
        out = self._new(B, Hp, Wp, C)
        self._c(self.lib.ocr_maxpool(_lib.ptr(x), B, H, W, C, ph, pw, s_h, s_w, _lib.ptr(out), self._sh()), "ocr_maxpool")
        return out

    def _wgrad(self, At, lda, Wt, ldw, D, ldd, batch_stride, M, N, R, shifts=(0,), a_rows=0, a_row=None):
        nb = len(shifts)
        need = ctypes.c_size_t(0)
        self._c(self.lib.ocr_gemm_wgrad_scratch_bytes(M, N, R, nb, ctypes.byref(need)), "ocr_gemm_wgrad_scratch_bytes")
        if self.wscratch is None or self.wscratch.numel() < need.value:
            self.wscratch = torch.empty(max(need.value, 1 << 24), dtype=torch.uint8, device=self.device)
        arr = (ctypes.c_int32 * nb)(*shifts)
        rows = (ctypes.c_int32 * nb)(*(a_row if a_row is not None else [0] * nb))
        self._c(self.lib.ocr_gemm_tf32_wgrad(At, lda, Wt, ldw, D, ldd, batch_stride, M, N, R, nb, arr, rows, a_rows, _lib.ptr(self.wscratch),
                                             self.wscratch.numel(), self._sh()), "ocr_gemm_tf32_wgrad")

    def _transposed(self, x2d, src_shift=0):
        """[R, C] -> [C, ld] with the long dimension contiguous (ld = R rounded up to 4); out[c, r] = x[r + src_shift, c]."""
        R, C = x2d.shape
        ld = (R + 3) // 4 * 4
        out = self._new(C, ld)
        self._c(self.lib.ocr_transpose(_lib.ptr(x2d), R, C, x2d.stride(0), _lib.ptr(out), ld, src_shift, self._sh()), "ocr_transpose")
        return out, ld

    def _planar(self, x, ncopies=1):
        B, H, W, C = x.shape
        R = B * (H + 2) * self.lib.ocr_planar_pad_pitch(W)      # a multiple of 4
        out = self._new(ncopies, C, R)
        self._c(self.lib.ocr_nhwc_to_planar_pad(_lib.ptr(x), B, H, W, C, _lib.ptr(out), R, ncopies, C * R, self._sh()), "ocr_nhwc_to_planar_pad")
        return out, R

    def _planar_blocked(self, x, ncopies=1):
        """Zero-ringed K-blocked planar copies [R/32][ncopies*C][32] (ocr_nhwc_to_planar_blocked)."""
        B, H, W, C = x.shape
        R = B * (H + 2) * self.lib.ocr_planar_pad_pitch32(W)     # a multiple of 32
        out = self._new(R // 32, ncopies * C, 32)
        self._c(self.lib.ocr_nhwc_to_planar_blocked(_lib.ptr(x), B, H, W, C, _lib.ptr(out), ncopies * C, 0, ncopies, self._sh()), "ocr_nhwc_to_planar_blocked")
        return out, R

    def _conv_wgrad(self, x, dy, name, xp=None):
        """d kernel [3,3,C,Cout] = sum over pixels of (3x3 patch of x) x dy.  Tap (i, j) is the zero-ringed planar copy of x
        shifted by j-1 pixels (three copies: TMA box origins must be 16-byte aligned) and by i-1 padded rows (a multiple of
        four elements): nine views, one launch."""
        B, H, W, C = x.shape
        Co = dy.shape[3]
        if getattr(self, "blocked_planar", True) and C % 4 == 0 and Co % 4 == 0:
            # K-blocked operands: a k-step's 128 bytes x C rows are one contiguous run (the channel planes of the plain planar copy
            # lie megabytes apart)
            Wp = self.lib.ocr_planar_pad_pitch32(W)
            xp, R = xp if xp is not None else self._planar_blocked(x, 3)
            dyp, _ = self._planar_blocked(dy, 1)
            shifts = [(i - 1) * Wp for i in range(3) for j in range(3)]
            a_row = [j * C for i in range(3) for j in range(3)]
            need = ctypes.c_size_t(0)
            self._c(self.lib.ocr_gemm_wgrad_scratch_bytes(C, Co, R, 9, ctypes.byref(need)), "ocr_gemm_wgrad_scratch_bytes")
            if self.wscratch is None or self.wscratch.numel() < need.value:
                self.wscratch = torch.empty(max(need.value, 1 << 24), dtype=torch.uint8, device=self.device)
            arr, rows = (ctypes.c_int32 * 9)(*shifts), (ctypes.c_int32 * 9)(*a_row)
            self._c(self.lib.ocr_gemm_tf32_wgrad_blocked(_lib.ptr(xp), 3 * C, _lib.ptr(dyp), Co, _lib.ptr(self.grads["convnet/%s/kernel" % name]), Co, C * Co,
                                                         C, Co, R, 9, arr, rows, _lib.ptr(self.wscratch), self.wscratch.numel(), self._sh()),
                    "ocr_gemm_tf32_wgrad_blocked")
            return
        Wp = self.lib.ocr_planar_pad_pitch(W)
        xp, R = xp if xp is not None else self._planar(x, 3)
        dyp, _ = self._planar(dy, 1)
        shifts = [(i - 1) * Wp for i in range(3) for j in range(3)]
        a_row = [j * C for i in range(3) for j in range(3)]
        self._wgrad(_lib.ptr(xp), R, _lib.ptr(dyp), R, _lib.ptr(self.grads["convnet/%s/kernel" % name]), Co, C * Co, C, Co, R, shifts, 3 * C, a_row)

    def _allreduce(self, t, async_op=False):
        import torch.distributed as dist
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, group=None if self.pg is True else self.pg, async_op=async_op)

    # ------------------------------------------------------------------ the step
    def _host_inputs(self, image, width, label):
        """Host side of a step: argument checks with TensorFlow's messages, sequence lengths (model.py:152-163), flat labels."""
        _lib.require_cuda(image)
        B, Hh, Ww, one = image.shape
        if one != 1:
            raise ValueError("image must be [B, H, W, 1]")
        x = image.contiguous()
        if x.dtype != torch.uint8:
            x = x.float()
        if torch.is_tensor(width) and width.is_cuda:
            seq_len_host = (torch.div(width.to(torch.int32) - 2, 2, rounding_mode="floor") - 2).tolist()
        else:
            seq_len_host = ((np.asarray(width, dtype=np.int64).reshape(-1) - 2) // 2 - 2).tolist()
        if len(seq_len_host) != B:
            raise ValueError("width must have one entry per image")
        T = (Ww - 2) // 2 - 2
        C = self.logits_w.shape[0]
        if isinstance(label, (list, tuple)) and len(label) and isinstance(label[0], (list, tuple, np.ndarray)) and not torch.is_tensor(label[0]):
            lengths = [len(l) for l in label]
            flat_host = torch.tensor([int(v) for l in label for v in l], dtype=torch.int32)
        else:
            _, _, lengths, flat_host = ctc._labels_to_flat(label, B, "cpu")
        ctc._validate_ctc(flat_host, lengths, seq_len_host, T, C, False)
        off = np.zeros(B + 1, np.int32)
        np.cumsum(lengths, out=off[1:])
        return x, seq_len_host, flat_host, off, (max(lengths) if lengths else 0)

    def forward_backward(self, image, width, label):
        """Forward in TRAIN mode + backward; fills self.grad (gradient of the MEAN CTC loss over this replica's batch).
        Returns the per-example losses [B] (device)."""
        x, seq_len_host, flat_host, off, max_len = self._host_inputs(image, width, label)
        dev = self.device
        seq_len = torch.tensor(seq_len_host, dtype=torch.int32).to(dev, non_blocking=True)
        flat = flat_host.to(dev) if flat_host.numel() else torch.zeros(1, dtype=torch.int32, device=dev)
        offsets = torch.from_numpy(off).to(dev)
        widths_dev = None
        if x.dtype == torch.uint8 and x.shape[1] == 31:      # raw mjsynth rows: training-side preprocessing on the device
            widths_dev = torch.as_tensor(np.asarray(width.cpu() if torch.is_tensor(width) else width, dtype=np.int32).reshape(-1)).to(dev, non_blocking=True)
        losses = self._backward_rnn(*self._forward(x, seq_len, flat, offsets, max_len, widths_dev))
        works = []
        if self.world > 1:   # the RNN + logits bucket is complete: its all-reduce runs behind the conv backward
            works.append(self._allreduce(self.grad[:self.n_rnn_floats], async_op=True))
        self._backward_conv()
        if self.world > 1:
            works.append(self._allreduce(self.grad[self.n_rnn_floats:], async_op=True))
            for w in works:
                w.wait()
        return losses

    def preprocess_train(self, image_u8, widths):
        """mjsynth._preprocess_image + the batcher's 0.0 padding (mjsynth.py:185-194,56,69) on the device:
        uint8 [B,31,W,1] rows + widths [B] int32 (device) -> float32 [B,32,W,1]."""
        B, Hin, Ww, _ = image_u8.shape
        out = self._new(B, Hin + 1, Ww, 1)
        self._c(self.lib.ocr_preprocess_train(_lib.ptr(image_u8), B, Hin, Ww, _lib.ptr(widths), _lib.ptr(out), self._sh()), "ocr_preprocess_train")
        return out

    def _forward(self, x, seq_len, flat, offsets, max_len, widths=None):
        lib, sh = self.lib, self._sh()
        if x.dtype == torch.uint8 and x.shape[1] == 31:
            if widths is None:
                raise ValueError("31-row uint8 images (raw mjsynth crops) need their widths for the training-side padding")
            x = self.preprocess_train(x, widths)
        B, Hh, Ww, one = x.shape
        is_u8 = x.dtype == torch.uint8
        P, G = self.params, self.grads
        scr = _lib.ptr(self.scratch)
        saved = {}
        # ---------------- forward: convnet_layers(mode=TRAIN)
        w1, b1 = P["convnet/conv1/kernel"], P["convnet/conv1/bias"]
        a = self._new(B, Hh - 2, Ww - 2, w1.shape[-1])
        self._c(lib.ocr_conv1_3x3_valid(_lib.ptr(x), int(is_u8), B, Hh, Ww, _lib.ptr(w1), _lib.ptr(b1), w1.shape[-1], _lib.ptr(a), sh), "ocr_conv1_3x3_valid")
        saved["conv1"] = dict(out=a)
        names = [lp[3] for lp in LAYER_PARAMS]
        pre_pooled = None            # the pool in front of this layer, already taken by the previous layer's batch-norm pass
        pre_arg = None               # ... and the arguments of its maxima (ocr_bn_relu_apply_pool_arg)
        for (filters, k, padding, name, bn) in LAYER_PARAMS[1:]:
            S = {}
            ph, pw, s_h, s_w = _POOL_BEFORE[name]
            if (ph, pw, s_h, s_w) != (1, 1, 1, 1):
                S["pool_in"], S["pool"] = a, (ph, pw, s_h, s_w)
                a = pre_pooled if pre_pooled is not None else self._pool(a, ph, pw, s_h, s_w)
                if pre_arg is not None:
                    saved[pre_arg[0]]["pool_arg"] = pre_arg[1:]     # the layer in front back-propagates through this pool itself
            pre_pooled = pre_arg = None
            S["x"] = a
            wf, _ = self.conv_w[name]
            bias = P["convnet/%s/bias" % name]
            if not bn:
                a = self._conv(a, wf, bias, filters, relu=True)
                S["out"] = a
            else:
                y = self._conv(a, wf, bias, filters, relu=False)
                rows = y.numel() // filters
                q = "convnet/%s/batch_norm/" % name
                sums = torch.empty(2 * filters, dtype=torch.float64, device=self.device)
                self._c(lib.ocr_bn_batch_sums(_lib.ptr(y), rows, filters, _lib.ptr(sums), sh), "ocr_bn_batch_sums")
                n_stat = rows
                if self.sync_bn:
                    self._allreduce(sums)
                    n_stat = rows * self.world
                mean, inv_std = self._new(filters), self._new(filters)
                self._c(lib.ocr_bn_finalize(_lib.ptr(sums), n_stat, filters, BN_EPS, BN_MOMENTUM, _lib.ptr(mean), _lib.ptr(inv_std),
                                            _lib.ptr(self.stats[q + "moving_mean"]), _lib.ptr(self.stats[q + "moving_variance"]), sh), "ocr_bn_finalize")
                nxt = names.index(name) + 1
                npool = _POOL_BEFORE[names[nxt]] if nxt < len(names) else (1, 1, 1, 1)
                Bn, Hn, Wn, _ = y.shape
                fused_pool = npool[:3] == (2, 2, 2) and npool[3] in (1, 2) and Hn >= 2 and Wn >= 2
                if fused_pool and self.pool_arg and 256 % (filters // 4) == 0 and Bn * Hn * Wn < 2 ** 31 - 1:
                    # normalise + ReLU + the pool in front of the next layer in one pass over y (model.py:105-116); the full-size
                    # activation is never written: the arguments of the maxima (a byte per four pooled channels) carry the pool's
                    # gradient into this layer's batch-norm backward
                    Hp_, Wp_ = (Hn - 2) // 2 + 1, (Wn - 2) // npool[3] + 1
                    pre_pooled = self._new(Bn, Hp_, Wp_, filters)
                    arg = torch.empty(Bn * Hp_ * Wp_ * (filters // 4), dtype=torch.uint8, device=self.device)
                    self._c(lib.ocr_bn_relu_apply_pool_arg(_lib.ptr(y), Bn, Hn, Wn, filters, _lib.ptr(mean), _lib.ptr(inv_std), _lib.ptr(P[q + "gamma"]),
                                                           _lib.ptr(P[q + "beta"]), npool[3], _lib.ptr(pre_pooled), _lib.ptr(arg), sh), "ocr_bn_relu_apply_pool_arg")
                    pre_arg = (name, arg, (Bn, Hn, Wn, filters), npool[3])
                    a = None
                elif fused_pool:
                    a = self._new(*y.shape)
                    pre_pooled = self._new(Bn, (Hn - 2) // 2 + 1, (Wn - 2) // npool[3] + 1, filters)
                    self._c(lib.ocr_bn_relu_apply_pool(_lib.ptr(y), Bn, Hn, Wn, filters, _lib.ptr(mean), _lib.ptr(inv_std), _lib.ptr(P[q + "gamma"]),
                                                       _lib.ptr(P[q + "beta"]), _lib.ptr(a), npool[3], _lib.ptr(pre_pooled), sh), "ocr_bn_relu_apply_pool")
                else:
                    a = self._new(*y.shape)
                    self._c(lib.ocr_bn_relu_apply(_lib.ptr(y), rows, filters, _lib.ptr(mean), _lib.ptr(inv_std), _lib.ptr(P[q + "gamma"]),
                                                  _lib.ptr(P[q + "beta"]), _lib.ptr(a), sh), "ocr_bn_relu_apply")
                S.update(y=y, mean=mean, inv_std=inv_std, out=a, n_stat=n_stat)
            saved[name] = S
        if self.overlap and self.early_planar:
            # the planar copies of the layer INPUTS that the weight gradients contract over depend on the forward pass only: they
            # are made on the side stream now, beside the recurrent layers (latency-bound, bandwidth idle), instead of at the end
            # of the step, where the side stream -- planar copies + the split-K contractions of conv4 .. conv2 -- is what the
            # optimizer waits for (0.6 ms of copies at B = 256)
            with self._on_side(*[saved[n]["x"] for n in names[1:]]):
                for n in names[1:]:
                    saved[n]["xp"] = (self._planar_blocked if self.blocked_planar else self._planar)(saved[n]["x"], 3)
        Bn, Hn, Wn, Cn = a.shape
        seq = self._new(Wn, Bn, Cn)
        self._c(lib.ocr_rows_max_to_seq(_lib.ptr(a), Bn, Hn, Wn, Cn, _lib.ptr(seq), sh), "ocr_rows_max_to_seq")
        T = Wn
        # ---------------- forward: rnn_layers
        need = ctypes.c_size_t(0)
        Hmax = max(self.rnn_sizes)
        if self.cell_type == "lstm":
            self._c(lib.ocr_birnn_lstm_train_workspace_bytes(T, B, Hmax, ctypes.byref(need)), "ocr_birnn_lstm_train_workspace_bytes")
        else:
            self._c(lib.ocr_birnn_gru_train_workspace_bytes(T, B, Hmax, ctypes.byref(need)), "ocr_birnn_gru_train_workspace_bytes")
        ws = torch.empty(need.value, dtype=torch.uint8, device=self.device)
        rnn_saved = []
        xin = seq
        for scope, L in zip(("bdrnn1", "bdrnn2"), self.rnn_w):
            I, H = L["I"], L["H"]
            if self.cell_type == "gru":
                out, act, rh = self._new(T, B, 2 * H), self._new(T * B, 6 * H), self._new(T, B, 2 * H)
                self._c(lib.ocr_birnn_gru_train_fwd(_lib.ptr(xin), T, B, I, H, _lib.ptr(seq_len), _lib.ptr(L["wx"]), _lib.ptr(L["whg"]), _lib.ptr(L["whc"]),
                                                    _lib.ptr(L["bias"]), _lib.ptr(out), _lib.ptr(act), _lib.ptr(rh), _lib.ptr(ws), need.value, sh),
                        "ocr_birnn_gru_train_fwd")
                rnn_saved.append(dict(x=xin, out=out, gates=act, rh=rh))
                xin = out
                continue
            out, gates, cs = self._new(T, B, 2 * H), self._new(T * B, 8 * H), self._new(T, B, 2 * H)
            o = self.offsets["rnn/%s/fw/lstm_cell/bias" % scope]
            bias8 = self.theta[o:o + 8 * H]     # fw | bw biases are adjacent in the flat buffer
            self._c(lib.ocr_birnn_lstm_train_fwd(_lib.ptr(xin), T, B, I, H, _lib.ptr(seq_len), _lib.ptr(L["wx"]), _lib.ptr(L["wh"]), _lib.ptr(bias8),
                                                 _lib.ptr(out), _lib.ptr(gates), _lib.ptr(cs), _lib.ptr(ws), need.value, sh), "ocr_birnn_lstm_train_fwd")
            rnn_saved.append(dict(x=xin, out=out, gates=gates, cs=cs))
            xin = out
        F = xin.shape[2]
        C = self.logits_w.shape[0]
        R = T * B
        logits = self._new(T, B, C)
        self._c(lib.ocr_gemm_tf32(_lib.ptr(xin), F, _lib.ptr(self.logits_w), F, _lib.ptr(P["rnn/logits/bias"]), _lib.ptr(logits), C, R, C, F, 1, sh), "ocr_gemm_tf32")
        # ---------------- ctc_loss_layer: mean over the batch (model.py:224-229); the kernel returns d mean / d logits
        losses, dlog, _ = ctc.ctc_loss_raw(logits, flat, offsets, seq_len, max_len, want_grad=True, grad_scale=1.0 / B)
        self.last_logits, self.last_seq_len = logits, seq_len
        return losses, logits, dlog, rnn_saved, saved, seq_len, ws, need, x

    def _backward_rnn(self, losses, logits, dlog, rnn_saved, saved, seq_len, ws, need, x):
        lib, sh = self.lib, self._sh()
        P, G = self.params, self.grads
        scr = _lib.ptr(self.scratch)
        T, B, C = logits.shape
        R = T * B
        F = rnn_saved[-1]["out"].shape[2]
        # ---------------- backward: logits layer (dense + ReLU, model.py:216-220)
        self._c(lib.ocr_relu_bwd(_lib.ptr(logits), _lib.ptr(dlog), dlog.numel(), _lib.ptr(dlog), sh), "ocr_relu_bwd")
        dz = dlog.view(R, C)
        scr_s = _lib.ptr(self.scratch_side)
        with self._on_side(dz, rnn_saved[-1]["out"]):
            self._c(lib.ocr_colsum(_lib.ptr(dz), R, C, C, _lib.ptr(G["rnn/logits/bias"]), scr_s, self._sh()), "ocr_colsum")
            dzT, ldz = self._transposed(dz)
            outT, ldo = self._transposed(rnn_saved[-1]["out"].view(R, F))
            self._wgrad(_lib.ptr(outT), ldo, _lib.ptr(dzT), ldz, _lib.ptr(G["rnn/logits/kernel"]), C, 0, F, C, R, [0])
            del dzT
        dout = self._new(T, B, F)
        self._c(lib.ocr_gemm_tf32(_lib.ptr(dz), C, _lib.ptr(P["rnn/logits/kernel"]), C, None, _lib.ptr(dout), F, R, F, C, 0, sh), "ocr_gemm_tf32")
        # ---------------- backward: recurrent layers (BPTT), last layer first
        for li in (1, 0):
            scope = ("bdrnn1", "bdrnn2")[li]
            L, S = self.rnn_w[li], rnn_saved[li]
            I, H = L["I"], L["H"]
            if self.cell_type == "gru":
                dout, outT, ldo = self._backward_gru_layer(scope, L, S, dout, outT, ldo, T, B, seq_len, ws, need)
                continue
            self._c(lib.ocr_birnn_lstm_bwd(_lib.ptr(dout), T, B, H, _lib.ptr(seq_len), _lib.ptr(S["gates"]), _lib.ptr(S["cs"]), _lib.ptr(L["wh_rows"]),
                                           _lib.ptr(ws), need.value, sh), "ocr_birnn_lstm_bwd")
            dG = S["gates"]                                   # [R, 8H] gradient of the gate pre-activations
            with self._on_side(dG, S["x"], S["out"]):
                o = self.offsets["rnn/%s/fw/lstm_cell/bias" % scope]
                self._c(lib.ocr_colsum(_lib.ptr(dG), R, 8 * H, 8 * H, _lib.ptr(self.grad[o:o + 8 * H]), scr_s, self._sh()), "ocr_colsum")
                dGT, ldg = self._transposed(dG)
                xT, ldx = self._transposed(S["x"].reshape(R, I))
                for d, dn in enumerate(("fw", "bw")):
                    gk = G["rnn/%s/%s/lstm_cell/kernel" % (scope, dn)]                 # [I+H, 4H]
                    Wt = ctypes.c_void_p(dGT.data_ptr() + d * 4 * H * ldg * 4)
                    self._wgrad(_lib.ptr(xT), ldx, Wt, ldg, _lib.ptr(gk), 4 * H, 0, I, 4 * H, R, [0])
                    # h_{prev}: the layer's own output one frame earlier (forward) / later (backward direction)
                    shift = -B if d == 0 else B
                    if B % 4 == 0:      # a frame shift of the transposed output is an aligned TMA coordinate offset
                        At, lda, shifts = ctypes.c_void_p(outT.data_ptr() + d * H * ldo * 4), ldo, [shift]
                    else:
                        hprevT, lda = self._transposed(S["out"].view(R, 2 * H)[:, d * H:(d + 1) * H], shift)
                        At, shifts = _lib.ptr(hprevT), [0]
                    self._wgrad(At, lda, Wt, ldg, ctypes.c_void_p(gk.data_ptr() + I * 4 * H * 4), 4 * H, 0, H, 4 * H, R, shifts)
                outT, ldo = xT, ldx          # the input of layer 2 is the output of layer 1
                del dGT
            dx = self._new(T, B, I)
            self._c(lib.ocr_gemm_tf32(_lib.ptr(dG), 8 * H, _lib.ptr(L["wxcat"]), 8 * H, None, _lib.ptr(dx), I, R, I, 8 * H, 0, sh), "ocr_gemm_tf32")
            dout = dx
            S.clear()
        del outT
        self._join_side()     # the RNN + logits bucket of the flat gradient is complete on the current stream
        self._conv_state = (saved, dout, x)
        return losses

    def _backward_gru_layer(self, scope, L, S, dout, outT, ldo, T, B, seq_len, ws, need):
        """BPTT of one bidirectional GRU layer, then d kernel / d bias / d input as contractions.  Returns (d input, input^T, ld)."""
        lib, sh, G = self.lib, self._sh(), self.grads
        scr = _lib.ptr(self.scratch)
        I, H = L["I"], L["H"]
        R = T * B
        vp = ctypes.c_void_p
        self._c(lib.ocr_birnn_gru_bwd(_lib.ptr(dout), T, B, H, _lib.ptr(seq_len), _lib.ptr(S["gates"]), _lib.ptr(S["out"]), _lib.ptr(L["wg_rows"]),
                                      _lib.ptr(L["wc_rows"]), _lib.ptr(ws), need.value, sh), "ocr_birnn_gru_bwd")
        dA = S["gates"]                                      # [R, 6H]: per direction d z_r | d z_u | d z_c
        dx = self._new(T, B, I)
        self._c(lib.ocr_gemm_tf32(_lib.ptr(dA), 6 * H, _lib.ptr(L["wxcat"]), 6 * H, None, _lib.ptr(dx), I, R, I, 6 * H, 0, sh), "ocr_gemm_tf32")
        ctx = self._on_side(dA, S["x"], S["out"], S["rh"])
        ctx.__enter__()
        scr = _lib.ptr(self.scratch_side)
        sh = self._sh()
        dAT, lda_ = self._transposed(dA)
        xT, ldx = self._transposed(S["x"].reshape(R, I))
        rhT, ldr = self._transposed(S["rh"].view(R, 2 * H))
        for d, dn in enumerate(("fw", "bw")):
            q = "rnn/%s/%s/gru_cell/" % (scope, dn)
            gk, ck = G[q + "gates/kernel"], G[q + "candidate/kernel"]
            col = d * 3 * H
            self._c(lib.ocr_colsum(vp(dA.data_ptr() + col * 4), R, 2 * H, 6 * H, _lib.ptr(G[q + "gates/bias"]), scr, sh), "ocr_colsum")
            self._c(lib.ocr_colsum(vp(dA.data_ptr() + (col + 2 * H) * 4), R, H, 6 * H, _lib.ptr(G[q + "candidate/bias"]), scr, sh), "ocr_colsum")
            Wg = vp(dAT.data_ptr() + col * lda_ * 4)                 # rows d z_r, d z_u of this direction
            Wc = vp(dAT.data_ptr() + (col + 2 * H) * lda_ * 4)       # rows d z_c
            self._wgrad(_lib.ptr(xT), ldx, Wg, lda_, _lib.ptr(gk), 2 * H, 0, I, 2 * H, R, [0])
            self._wgrad(_lib.ptr(xT), ldx, Wc, lda_, _lib.ptr(ck), H, 0, I, H, R, [0])
            shift = -B if d == 0 else B                              # h_prev = the layer's output one frame earlier / later
            if B % 4 == 0:
                At, lda, shifts = vp(outT.data_ptr() + d * H * ldo * 4), ldo, [shift]
            else:
                hprevT, lda = self._transposed(S["out"].view(R, 2 * H)[:, d * H:(d + 1) * H], shift)
                At, shifts = _lib.ptr(hprevT), [0]
            self._wgrad(At, lda, Wg, lda_, vp(gk.data_ptr() + I * 2 * H * 4), 2 * H, 0, H, 2 * H, R, shifts)
            self._wgrad(vp(rhT.data_ptr() + d * H * ldr * 4), ldr, Wc, lda_, vp(ck.data_ptr() + I * H * 4), H, 0, H, H, R, [0])
        ctx.__exit__(None, None, None)
        S.clear()
        return dx, xT, ldx

    def _backward_conv(self):
        lib, sh = self.lib, self._sh()
        P, G = self.params, self.grads
        scr = _lib.ptr(self.scratch)
        saved, dout, x = self._conv_state
        self._conv_state = None
        is_u8 = x.dtype == torch.uint8
        B, Hh, Ww, _ = x.shape
        S = saved["conv8"]
        a8 = S["out"]
        Bn, Hn, Wn, Cn = a8.shape
        da = self._new(*a8.shape)
        self._c(lib.ocr_rows_max_to_seq_bwd(_lib.ptr(a8), _lib.ptr(dout), Bn, Hn, Wn, Cn, _lib.ptr(da), sh), "ocr_rows_max_to_seq_bwd")
        for (filters, k, padding, name, bn) in reversed(LAYER_PARAMS[1:]):
            S = saved[name]
            xin = S["x"]
            rows = da.numel() // filters
            if bn and "pool_arg" in S:
                # the gradient w.r.t. the activation is MaxPoolGrad(dpool): formed from the arguments of the maxima inside the two
                # batch-norm backward passes (dpool = the input gradient of the layer behind the pool, held in `da`)
                q = "convnet/%s/batch_norm/" % name
                arg, (Bn, Hn, Wn, _), sw_ = S["pool_arg"]
                rows = Bn * Hn * Wn
                sums = torch.empty(2 * filters, dtype=torch.float64, device=self.device)
                args = (_lib.ptr(S["mean"]), _lib.ptr(S["inv_std"]), _lib.ptr(P[q + "gamma"]), _lib.ptr(P[q + "beta"]))
                self._c(lib.ocr_bn_relu_bwd_sums_pool(_lib.ptr(S["y"]), _lib.ptr(da), _lib.ptr(arg), Bn, Hn, Wn, filters, sw_, *args, _lib.ptr(sums),
                                                      _lib.ptr(G[q + "gamma"]), _lib.ptr(G[q + "beta"]), sh), "ocr_bn_relu_bwd_sums_pool")
                if self.sync_bn:
                    self._allreduce(sums)
                dy = self._new(Bn, Hn, Wn, filters)
                self._c(lib.ocr_bn_relu_bwd_apply_bias_pool(_lib.ptr(S["y"]), _lib.ptr(da), _lib.ptr(arg), Bn, Hn, Wn, filters, sw_, S["n_stat"], *args,
                                                            _lib.ptr(sums), _lib.ptr(dy), _lib.ptr(G["convnet/%s/bias" % name]), scr, sh),
                        "ocr_bn_relu_bwd_apply_bias_pool")
            elif bn:
                q = "convnet/%s/batch_norm/" % name
                sums = torch.empty(2 * filters, dtype=torch.float64, device=self.device)
                args = (_lib.ptr(S["mean"]), _lib.ptr(S["inv_std"]), _lib.ptr(P[q + "gamma"]), _lib.ptr(P[q + "beta"]))
                self._c(lib.ocr_bn_relu_bwd_sums(_lib.ptr(S["y"]), _lib.ptr(da), rows, filters, *args, _lib.ptr(sums), _lib.ptr(G[q + "gamma"]),
                                                 _lib.ptr(G[q + "beta"]), sh), "ocr_bn_relu_bwd_sums")
                if self.sync_bn:
                    self._allreduce(sums)
                dy = da
                if 256 % (filters // 4) == 0:      # the bias gradient (column sums of dy) comes out of the same pass
                    self._c(lib.ocr_bn_relu_bwd_apply_bias(_lib.ptr(S["y"]), _lib.ptr(da), rows, S["n_stat"], filters, *args, _lib.ptr(sums), _lib.ptr(dy),
                                                           _lib.ptr(G["convnet/%s/bias" % name]), scr, sh), "ocr_bn_relu_bwd_apply_bias")
                else:
                    self._c(lib.ocr_bn_relu_bwd_apply(_lib.ptr(S["y"]), _lib.ptr(da), rows, S["n_stat"], filters, *args, _lib.ptr(sums), _lib.ptr(dy), sh),
                            "ocr_bn_relu_bwd_apply")
                    self._c(lib.ocr_colsum(_lib.ptr(dy), rows, filters, filters, _lib.ptr(G["convnet/%s/bias" % name]), scr, sh), "ocr_colsum")
            else:
                dy = da
                self._c(lib.ocr_relu_bwd_bias(_lib.ptr(S["out"]), _lib.ptr(da), rows, filters, _lib.ptr(dy), _lib.ptr(G["convnet/%s/bias" % name]), scr, sh),
                        "ocr_relu_bwd_bias")
            with self._on_side(xin, dy):
                self._conv_wgrad(xin, dy, name, S.get("xp"))
            _, wd = self.conv_w[name]
            dxin = self._conv(dy, wd, self.zero_bias, xin.shape[3], relu=False)
            if "pool" in S and S["pool_in"] is None:
                da = dxin                  # the layer in front takes the pool's gradient from its stored arguments of the maxima
            elif "pool" in S:
                ph, pw, s_h, s_w = S["pool"]
                pin = S["pool_in"]
                da = self._new(*pin.shape)
                Bp, Hp_, Wp_, Cp = pin.shape
                self._c(lib.ocr_maxpool_bwd(_lib.ptr(pin), _lib.ptr(dxin), Bp, Hp_, Wp_, Cp, ph, pw, s_h, s_w, _lib.ptr(da), sh), "ocr_maxpool_bwd")
            else:
                da = dxin
            S.clear()
        a1 = saved["conv1"]["out"]
        rows = a1.numel() // a1.shape[3]
        self._c(lib.ocr_relu_bwd_bias(_lib.ptr(a1), _lib.ptr(da), rows, a1.shape[3], _lib.ptr(da), _lib.ptr(G["convnet/conv1/bias"]), scr, sh), "ocr_relu_bwd_bias")
        self._c(lib.ocr_conv1_wgrad(_lib.ptr(x), int(is_u8), B, Hh, Ww, _lib.ptr(da), a1.shape[3], _lib.ptr(G["convnet/conv1/kernel"]), scr, sh), "ocr_conv1_wgrad")
        self._join_side()

    def _lr_t(self):
        lr = learning_rate(self.global_step, **self.hp)
        t = self.global_step + 1
        return lr * math.sqrt(1.0 - self.beta2 ** t) / (1.0 - self.beta1 ** t)

    def apply_gradients(self, lr_t_device=None):
        """AdamOptimizer.apply_gradients with the decayed learning rate of the CURRENT global step, then global_step += 1."""
        vp = ctypes.c_void_p
        for (o, e) in self.adam_runs:      # one run = the whole buffer unless tune_scope froze some variables
            self._c(self.lib.ocr_adam_step(vp(self.theta.data_ptr() + 4 * o), vp(self.grad.data_ptr() + 4 * o), vp(self.adam_m.data_ptr() + 4 * o),
                                           vp(self.adam_v.data_ptr() + 4 * o), e - o, self._lr_t(), _lib.ptr(lr_t_device), self.beta1, self.beta2,
                                           self.epsilon, 1.0 / self.world, self._sh()), "ocr_adam_step")
        self.global_step += 1
        self.derive_layouts()

    def train_step(self, image, width, label):
        """[step_loss, step] = sess.run([train_op, global_step])  (train.py:196): returns the mean CTC loss (device scalar)."""
        losses = self.forward_backward(image, width, label)
        self.apply_gradients()
        return losses.mean()

    # ------------------------------------------------------------------ the step as CUDA graphs
    def capture(self, batch_size, width, height=32, max_label_len=None):
        """Record the step for a fixed input shape [batch_size, height, width, 1] uint8 as CUDA graphs (the ~1100 launches of
        a step otherwise cost more host time than device time at small per-GPU batches).  One replica: one graph.  Data
        parallel: three graphs (forward + recurrent backward | conv backward | Adam) with the two NCCL bucket all-reduces
        issued between them, so the first overlaps the conv backward.  Use train_step_captured afterwards."""
        dev = self.device
        T = (width - 2) // 2 - 2
        Lmax = int(max_label_len) if max_label_len else min(T, 127)
        g = dict(B=batch_size, W=width, H=height, Lmax=Lmax)
        g["image"] = torch.zeros((batch_size, height, width, 1), dtype=torch.uint8, device=dev)
        g["seq_len"] = torch.full((batch_size,), T, dtype=torch.int32, device=dev)
        g["flat"] = torch.zeros(batch_size * Lmax, dtype=torch.int32, device=dev)
        g["offsets"] = torch.arange(batch_size + 1, dtype=torch.int32, device=dev)
        g["lr_t"] = torch.zeros(1, dtype=torch.float32, device=dev)
        g["widths"] = torch.full((batch_size,), width, dtype=torch.int32, device=dev)
        g["h_widths"] = torch.zeros(batch_size, dtype=torch.int32).pin_memory()
        # pinned staging for the small per-step host values
        g["h_seq_len"] = torch.zeros(batch_size, dtype=torch.int32).pin_memory()
        g["h_flat"] = torch.zeros(batch_size * Lmax, dtype=torch.int32).pin_memory()
        g["h_offsets"] = torch.zeros(batch_size + 1, dtype=torch.int32).pin_memory()
        g["h_lr_t"] = torch.zeros(1, dtype=torch.float32).pin_memory()
        # warm-up on a side stream (lazy scratch allocations, cudaFuncSetAttribute calls) with a feasible dummy batch
        g["flat"][:batch_size] = 0
        step0 = self.global_step
        backup = [t.clone() for t in (self.theta, self.adam_m, self.adam_v)] + [v.clone() for v in self.stats.values()]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self._backward_rnn(*self._forward(g["image"], g["seq_len"], g["flat"], g["offsets"], Lmax, g["widths"]))
            self._backward_conv()
            self.apply_gradients(g["lr_t"])
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        pool = torch.cuda.graph_pool_handle()
        graphs = []

        def rec(fn):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, pool=pool):
                out = fn()
            graphs.append(gr)
            return out

        if self.world == 1:
            def whole():
                losses = self._backward_rnn(*self._forward(g["image"], g["seq_len"], g["flat"], g["offsets"], Lmax, g["widths"]))
                self._backward_conv()
                self.apply_gradients(g["lr_t"])
                return losses
            g["losses"] = rec(whole)
        else:
            g["losses"] = rec(lambda: self._backward_rnn(*self._forward(g["image"], g["seq_len"], g["flat"], g["offsets"], Lmax, g["widths"])))
            rec(self._backward_conv)
            rec(lambda: self.apply_gradients(g["lr_t"]))
        # undo the warm-up / capture side effects on the variables (captures do not execute, the warm-up did)
        for dst, src in zip([self.theta, self.adam_m, self.adam_v] + list(self.stats.values()), backup):
            dst.copy_(src)
        self.global_step = step0
        self.derive_layouts()
        g["graphs"] = graphs
        self._graph = g
        return self

    def train_step_captured(self, image, width, label):
        """train_step through the captured graphs.  image: uint8 [B,H,W,1] (host pinned or device); returns the per-example
        losses tensor (device, overwritten by the next step)."""
        g = self._graph
        if tuple(image.shape) != (g["B"], g["H"], g["W"], 1) or image.dtype != torch.uint8:
            raise ValueError("captured for uint8 images of shape %s" % ((g["B"], g["H"], g["W"], 1),))
        T = (g["W"] - 2) // 2 - 2
        C = self.logits_w.shape[0]
        seq_len_host = ((np.asarray(width, dtype=np.int64).reshape(-1) - 2) // 2 - 2)
        lengths = [len(l) for l in label]
        if len(lengths) != g["B"] or len(seq_len_host) != g["B"]:
            raise ValueError("width / label must have one entry per image")
        if max(lengths, default=0) > g["Lmax"]:
            raise ValueError("label longer than the captured maximum %d" % g["Lmax"])
        flat_host = torch.tensor([int(v) for l in label for v in l], dtype=torch.int32)
        ctc._validate_ctc(flat_host, lengths, seq_len_host.tolist(), T, C, False)
        # The pinned staging buffers are rewritten every call: wait until the previous call's host-to-device copies have
        # executed, or a host running ahead of the GPU would hand step N the labels / lengths / step size of step N+1.
        if g.get("staged") is not None:
            g["staged"].synchronize()
        g["h_seq_len"].copy_(torch.from_numpy(seq_len_host.astype(np.int32)))
        g["h_flat"][:flat_host.numel()] = flat_host
        g["h_offsets"][0] = 0
        g["h_offsets"][1:] = torch.from_numpy(np.cumsum(lengths).astype(np.int32))
        g["h_lr_t"][0] = self._lr_t()
        g["h_widths"].copy_(torch.from_numpy(np.asarray(width, dtype=np.int32).reshape(-1)))
        g["widths"].copy_(g["h_widths"], non_blocking=True)
        g["image"].copy_(image, non_blocking=True)
        g["seq_len"].copy_(g["h_seq_len"], non_blocking=True)
        g["flat"].copy_(g["h_flat"], non_blocking=True)
        g["offsets"].copy_(g["h_offsets"], non_blocking=True)
        g["lr_t"].copy_(g["h_lr_t"], non_blocking=True)
        if g.get("staged") is None:
            g["staged"] = torch.cuda.Event()
        g["staged"].record(torch.cuda.current_stream(self.device))
        gr = g["graphs"]
        if self.world == 1:
            gr[0].replay()
        else:
            gr[0].replay()
            if getattr(self, "skip_allreduce", False):   # timing aid (bench.py: exposed time of the collectives)
                gr[1].replay()
            else:
                w1 = self._allreduce(self.grad[:self.n_rnn_floats], async_op=True)
                gr[1].replay()
                w2 = self._allreduce(self.grad[self.n_rnn_floats:], async_op=True)
                w1.wait()
                w2.wait()
            gr[2].replay()
        self.global_step += 1
        return g["losses"]
