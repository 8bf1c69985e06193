"""CTC ops of the recognizer's hot path on B200: the stand-ins for the TensorFlow ops the
reference calls.

  ctc_loss                 <- tf.nn.ctc_loss                 (src/weinman/model.py:226)
  ctc_greedy_decoder       <- tf.nn.ctc_greedy_decoder       (src/weinman/validate.py:86)
  ctc_beam_search_decoder  <- tf.nn.ctc_beam_search_decoder  (src/weinman/test.py:84, client.py:227)
  edit_distance            <- tf.edit_distance               (src/weinman/test.py:90)
  sparse_tensor_to_dense   <- tf.sparse_tensor_to_dense      (src/weinman/validate.py:91)

Same argument order, shapes and dtypes as the TensorFlow ops (logits time-major [T,B,C] float32,
blank = C-1, sparse labels as (indices[N,2] int64, values[N] int32, dense_shape[2])); tensors are
torch CUDA tensors.  All arithmetic runs in the hand-written sm_100a kernels of libocr_b200.so.
"""
import collections
import ctypes

import numpy as np
import torch

from . import _lib

SparseTensor = collections.namedtuple("SparseTensor", ["indices", "values", "dense_shape"])


# ----------------------------------------------------------------------------- label handling
def _labels_to_flat(labels, batch_size, device):
    """Accepts a SparseTensor-like triple, (values, lengths), or a list of int sequences.
    Returns (flat int32 [N] on device, offsets int32 [B+1] on device, lengths list, host flat tensor).
    All host work is vectorised (numpy): this sits on the end-to-end path of every step."""
    if isinstance(labels, (list, tuple)) and len(labels) == 3 and torch.is_tensor(labels[0]) and labels[0].dim() == 2:
        indices, values, _ = labels
        rows = indices[:, 0].to("cpu", torch.int64).numpy()
        if rows.size > 1 and bool((rows[1:] < rows[:-1]).any()):
            raise ValueError("labels.indices must be ordered by batch (row-major), as tf.nn.ctc_loss requires")
        lengths = np.bincount(rows, minlength=batch_size) if rows.size else np.zeros(batch_size, np.int64)
        flat_np = values.to("cpu", torch.int32).numpy()
    elif isinstance(labels, (list, tuple)) and len(labels) == 2 and torch.is_tensor(labels[0]) and torch.is_tensor(labels[1]):
        flat_np = labels[0].to("cpu", torch.int32).numpy()
        lengths = labels[1].to("cpu", torch.int64).numpy()
    else:
        lengths = np.fromiter((len(l) for l in labels), dtype=np.int64, count=len(labels))
        flat_np = np.fromiter((v for l in labels for v in l), dtype=np.int32, count=int(lengths.sum()))
    if len(lengths) != batch_size:
        raise ValueError("labels describe %d examples but logits have batch %d" % (len(lengths), batch_size))
    off = np.zeros(batch_size + 1, np.int32)
    np.cumsum(lengths, out=off[1:])
    if flat_np.size != int(off[-1]):
        raise ValueError("label values/lengths mismatch")
    flat_host = torch.from_numpy(np.ascontiguousarray(flat_np))
    if str(device) == "cpu":
        return flat_host, torch.from_numpy(off), lengths.tolist(), flat_host
    offsets = torch.from_numpy(off).to(device, non_blocking=True)
    flat = flat_host.to(device, non_blocking=True) if flat_np.size else torch.zeros(1, dtype=torch.int32, device=device)
    return flat, offsets, lengths.tolist(), flat_host


def _validate_ctc(flat_host, lengths, seq_len_host, T, C, ignore_longer_outputs_than_inputs):
    """Host-side argument validation with TensorFlow's error texts; vectorised."""
    sl = np.asarray(seq_len_host, dtype=np.int64).reshape(-1)
    if sl.size and (int(sl.max()) > T or int(sl.min()) < 0):
        raise ValueError("sequence_length(b) <= %d required (max_time)" % T)
    flat = np.asarray(flat_host, dtype=np.int64).reshape(-1)
    if flat.size and (int(flat.max()) >= C - 1 or int(flat.min()) < 0):
        raise ValueError("Saw a non-null label (index >= num_classes - 1) following a null label, or a label "
                         "outside [0, %d): labels must be < num_classes - 1 = %d" % (C - 1, C - 1))
    if ignore_longer_outputs_than_inputs or not flat.size:
        return
    ln = np.asarray(lengths, dtype=np.int64).reshape(-1)
    # a label of length L needs at most 2L - 1 frames (every neighbour repeated): nothing to count when all have that many
    if bool(((sl >= 2 * ln) | (sl <= 0)).all()):
        return
    off = np.zeros(ln.size + 1, np.int64)
    np.cumsum(ln, out=off[1:])
    rep = np.zeros(flat.size + 1, np.int64)           # rep[i+1] = repeats among flat[:i+1] that are not example starts
    eq = np.zeros(flat.size, np.int64)
    eq[1:] = flat[1:] == flat[:-1]
    eq[off[:-1][ln > 0]] = 0
    np.cumsum(eq, out=rep[1:])
    need = ln + rep[off[1:]] - rep[off[:-1]]
    bad = np.nonzero((sl > 0) & (need > sl))[0]
    if bad.size:
        b = int(bad[0])
        raise ValueError("Not enough time for target transition sequence (required: %d, available: %d)%d"
                         "You can turn this error into a warning by using the flag "
                         "ignore_longer_outputs_than_inputs" % (int(need[b]), int(sl[b]), b))


def ctc_loss_raw(logits, flat, offsets, seq_len, max_label_len, want_grad=True, grad_scale=1.0):
    """Thin call into ocr_ctc_loss: returns (loss[B], grad[T,B,C] or None, status[B])."""
    _lib.require_cuda(logits, flat, offsets, seq_len)
    lib = _lib.load()
    T, B, C = logits.shape
    logits = logits.contiguous()
    loss = torch.empty(B, dtype=torch.float32, device=logits.device)
    grad = torch.empty_like(logits) if want_grad else None
    status = torch.empty(B, dtype=torch.int32, device=logits.device)
    need = ctypes.c_size_t(0)
    _lib.check(lib.ocr_ctc_loss_workspace_bytes(T, B, C, int(max_label_len), ctypes.byref(need)), "ocr_ctc_loss_workspace_bytes")
    ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=logits.device) if need.value else None
    _lib.check(lib.ocr_ctc_loss(_lib.ptr(logits), T, B, C, _lib.ptr(flat), _lib.ptr(offsets), _lib.ptr(seq_len),
                                int(max_label_len), _lib.ptr(loss), _lib.ptr(grad), _lib.ptr(status),
                                float(grad_scale), _lib.ptr(ws), need.value, _lib.stream_handle()), "ocr_ctc_loss")
    return loss, grad, status


class _CtcLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, flat, offsets, seq_len, max_label_len):
        loss, grad, _ = ctc_loss_raw(logits, flat, offsets, seq_len, max_label_len, want_grad=True)
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        (grad,) = ctx.saved_tensors
        return grad * gloss.view(1, -1, 1), None, None, None, None


def ctc_loss(labels, inputs, sequence_length, preprocess_collapse_repeated=False, ctc_merge_repeated=True,
             ignore_longer_outputs_than_inputs=False, time_major=True):
    """tf.nn.ctc_loss(labels, inputs, sequence_length, ..., time_major=True) -> loss [B] float32.

    The per-example gradient w.r.t. `inputs` is produced by the same kernel launch and handed to
    autograd, exactly as TensorFlow's op emits it as a second output."""
    if preprocess_collapse_repeated or not ctc_merge_repeated:
        raise NotImplementedError("only the reference's configuration (preprocess_collapse_repeated=False, "
                                  "ctc_merge_repeated=True; model.py:226) is implemented")
    if not time_major:
        inputs = inputs.transpose(0, 1)
    _lib.require_cuda(inputs)
    T, B, C = inputs.shape
    flat, offsets, lengths, flat_host = _labels_to_flat(labels, B, inputs.device)
    sl_host = sequence_length if not sequence_length.is_cuda else sequence_length.cpu()   # host lengths: no device round trip
    seq_len = sequence_length.to(device=inputs.device, dtype=torch.int32, non_blocking=True).contiguous()
    _validate_ctc(flat_host, lengths, sl_host.numpy(), T, C, ignore_longer_outputs_than_inputs)
    return _CtcLossFn.apply(inputs.contiguous().float(), flat, offsets, seq_len, max(lengths) if lengths else 0)


# ----------------------------------------------------------------------------- decoders
def _dense_to_sparse(decoded, lengths):
    """[B,T] -1 padded + lengths -> SparseTensor(indices [N,2] int64, values [N] int64, dense_shape [2])."""
    B, T = decoded.shape
    max_len = int(lengths.max().item()) if B else 0
    mask = torch.arange(T, device=decoded.device).unsqueeze(0) < lengths.unsqueeze(1)
    idx = mask.nonzero()
    return SparseTensor(idx.to(torch.int64), decoded[mask], torch.tensor([B, max_len], dtype=torch.int64))


def sparse_tensor_to_dense(sp, default_value=0):
    B, L = int(sp.dense_shape[0]), int(sp.dense_shape[1])
    out = torch.full((B, L), default_value, dtype=sp.values.dtype, device=sp.values.device)
    if sp.values.numel():
        out[sp.indices[:, 0], sp.indices[:, 1]] = sp.values
    return out


def ctc_greedy_decode_raw(logits, seq_len, merge_repeated=True):
    """-> (decoded int64 [B,T] -1 padded, lengths int32 [B], neg_sum_logits f32 [B])."""
    _lib.require_cuda(logits, seq_len)
    lib = _lib.load()
    T, B, C = logits.shape
    logits = logits.contiguous()
    dec = torch.empty((B, T), dtype=torch.int64, device=logits.device)
    ln = torch.empty(B, dtype=torch.int32, device=logits.device)
    ns = torch.empty(B, dtype=torch.float32, device=logits.device)
    _lib.check(lib.ocr_ctc_greedy_decode(_lib.ptr(logits), T, B, C, _lib.ptr(seq_len), int(bool(merge_repeated)),
                                         _lib.ptr(dec), _lib.ptr(ln), _lib.ptr(ns), _lib.stream_handle()),
               "ocr_ctc_greedy_decode")
    return dec, ln, ns


def ctc_greedy_decoder(inputs, sequence_length, merge_repeated=True):
    """tf.nn.ctc_greedy_decoder -> ([SparseTensor decoded], neg_sum_logits [B,1])."""
    seq_len = sequence_length.to(device=inputs.device, dtype=torch.int32).contiguous()
    if int(seq_len.max().item()) > inputs.shape[0]:
        raise ValueError("sequence_length(b) <= %d required" % inputs.shape[0])
    dec, ln, ns = ctc_greedy_decode_raw(inputs.float(), seq_len, merge_repeated)
    return [_dense_to_sparse(dec, ln)], ns.view(-1, 1)


def ctc_beam_search_raw(logits, seq_len, beam_width=100, top_paths=1, merge_repeated=True, normalize=True):
    """-> (decoded int64 [B,top_paths,T] -1 padded, lengths int32 [B,top_paths], log_prob f32 [B,top_paths])."""
    _lib.require_cuda(logits, seq_len)
    lib = _lib.load()
    T, B, C = logits.shape
    logits = logits.contiguous()
    dec = torch.empty((B, top_paths, T), dtype=torch.int64, device=logits.device)
    ln = torch.empty((B, top_paths), dtype=torch.int32, device=logits.device)
    lp = torch.empty((B, top_paths), dtype=torch.float32, device=logits.device)
    need = ctypes.c_size_t(0)
    _lib.check(lib.ocr_ctc_beam_search_workspace_bytes(T, B, C, int(beam_width), ctypes.byref(need)),
               "ocr_ctc_beam_search_workspace_bytes")
    ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=logits.device)
    _lib.check(lib.ocr_ctc_beam_search(_lib.ptr(logits), T, B, C, _lib.ptr(seq_len), int(beam_width), int(top_paths),
                                       int(bool(merge_repeated)), int(bool(normalize)), _lib.ptr(dec), _lib.ptr(ln),
                                       _lib.ptr(lp), _lib.ptr(ws), need.value, _lib.stream_handle()),
               "ocr_ctc_beam_search")
    return dec, ln, lp


def ctc_beam_search_decoder(inputs, sequence_length, beam_width=100, top_paths=1, merge_repeated=True,
                            normalize=True):
    """tf.nn.ctc_beam_search_decoder -> ([SparseTensor] * top_paths, log_probability [B,top_paths])."""
    if top_paths > beam_width:
        raise ValueError("top_paths (%d) must be <= beam_width (%d)" % (top_paths, beam_width))
    seq_len = sequence_length.to(device=inputs.device, dtype=torch.int32).contiguous()
    if int(seq_len.max().item()) > inputs.shape[0]:
        raise ValueError("sequence_length(b) <= %d required" % inputs.shape[0])
    dec, ln, lp = ctc_beam_search_raw(inputs.float(), seq_len, beam_width, top_paths, merge_repeated, normalize)
    return [_dense_to_sparse(dec[:, p].contiguous(), ln[:, p].contiguous()) for p in range(top_paths)], lp


def edit_distance(hypothesis, truth, normalize=True):
    """tf.edit_distance(hypothesis SparseTensor, truth SparseTensor, normalize) -> float32 [B]."""
    lib = _lib.load()
    B = int(hypothesis.dense_shape[0])
    dev = hypothesis.values.device
    hyp = sparse_tensor_to_dense(SparseTensor(hypothesis.indices, hypothesis.values.to(torch.int64), hypothesis.dense_shape), -1)
    if hyp.shape[1] == 0:
        hyp = torch.full((B, 1), -1, dtype=torch.int64, device=dev)
    hyp = hyp.contiguous()
    flat, offsets, lengths, _ = _labels_to_flat((truth.indices, truth.values, truth.dense_shape), B, dev)
    dist = torch.empty(B, dtype=torch.float32, device=dev)
    _lib.check(lib.ocr_edit_distance(_lib.ptr(hyp), hyp.shape[1], None, _lib.ptr(flat), _lib.ptr(offsets), B,
                                     max(lengths) if lengths else 0, _lib.ptr(dist), _lib.stream_handle()),
               "ocr_edit_distance")
    if normalize:
        tl = torch.tensor(lengths, dtype=torch.float32, device=dev)
        dist = dist / tl  # TF: inf when the truth is empty and the hypothesis is not
    return dist
