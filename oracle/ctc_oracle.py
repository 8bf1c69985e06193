"""TEST INFRASTRUCTURE ONLY -- numpy/ctypes front end of the C oracle (oracle/ctc_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (cnn_lstm_ctc_ocr_b200) never does.

Function-for-function it restates the TensorFlow ops the reference calls:
  ctc_loss           <- tf.nn.ctc_loss           (/root/reference/src/weinman/model.py:224-229)
  ctc_greedy_decoder <- tf.nn.ctc_greedy_decoder (/root/reference/src/weinman/validate.py:81-92)
  ctc_beam_search_decoder <- tf.nn.ctc_beam_search_decoder (src/weinman/test.py:84-88)
  edit_distance      <- tf.edit_distance         (src/weinman/test.py:90)
Parity pin: see the header of ctc_oracle.c ("parity unpinned" against a live TensorFlow).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_ctc.so")
_lib = None


def build(force=False):
    """Compile oracle/ctc_oracle.c with gcc (recipe: oracle/Makefile)."""
    src = [os.path.join(_HERE, f) for f in ("ctc_oracle.c", "det_math.h", "Makefile")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src)):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "CC=gcc"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_det_expf.restype = ctypes.c_float
        _lib.oracle_det_expf.argtypes = [ctypes.c_float]
        _lib.oracle_det_logf.restype = ctypes.c_float
        _lib.oracle_det_logf.argtypes = [ctypes.c_float]
        _lib.oracle_det_lse2.restype = ctypes.c_float
        _lib.oracle_det_lse2.argtypes = [ctypes.c_float, ctypes.c_float]
    return _lib


def max_threads():
    return int(lib().oracle_max_threads())


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _flat_labels(labels):
    """labels: list of int sequences -> (flat int32, offsets int32[B+1])."""
    off = np.zeros(len(labels) + 1, np.int32)
    for i, l in enumerate(labels):
        off[i + 1] = off[i] + len(l)
    flat = np.zeros(max(int(off[-1]), 1), np.int32)
    for i, l in enumerate(labels):
        flat[off[i]:off[i + 1]] = np.asarray(l, np.int32)
    return flat, off


def ctc_loss(logits, labels, seq_len, want_grad=True, nthreads=1, f64=False):
    """logits [T,B,C] f32; labels list of int lists; seq_len [B] -> (loss[B], grad[T,B,C], status[B]).
    f64=True evaluates the same recursion in float64 (the ground truth tolerances are stated against)."""
    logits = np.ascontiguousarray(logits, np.float32)
    T, B, C = logits.shape
    flat, off = _flat_labels(labels)
    seq_len = np.ascontiguousarray(seq_len, np.int32)
    loss = np.zeros(B, np.float32)
    grad = np.zeros_like(logits) if want_grad else None
    status = np.zeros(B, np.int32)
    fn = lib().oracle_ctc_loss_f64 if f64 else lib().oracle_ctc_loss
    fn(_p(logits), T, B, C, _p(flat), _p(off), _p(seq_len), _p(loss),
                          _p(grad) if want_grad else None, _p(status), int(nthreads))
    return loss, grad, status


def ctc_greedy_decoder(logits, seq_len, merge_repeated=True):
    """-> (decoded int64 [B,T] (-1 padded), lengths int32 [B], neg_sum_logits f32 [B,1])."""
    logits = np.ascontiguousarray(logits, np.float32)
    T, B, C = logits.shape
    seq_len = np.ascontiguousarray(seq_len, np.int32)
    dec = np.full((B, max(T, 1)), -1, np.int64)
    ln = np.zeros(B, np.int32)
    ns = np.zeros(B, np.float32)
    lib().oracle_ctc_greedy(_p(logits), T, B, C, _p(seq_len), int(bool(merge_repeated)), _p(dec), _p(ln), _p(ns), 1)
    return dec[:, :T], ln, ns.reshape(B, 1)


def ctc_beam_search_decoder(logits, seq_len, beam_width=100, top_paths=1, merge_repeated=True,
                            normalize=True, det_math=True, nthreads=1):
    """-> (decoded int64 [B,top_paths,T] (-1 padded), lengths [B,top_paths], log_prob f32 [B,top_paths])."""
    logits = np.ascontiguousarray(logits, np.float32)
    T, B, C = logits.shape
    seq_len = np.ascontiguousarray(seq_len, np.int32)
    dec = np.full((B, top_paths, max(T, 1)), -1, np.int64)
    ln = np.zeros((B, top_paths), np.int32)
    lp = np.zeros((B, top_paths), np.float32)
    rc = lib().oracle_ctc_beam(_p(logits), T, B, C, _p(seq_len), int(beam_width), int(top_paths),
                               int(bool(merge_repeated)), int(bool(normalize)), int(bool(det_math)),
                               _p(dec), _p(ln), _p(lp), int(nthreads))
    if rc != 0:
        raise ValueError("top_paths > beam_width")
    return dec[:, :, :T], ln, lp


def edit_distance(hyp, truth):
    """hyp, truth: lists of int sequences -> float32 [B] Levenshtein distances (normalize=False)."""
    B = len(hyp)
    hf, ho = _flat_labels(hyp)
    tf_, to = _flat_labels(truth)
    hf = hf.astype(np.int64)
    tf_ = tf_.astype(np.int64)
    out = np.zeros(B, np.float32)
    lib().oracle_edit_distance(_p(hf), _p(ho), _p(tf_), _p(to), B, _p(out))
    return out


def densify(dec, ln):
    """sparse_tensor_to_dense(default_value=-1) of a decode: trim to the longest row (validate.py:91)."""
    m = int(ln.max()) if ln.size else 0
    return dec[..., :m]
