"""GPU tier: the sm_100a CTC kernels, called through the C ABI, against the CPU oracle.

Tolerances:
  loss  : rel 1e-5 vs the float64 recursion and vs the TF-faithful float32 oracle;
  grad  : fast kernel (linear-domain lattice, exact power-of-two rescaling): abs 1e-5 vs the float64
          recursion.  Against the TF-faithful float32 oracle the bar is 3e-4, because THAT recursion
          (TF's own float32 log-domain arithmetic) carries ~1 ulp(|alpha|) ~ 3e-5 of noise per step
          (oracle-f32 vs oracle-f64 differ by up to ~1e-4 on the same inputs, asserted in
          test_oracle_golden.py).  The general (long-sequence) kernel is log-domain float32 like TF
          and is held to 3e-4 / 3e-3 (T=509);
  decode: bit-exact (labels, lengths, and neg_sum_logits float bits).
"""
import os

import numpy as np
import pytest
import torch

from util import cfg2_inputs, make_labels

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "tf_unit_vectors.npz"))


def _gpu_loss(x, labels, seq_len, want_grad=True, path=0):
    from cnn_lstm_ctc_ocr_b200 import ctc, _lib
    dev = torch.device("cuda:0")
    _lib.check(_lib.load().ocr_ctc_loss_set_path(path), "set_path")
    xt = torch.tensor(x, device=dev)
    flat, off, lengths, _ = ctc._labels_to_flat(labels, x.shape[1], dev)
    sl = torch.tensor(np.asarray(seq_len, np.int32), device=dev)
    loss, grad, st = ctc.ctc_loss_raw(xt, flat, off, sl, max(lengths) if lengths else 0, want_grad)
    torch.cuda.synchronize()
    _lib.load().ocr_ctc_loss_set_path(0)
    return loss.cpu().numpy(), (grad.cpu().numpy() if want_grad else None), st.cpu().numpy()


def test_loss_tf_unit_vectors():
    with np.errstate(divide="ignore"):
        logits = np.stack([np.log(G["loss_p0"]), np.log(G["loss_p1"])], axis=1)
    loss, grad, st = _gpu_loss(logits, [G["loss_targets0"].tolist(), G["loss_targets1"].tolist()], [5, 5])
    assert st.tolist() == [0, 0]
    np.testing.assert_allclose(loss, [G["loss_value0"], G["loss_value1"]], atol=2e-5)
    np.testing.assert_allclose(grad[:, 0], G["loss_g0"], atol=2e-6)
    np.testing.assert_allclose(grad[:, 1], G["loss_g1"], atol=2e-6)


@pytest.mark.parametrize("path", [0, 2, 1, 8])   # 8: the streaming kernel (csrc/ctc_loss_stream.cuh)
@pytest.mark.parametrize("C,ragged,relu", [(63, False, False), (63, True, False), (63, True, True), (96, True, False)])
def test_loss_cfg2_vs_oracle(oracle, C, ragged, relu, path):
    x, labels, seq_len = cfg2_inputs(seed=1, T=64, B=256, C=C, ragged=ragged, relu=relu)
    loss, grad, st = _gpu_loss(x, labels, seq_len, path=path)
    l32, g32, s32 = oracle.ctc_loss(x, labels, seq_len, nthreads=8)
    l64, g64, _ = oracle.ctc_loss(x, labels, seq_len, nthreads=8, f64=True)
    assert (st == s32).all() and (st == 0).all()
    np.testing.assert_allclose(loss, l64, rtol=1e-5)
    np.testing.assert_allclose(loss, l32, rtol=1e-5)
    assert np.abs(grad - g32).max() < 3e-4
    assert np.abs(grad - g64).max() < (3e-4 if path == 1 else 1e-5)
    # size-independent properties: rows inside the sequence sum to zero, rows past it are zero
    for b in range(0, 256, 17):
        assert (grad[seq_len[b]:, b] == 0).all()
        assert np.abs(grad[:seq_len[b], b].sum(-1)).max() < (3e-4 if path == 1 else 2e-6)


def test_loss_many_waves_prefetch_and_pdl(oracle):
    """B = 2048 is several waves of the fast kernel (512 CTAs of 4 sequences, 296 resident on a B200): the L2 prefetch of the
    successor group is live.  Parity against the oracle, and the launch-time knobs (prefetch distance, programmatic
    dependent launch, speculative box requests) must not change a single bit."""
    from cnn_lstm_ctc_ocr_b200 import _lib
    lib = _lib.load()
    x, labels, seq_len = cfg2_inputs(seed=5, T=64, B=2048, C=63, ragged=True)
    l64, g64, _ = oracle.ctc_loss(x, labels, seq_len, nthreads=8, f64=True)
    outs = []
    try:
        for prefetch, pdl, spec in ((-1, 1, 1), (0, 1, 1), (37, 1, 0), (-1, 0, 0)):
            _lib.check(lib.ocr_debug_ctc_prefetch(prefetch), "prefetch")
            _lib.check(lib.ocr_debug_ctc_pdl(pdl), "pdl")
            _lib.check(lib.ocr_debug_ctc_speculate(spec), "speculate")
            loss, grad, st = _gpu_loss(x, labels, seq_len)
            assert (st == 0).all()
            outs.append((loss, grad))
    finally:
        lib.ocr_debug_ctc_prefetch(-1)
        lib.ocr_debug_ctc_pdl(1)
        lib.ocr_debug_ctc_speculate(1)
    np.testing.assert_allclose(outs[0][0], l64, rtol=1e-5)
    assert np.abs(outs[0][1] - g64).max() < 1e-5
    for loss, grad in outs[1:]:
        assert np.array_equal(loss, outs[0][0]) and np.array_equal(grad, outs[0][1])
    for b in range(0, 2048, 101):
        assert (outs[0][1][seq_len[b]:, b] == 0).all()


@pytest.mark.parametrize("T,B,C,maxlen,path", [
    (64, 255, 63, 16, 0),    # B*C not a multiple of 4: LSU path, partial last group
    (64, 8, 63, 16, 0),      # one full TMA group of 8 / two of 4
    (61, 37, 96, 20, 0),     # partial last group on the TMA path
    (61, 37, 96, 20, 2),
    (40, 21, 7, 30, 0),      # labels as long as the sequence allows, tiny alphabet
    (125, 64, 96, 24, 0),    # cfg3 shape: one sequence per CTA
    (200, 9, 30, 60, 0),     # 2 label pairs per lane
    (250, 5, 20, 100, 0),    # 4 label pairs per lane
    (3, 4, 5, 2, 0), (1, 4, 5, 1, 0), (2, 3, 4, 1, 2),
    (64, 8, 63, 16, 8), (61, 36, 63, 20, 8), (125, 64, 63, 24, 8), (48, 6, 66, 30, 8),   # streaming kernel: partial last box, G = 4 and 2
])
def test_loss_shapes_vs_f64(oracle, T, B, C, maxlen, path):
    rng = np.random.default_rng(T * 1000 + B)
    x = (rng.standard_normal((T, B, C)) * 2).astype(np.float32)
    seq_len = rng.integers(max(1, T // 2), T + 1, B).astype(np.int32)
    seq_len[0] = T
    labels = make_labels(rng, B, seq_len, max_len=maxlen, num_labels=C - 1, repeat_p=0.3)
    if B > 2:
        labels[1] = []          # empty label: only the all-blank alignment
        seq_len[2] = 0          # zero-length sequence: loss 0, grad 0
    loss, grad, st = _gpu_loss(x, labels, seq_len, path=path)
    l64, g64, s64 = oracle.ctc_loss(x, labels, seq_len, nthreads=8, f64=True)
    assert (st == s64).all()
    np.testing.assert_allclose(loss, l64, rtol=1e-5, atol=1e-6)
    assert np.abs(grad - g64).max() < 2e-5
    loss2, _, _ = _gpu_loss(x, labels, seq_len, want_grad=False, path=path)
    assert (loss2 == loss).all()


def test_loss_extreme_dynamic_range(oracle):
    """Peaked (confident, often wrong) distributions: per-frame label probabilities down to ~e^-50."""
    rng = np.random.default_rng(77)
    T, B, C = 64, 32, 63
    x = (rng.standard_normal((T, B, C)) * 8).astype(np.float32)
    seq_len = np.full(B, T, np.int32)
    labels = make_labels(rng, B, seq_len, max_len=16, num_labels=C - 1)
    loss, grad, st = _gpu_loss(x, labels, seq_len)
    l64, g64, s64 = oracle.ctc_loss(x, labels, seq_len, nthreads=8, f64=True)
    assert (st == 0).all()
    np.testing.assert_allclose(loss, l64, rtol=1e-5)
    # sequences whose lattice leaves the float32 range of the linear-domain kernel are recomputed by the
    # log-domain kernel, which has TF's own float32 noise level (~ulp(|log p|) = 6e-5 per step at |log p| ~ 1000)
    assert np.abs(grad - g64).max() < 1e-3


def test_loss_redo_in_kernel_tail_equals_gate_launch(oracle):
    """Sequences the fast kernel flags (lattice outside the float32 range) are redone by the exact routine either in the
    fast kernel's own tail (default) or by the separate gate launch: same bits both ways, with and without a gradient."""
    from cnn_lstm_ctc_ocr_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(78)
    T, B, C = 64, 64, 63
    x = (rng.standard_normal((T, B, C)) * 8).astype(np.float32)
    seq_len = rng.integers(T // 2, T + 1, B).astype(np.int32)
    labels = make_labels(rng, B, seq_len, max_len=16, num_labels=C - 1)
    _, _, flags = _gpu_loss(x, labels, seq_len, path=3)          # diagnostics path: flags left in status
    assert 0 < int((flags == 100).sum()) < B                      # some sequences take the redo, some do not
    l64, g64, s64 = oracle.ctc_loss(x, labels, seq_len, nthreads=8, f64=True)
    res = {}
    try:
        for inline in (1, 0):
            _lib.check(lib.ocr_debug_ctc_inline_redo(inline), "inline_redo")
            res[inline] = _gpu_loss(x, labels, seq_len) + (_gpu_loss(x, labels, seq_len, want_grad=False)[0],)
    finally:
        lib.ocr_debug_ctc_inline_redo(1)
    for inline in (1, 0):
        loss, grad, st, loss_only = res[inline]
        assert (st == 0).all()
        np.testing.assert_allclose(loss, l64, rtol=1e-5)
        assert np.abs(grad - g64).max() < 1e-3
        assert np.array_equal(loss, loss_only)
    assert np.array_equal(res[1][0], res[0][0]) and np.array_equal(res[1][1], res[0][1])


def test_loss_edge_cases(oracle):
    rng = np.random.default_rng(3)
    T, C = 6, 5
    x = rng.standard_normal((T, 5, C)).astype(np.float32)
    labels = [[1, 1, 1], [0], [1, 1, 1], [2], []]
    sl = [4, 0, 5, 3, 6]
    loss, grad, st = _gpu_loss(x, labels, sl)
    lo, go, so = oracle.ctc_loss(x, labels, sl)
    assert st.tolist() == so.tolist() == [2, 0, 0, 0, 0]
    np.testing.assert_allclose(loss, lo, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(grad, go, atol=2e-6)
    x2 = x.copy()
    x2[0, 2, :] = [0, -np.inf, 0, 0, -np.inf]
    loss, grad, st = _gpu_loss(x2, labels, sl)
    lo, go, so = oracle.ctc_loss(x2, labels, sl)
    assert st[2] == 1 and np.isinf(loss[2]) and so[2] == 1
    np.testing.assert_allclose(grad[:, 2], go[:, 2], atol=2e-6)  # "no valid path": gradient = softmax


def test_loss_long_sequences_use_workspace(oracle):
    """T=509 (a 1024-px crop), long labels: lattice no longer fits in shared memory."""
    rng = np.random.default_rng(9)
    T, B, C = 509, 6, 96
    x = rng.standard_normal((T, B, C)).astype(np.float32)
    sl = np.array([509, 400, 509, 37, 255, 509], np.int32)
    labels = make_labels(rng, B, sl, max_len=120, num_labels=95)
    labels[0] = [int(v) for v in rng.integers(0, 95, 120)]
    from cnn_lstm_ctc_ocr_b200 import _lib
    import ctypes
    need = ctypes.c_size_t(0)
    _lib.load().ocr_ctc_loss_workspace_bytes(T, B, C, 120, ctypes.byref(need))
    assert need.value > 1 << 20  # no shared-memory configuration fits: general kernel + global lattice workspace
    loss, grad, st = _gpu_loss(x, labels, sl)
    l64, g64, _ = oracle.ctc_loss(x, labels, sl, f64=True, nthreads=6)
    assert (st == 0).all()
    np.testing.assert_allclose(loss, l64, rtol=1e-5)
    assert np.abs(grad - g64).max() < 3e-3


def test_loss_autograd_and_validation():
    from cnn_lstm_ctc_ocr_b200 import ctc
    x, labels, seq_len = cfg2_inputs(seed=4, T=32, B=8, C=20)
    dev = torch.device("cuda:0")
    xt = torch.tensor(x, device=dev, requires_grad=True)
    loss = ctc.ctc_loss(labels, xt, torch.tensor(seq_len))
    loss.mean().backward()
    ref = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    tl = torch.nn.functional.ctc_loss(torch.log_softmax(ref, -1), torch.tensor(sum(labels, [])),
                                      torch.tensor(seq_len.astype(np.int64)), torch.tensor([len(l) for l in labels]),
                                      blank=19, reduction="none")
    tl.mean().backward()
    np.testing.assert_allclose(loss.detach().cpu().numpy(), tl.detach().numpy(), rtol=1e-5)
    np.testing.assert_allclose(xt.grad.cpu().numpy(), ref.grad.numpy(), atol=1e-5)
    # sparse-triple labels (the reference's SparseTensor), same result
    idx = torch.tensor([[b, i] for b, l in enumerate(labels) for i in range(len(l))])
    vals = torch.tensor(sum(labels, []), dtype=torch.int32)
    loss2 = ctc.ctc_loss((idx, vals, torch.tensor([8, 16])), xt.detach(), torch.tensor(seq_len))
    assert torch.equal(loss2, loss.detach())
    with pytest.raises(ValueError, match="Not enough time"):
        ctc.ctc_loss([[1, 1, 1, 1]] + labels[1:], xt.detach(), torch.tensor([5] + seq_len[1:].tolist()))
    with pytest.raises(ValueError):
        ctc.ctc_loss([[19]] + labels[1:], xt.detach(), torch.tensor(seq_len))


@pytest.mark.parametrize("C,relu", [(63, False), (96, True), (5, True)])
def test_greedy_bit_exact(oracle, C, relu):
    from cnn_lstm_ctc_ocr_b200 import ctc
    x, _, seq_len = cfg2_inputs(seed=2, T=61, B=300, C=C, relu=relu, scale=0.5 if relu else 1.0)
    seq_len[3] = 0
    dev = torch.device("cuda:0")
    for merge in (True, False):
        dec, ln, ns = ctc.ctc_greedy_decode_raw(torch.tensor(x, device=dev), torch.tensor(seq_len, device=dev), merge)
        od, ol, on = oracle.ctc_greedy_decoder(x, seq_len, merge)
        assert (dec.cpu().numpy() == od).all()
        assert (ln.cpu().numpy() == ol).all()
        assert (ns.cpu().numpy().view(np.uint32) == on.ravel().view(np.uint32)).all()  # float bits
    sp, nsl = ctc.ctc_greedy_decoder(torch.tensor(x, device=dev), torch.tensor(seq_len))
    dense = ctc.sparse_tensor_to_dense(sp[0], -1).cpu().numpy()
    assert (dense == oracle.densify(od if False else oracle.ctc_greedy_decoder(x, seq_len, True)[0], oracle.ctc_greedy_decoder(x, seq_len, True)[1])).all()
    assert nsl.shape == (300, 1)


def test_greedy_tf_unit_vectors():
    from cnn_lstm_ctc_ocr_b200 import ctc
    with np.errstate(divide="ignore"):
        logits = np.stack([np.log(G["greedy_p0"]), np.log(G["greedy_p1"])], axis=1)
    dev = torch.device("cuda:0")
    dec, ln, ns = ctc.ctc_greedy_decode_raw(torch.tensor(logits, device=dev), torch.tensor([4, 5], dtype=torch.int32, device=dev))
    assert ln.tolist() == [2, 3]
    assert dec[0, :2].tolist() == [0, 1] and dec[1, :3].tolist() == [1, 1, 0]


def test_edit_distance_vs_oracle(oracle):
    from cnn_lstm_ctc_ocr_b200 import ctc
    rng = np.random.default_rng(5)
    B = 200
    hyp = [rng.integers(0, 6, rng.integers(0, 70)).tolist() for _ in range(B)]
    tru = [rng.integers(0, 6, rng.integers(0, 70)).tolist() for _ in range(B)]
    hyp[0], tru[0] = [], []
    hyp[1], tru[1] = [1, 2], []
    hyp[2], tru[2] = [], [3]
    dev = torch.device("cuda:0")

    def sparse(seqs, dt):
        idx = torch.tensor([[b, i] for b, l in enumerate(seqs) for i in range(len(l))], dtype=torch.int64).reshape(-1, 2)
        vals = torch.tensor(sum(seqs, []), dtype=dt)
        return ctc.SparseTensor(idx.to(dev), vals.to(dev), torch.tensor([len(seqs), max(len(l) for l in seqs)]))
    d = ctc.edit_distance(sparse(hyp, torch.int64), sparse(tru, torch.int32), normalize=False)
    assert (d.cpu().numpy() == oracle.edit_distance(hyp, tru)).all()
