"""CPU tier: pins the oracle (oracle/ctc_oracle.c) on upstream TensorFlow's unit-test vectors,
on brute-force path enumeration and on torch.nn.functional.ctc_loss."""
import itertools
import os

import numpy as np
import pytest
import torch

from util import cfg2_inputs

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "tf_unit_vectors.npz"))


def test_tf_ctc_loss_unit_vectors(oracle):
    with np.errstate(divide="ignore"):
        logits = np.stack([np.log(G["loss_p0"]), np.log(G["loss_p1"])], axis=1)
    loss, grad, st = oracle.ctc_loss(logits, [G["loss_targets0"].tolist(), G["loss_targets1"].tolist()], [5, 5])
    assert st.tolist() == [0, 0]
    np.testing.assert_allclose(loss, [G["loss_value0"], G["loss_value1"]], rtol=0, atol=2e-5)
    np.testing.assert_allclose(grad[:, 0], G["loss_g0"], atol=2e-6)
    np.testing.assert_allclose(grad[:, 1], G["loss_g1"], atol=2e-6)


def test_tf_greedy_unit_vectors(oracle):
    with np.errstate(divide="ignore"):
        logits = np.stack([np.log(G["greedy_p0"]), np.log(G["greedy_p1"])], axis=1)
    dec, ln, ns = oracle.ctc_greedy_decoder(logits, G["greedy_seq_len"])
    assert ln.tolist() == [2, 3]
    assert dec[0, :2].tolist() == G["greedy_decode0"].tolist()
    assert dec[1, :3].tolist() == G["greedy_decode1"].tolist()
    assert (dec[0, 2:] == -1).all() and (dec[1, 3:] == -1).all()
    truth = [np.sum(-np.log([1.0, 0.6, 0.6, 0.9])), np.sum(-np.log([0.9] * 5))]
    np.testing.assert_allclose(ns.ravel(), truth, rtol=1e-6)


@pytest.mark.parametrize("det", [False, True])
def test_tf_beam_unit_vectors(oracle, det):
    bl = np.log(G["beam_p"]) + G["beam_offset"]
    inp = np.concatenate([bl[:5], np.zeros((3, 6), np.float32)])[:, None, :]
    dec, ln, lp = oracle.ctc_beam_search_decoder(inp, [5], beam_width=2, top_paths=2, merge_repeated=True,
                                                 normalize=False, det_math=det)
    assert dec[0, 0, :ln[0, 0]].tolist() == G["beam_decode0"].tolist()
    assert dec[0, 1, :ln[0, 1]].tolist() == G["beam_decode1"].tolist()
    np.testing.assert_allclose(lp[0], G["beam_logprob_vmax"], atol=2e-6)
    # log-softmax variant: same beams, scores shifted by the per-frame normaliser
    dec2, ln2, lp2 = oracle.ctc_beam_search_decoder(inp, [5], beam_width=2, top_paths=2, normalize=True, det_math=det)
    assert (dec2 == dec).all()
    shift = np.sum(np.log(G["beam_p"][:5].max(axis=1)))
    np.testing.assert_allclose(lp2[0], G["beam_logprob_vmax"] + shift, atol=1e-5)


def _brute_force(logits, label, blank, merge=True):
    """Sum of path probabilities over all alignments that collapse to `label` (T small)."""
    T, C = logits.shape
    y = np.exp(logits - logits.max(1, keepdims=True)).astype(np.float64)
    y /= y.sum(1, keepdims=True)
    total = 0.0
    for path in itertools.product(range(C), repeat=T):
        out, prev = [], -1
        for c in path:
            if c != blank and c != prev:
                out.append(c)
            prev = c
        if out == list(label):
            total += np.prod([y[t, c] for t, c in enumerate(path)])
    return total


@pytest.mark.parametrize("label", [[], [0], [1, 1], [0, 1], [2, 0, 2], [1, 1, 1]])
def test_loss_vs_brute_force(oracle, label):
    rng = np.random.default_rng(len(label) + 7)
    T, C = 5, 4
    x = rng.standard_normal((T, 1, C)).astype(np.float32) * 2
    p = _brute_force(x[:, 0], label, C - 1)
    for f64 in (False, True):
        loss, grad, st = oracle.ctc_loss(x, [label], [T], f64=f64)
        assert st[0] == 0
        np.testing.assert_allclose(loss[0], -np.log(p), rtol=2e-6 if f64 else 2e-5)
        # finite-difference check of the gradient through the brute force
        eps = 1e-3
        for (t, k) in [(0, 0), (2, 1), (4, 3)]:
            xp = x.copy(); xp[t, 0, k] += eps
            xm = x.copy(); xm[t, 0, k] -= eps
            fd = (-np.log(_brute_force(xp[:, 0], label, C - 1)) + np.log(_brute_force(xm[:, 0], label, C - 1))) / (2 * eps)
            assert abs(fd - grad[t, 0, k]) < 2e-3


def test_loss_edge_cases(oracle):
    rng = np.random.default_rng(3)
    T, C = 6, 5
    x = rng.standard_normal((T, 4, C)).astype(np.float32)
    # b0: infeasible (needs 2*3-1 = 5 frames incl. repeats, has 4); b1: zero-length; b2: exactly fits; b3: padded
    loss, grad, st = oracle.ctc_loss(x, [[1, 1, 1], [0], [1, 1, 1], [2]], [4, 0, 5, 3])
    assert st.tolist() == [2, 0, 0, 0]
    assert loss[1] == 0 and (grad[:, 1] == 0).all()
    assert np.isfinite(loss[2]) and (grad[5:, 2] == 0).all()
    assert (grad[3:, 3] == 0).all()
    np.testing.assert_allclose(grad[:5, 2].sum(-1), 0, atol=5e-6)  # rows of softmax - posterior sum to zero
    # no valid path: -inf logit on the only class that can start the alignment
    x2 = x.copy()
    x2[0, 2, :] = [0, -np.inf, 0, 0, -np.inf]  # label 1 and blank impossible at t=0
    loss, grad, st = oracle.ctc_loss(x2, [[1, 1, 1], [0], [1, 1, 1], [2]], [4, 0, 5, 3])
    assert st[2] == 1 and np.isinf(loss[2])


def test_loss_vs_torch_and_f64(oracle):
    x, labels, seq_len = cfg2_inputs(seed=1, T=64, B=64, C=63)
    l32, g32, _ = oracle.ctc_loss(x, labels, seq_len, nthreads=4)
    l64, g64, _ = oracle.ctc_loss(x, labels, seq_len, nthreads=4, f64=True)
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    tl = torch.nn.functional.ctc_loss(torch.log_softmax(xt, -1), torch.tensor(sum(labels, [])),
                                      torch.tensor(seq_len.astype(np.int64)), torch.tensor([len(l) for l in labels]),
                                      blank=62, reduction="none")
    tl.sum().backward()
    np.testing.assert_allclose(l64, tl.detach().numpy(), rtol=1e-6)
    np.testing.assert_allclose(g64, xt.grad.numpy(), atol=1e-6)
    # the TF-faithful float32 recursion is itself only this close to the exact answer
    np.testing.assert_allclose(l32, l64, rtol=2e-6)
    assert np.abs(g32 - g64).max() < 5e-4


def test_greedy_ties_and_merge(oracle):
    # ReLU logits produce exact ties: first maximum wins
    x = np.zeros((4, 1, 5), np.float32)
    x[1, 0, 2] = x[1, 0, 3] = 1.0  # tie between 2 and 3 -> 2
    x[2, 0, 2] = 1.0               # repeat of 2 -> merged
    x[3, 0, 4] = 1.0               # blank
    dec, ln, ns = oracle.ctc_greedy_decoder(x, [4])
    assert dec[0, :ln[0]].tolist() == [0, 2]
    dec, ln, _ = oracle.ctc_greedy_decoder(x, [4], merge_repeated=False)
    assert dec[0, :ln[0]].tolist() == [0, 2, 2]
    assert ns[0, 0] == -3.0


def test_beam_matches_exhaustive(oracle):
    """With a beam wider than the number of prefixes, the top path is the most probable labelling."""
    rng = np.random.default_rng(11)
    T, C = 4, 3
    for trial in range(5):
        x = (rng.standard_normal((T, 1, C)) * 2).astype(np.float32)
        dec, ln, lp = oracle.ctc_beam_search_decoder(x, [T], beam_width=128, top_paths=3, merge_repeated=False)
        cands = {}
        for L in range(0, T + 1):
            for lab in itertools.product(range(C - 1), repeat=L):
                p = _brute_force(x[:, 0], list(lab), C - 1)
                if p > 0:
                    cands[lab] = p
        best = sorted(cands.items(), key=lambda kv: -kv[1])[:3]
        for p_i, (lab, p) in enumerate(best):
            assert tuple(dec[0, p_i, :ln[0, p_i]].tolist()) == lab
            np.testing.assert_allclose(lp[0, p_i], np.log(p), atol=2e-5)


def test_beam_det_math_agrees_with_libm(oracle):
    x, _, seq_len = cfg2_inputs(seed=5, T=24, B=16, C=20, scale=3.0)
    a = oracle.ctc_beam_search_decoder(x, seq_len, 16, 2, True, det_math=True)
    b = oracle.ctc_beam_search_decoder(x, seq_len, 16, 2, True, det_math=False)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
    np.testing.assert_allclose(a[2], b[2], rtol=1e-5, atol=1e-5)


def test_det_math_accuracy(oracle):
    L = oracle.lib()
    xs = np.concatenate([np.linspace(-85, 0, 4001), -np.logspace(-8, 1.9, 500)]).astype(np.float32)
    err = max(abs(L.oracle_det_expf(float(v)) - np.exp(np.float64(v))) / np.exp(np.float64(v)) for v in xs)
    assert err < 3e-7
    ys = np.concatenate([np.linspace(1, 2, 2001), np.logspace(-3, 3, 2001)]).astype(np.float32)
    err = max(abs(L.oracle_det_logf(float(v)) - np.log(np.float64(v))) for v in ys)
    assert err < 5e-7
    assert L.oracle_det_lse2(float("-inf"), float("-inf")) == float("-inf")
    assert abs(L.oracle_det_lse2(0.0, 0.0) - np.log(2)) < 2e-7


def test_edit_distance(oracle):
    d = oracle.edit_distance([[1, 2, 3], [], [1, 2], [5]], [[1, 3], [4, 4], [1, 2], []])
    assert d.tolist() == [1.0, 2.0, 0.0, 1.0]
