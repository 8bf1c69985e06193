"""GPU tier: CTC beam search (sm_100a, warp per sequence) against the CPU oracle.
Bar: bit-exact labels, lengths AND float32 log-probability bits (both sides evaluate the same
deterministic exp/log recipe), including ReLU-tied logits and TF's order-dependent pruning."""
import os

import numpy as np
import pytest
import torch

from util import cfg2_inputs

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "tf_unit_vectors.npz"))


def _gpu_beam(x, seq_len, K, top_paths, merge, normalize=True):
    from cnn_lstm_ctc_ocr_b200 import ctc
    dev = torch.device("cuda:0")
    dec, ln, lp = ctc.ctc_beam_search_raw(torch.tensor(x, device=dev), torch.tensor(np.asarray(seq_len, np.int32), device=dev),
                                          K, top_paths, merge, normalize)
    torch.cuda.synchronize()
    return dec.cpu().numpy(), ln.cpu().numpy(), lp.cpu().numpy()


def _check(oracle, x, seq_len, K, top_paths, merge, normalize=True):
    d, l, p = _gpu_beam(x, seq_len, K, top_paths, merge, normalize)
    od, ol, op = oracle.ctc_beam_search_decoder(x, seq_len, K, top_paths, merge, normalize=normalize, det_math=True, nthreads=16)
    bad = np.nonzero((d != od).any(axis=(1, 2)) | (l != ol).any(axis=1))[0]
    assert bad.size == 0, "decode mismatch in sequences %s" % bad[:10]
    assert (p.view(np.uint32) == op.view(np.uint32)).all(), "log_prob bits differ: max abs %g" % np.abs(p - op).max()


def test_beam_tf_unit_vectors():
    bl = np.log(G["beam_p"]) + G["beam_offset"]
    inp = np.concatenate([bl[:5], np.zeros((3, 6), np.float32)])[:, None, :].astype(np.float32)
    d, l, p = _gpu_beam(inp, [5], 2, 2, True, normalize=False)
    assert d[0, 0, :l[0, 0]].tolist() == [1, 0] and d[0, 1, :l[0, 1]].tolist() == [0, 1, 0]
    np.testing.assert_allclose(p[0], G["beam_logprob_vmax"], atol=2e-6)


@pytest.mark.parametrize("C,scale,relu", [(63, 3.0, False), (96, 3.0, False), (63, 1.0, False), (96, 0.3, True), (63, 3.0, True)])
def test_beam_width128_bit_exact(oracle, C, scale, relu):
    x, _, seq_len = cfg2_inputs(seed=20 + C, T=64, B=96, C=C, relu=relu, scale=scale)
    seq_len[5] = 0
    seq_len[6] = 1
    _check(oracle, x, seq_len, 128, 1, True)


@pytest.mark.parametrize("K,top_paths,merge,normalize", [(1, 1, True, True), (2, 2, False, True), (16, 3, True, False),
                                                           (100, 1, False, True), (37, 5, True, True)])
def test_beam_variants_bit_exact(oracle, K, top_paths, merge, normalize):
    x, _, seq_len = cfg2_inputs(seed=K, T=40, B=64, C=20, scale=2.0, relu=(K % 2 == 0))
    _check(oracle, x, seq_len, K, top_paths, merge, normalize)


def test_beam_small_alphabet_and_long(oracle):
    x, _, seq_len = cfg2_inputs(seed=3, T=200, B=8, C=5, scale=1.0)
    _check(oracle, x, seq_len, 128, 2, True)
    x, _, seq_len = cfg2_inputs(seed=4, T=30, B=8, C=200, scale=2.0)
    _check(oracle, x, seq_len, 64, 1, False)


def test_beam_api_and_metrics(oracle):
    """test._get_testing's decode + metrics (src/weinman/test.py:84-99)."""
    from cnn_lstm_ctc_ocr_b200 import ctc
    x, labels, seq_len = cfg2_inputs(seed=8, T=32, B=16, C=30, scale=3.0)
    dev = torch.device("cuda:0")
    sp, lp = ctc.ctc_beam_search_decoder(torch.tensor(x, device=dev), torch.tensor(seq_len), beam_width=128, top_paths=1)
    od, ol, op = oracle.ctc_beam_search_decoder(x, seq_len, 128, 1, True, nthreads=8)
    dense = ctc.sparse_tensor_to_dense(sp[0], -1).cpu().numpy()
    assert (dense == oracle.densify(od[:, 0], ol[:, 0])).all()
    idx = torch.tensor([[b, i] for b, l in enumerate(labels) for i in range(len(l))], dtype=torch.int64)
    truth = ctc.SparseTensor(idx.to(dev), torch.tensor(sum(labels, []), dtype=torch.int32, device=dev), torch.tensor([16, 16]))
    dist = ctc.edit_distance(sp[0], truth, normalize=False).cpu().numpy()
    ref = oracle.edit_distance([od[b, 0, :ol[b, 0]].tolist() for b in range(16)], labels)
    assert (dist == ref).all()
    with pytest.raises(ValueError):
        ctc.ctc_beam_search_decoder(torch.tensor(x, device=dev), torch.tensor(seq_len), beam_width=2, top_paths=3)


def test_cta_kernel_equals_the_replay_kernel():
    """The default kernel (one CTA per sequence: counts, selection and TF's reset rule stated in closed form) against the
    warp-per-sequence kernel that replays TensorFlow's list updates one insertion at a time: same labels, lengths and score
    bits on flat, peaked and ReLU-tied inputs, several widths and alphabets."""
    from cnn_lstm_ctc_ocr_b200 import _lib
    lib = _lib.load()
    cases = [(63, 3.0, False, 128, 64, 64), (63, 8.0, False, 128, 64, 48), (96, 0.3, True, 128, 61, 32), (20, 2.0, True, 37, 40, 64),
             (5, 1.0, False, 128, 120, 8), (200, 2.0, False, 64, 30, 8), (63, 0.05, False, 128, 64, 32), (63, 20.0, True, 100, 64, 32)]
    for C, scale, relu, K, T, B in cases:
        x, _, seq_len = cfg2_inputs(seed=7 * C + K, T=T, B=B, C=C, relu=relu, scale=scale)
        outs = []
        try:
            for path in (0, 1):
                _lib.check(lib.ocr_debug_beam_path(path), "beam_path")
                outs.append(_gpu_beam(x, seq_len, K, min(3, K), True))
        finally:
            lib.ocr_debug_beam_path(0)
        for a, b_ in zip(outs[0], outs[1]):
            assert np.array_equal(a.view(np.uint32) if a.dtype == np.float32 else a, b_.view(np.uint32) if b_.dtype == np.float32 else b_), (C, scale, relu, K)
