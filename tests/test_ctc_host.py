"""CPU tier: host-side argument handling of cnn_lstm_ctc_ocr_b200.ctc (label packing, TensorFlow's validation rules and
error texts: SURVEY.md section 8b) against a plain-loop restatement.  No kernel is called."""
import numpy as np
import pytest
import torch

from cnn_lstm_ctc_ocr_b200 import ctc


def _required_frames(label):
    """tf.nn.ctc_loss: a target needs len(label) frames plus one blank between every pair of equal neighbours."""
    return len(label) + sum(1 for i in range(1, len(label)) if label[i] == label[i - 1])


def _first_error(labels, seq_len, T, C):
    if len(seq_len) and (max(seq_len) > T or min(seq_len) < 0):
        return "sequence_length"
    flat = [v for l in labels for v in l]
    if flat and (max(flat) >= C - 1 or min(flat) < 0):
        return "label"
    for b, (l, s) in enumerate(zip(labels, seq_len)):
        if s > 0 and _required_frames(l) > s:
            return "Not enough time for target transition sequence (required: %d, available: %d)%d" % (_required_frames(l), s, b)
    return None


@pytest.mark.parametrize("seed", range(40))
def test_validate_matches_plain_loops(seed):
    rng = np.random.default_rng(seed)
    T, C, B = 24, 7, int(rng.integers(1, 12))
    tight = seed % 2 == 1   # odd seeds: lengths close to the limit, repeats likely (small alphabet)
    seq_len = [int(rng.integers(0 if seed % 5 == 0 else 1, T + 1)) for _ in range(B)]
    labels = []
    for s in seq_len:
        hi = max(1, (s if tight else s // 2)) + 1
        n = int(rng.integers(0, hi))
        labels.append([int(v) for v in rng.integers(0, C - 1, n)])
    if seed % 7 == 3:
        labels[0] = labels[0] + [C - 1]          # a blank inside the labels
    if seed % 11 == 5:
        seq_len[-1] = T + 1                       # longer than max_time
    flat, off, lengths, flat_host = ctc._labels_to_flat(labels, B, "cpu")
    assert off.tolist() == np.concatenate([[0], np.cumsum([len(l) for l in labels])]).tolist()
    assert flat.tolist() == [v for l in labels for v in l] and lengths == [len(l) for l in labels]
    want = _first_error(labels, seq_len, T, C)
    if want is None:
        ctc._validate_ctc(flat_host, lengths, np.asarray(seq_len), T, C, False)
        return
    with pytest.raises(ValueError) as e:
        ctc._validate_ctc(flat_host, lengths, np.asarray(seq_len), T, C, False)
    if want.startswith("Not enough"):
        assert str(e.value).startswith(want)
        ctc._validate_ctc(flat_host, lengths, np.asarray(seq_len), T, C, True)   # ignore_longer_outputs_than_inputs
    elif want == "label":
        assert "num_classes - 1" in str(e.value)
    else:
        assert "sequence_length" in str(e.value)


def test_label_forms_agree():
    labels = [[1, 2, 2], [], [0]]
    a = ctc._labels_to_flat(labels, 3, "cpu")
    idx = torch.tensor([[0, 0], [0, 1], [0, 2], [2, 0]])
    b = ctc._labels_to_flat((idx, torch.tensor([1, 2, 2, 0], dtype=torch.int32), torch.tensor([3, 3])), 3, "cpu")
    c = ctc._labels_to_flat((torch.tensor([1, 2, 2, 0], dtype=torch.int32), torch.tensor([3, 0, 1])), 3, "cpu")
    for x in (b, c):
        assert x[0].tolist() == a[0].tolist() and x[1].tolist() == a[1].tolist() and list(x[2]) == list(a[2])
    with pytest.raises(ValueError):
        ctc._labels_to_flat(labels, 4, "cpu")
    with pytest.raises(ValueError):   # rows out of order
        ctc._labels_to_flat((torch.tensor([[1, 0], [0, 0]]), torch.tensor([1, 2], dtype=torch.int32), torch.tensor([2, 1])), 2, "cpu")
