"""Shared synthetic-input generators."""
import numpy as np


def make_labels(rng, B, seq_len, max_len=16, num_labels=62, repeat_p=0.15):
    """Label length U{1..min(max_len, feasible)}, ids U{0..num_labels-1}, ~15% forced adjacent repeats."""
    labels = []
    for b in range(B):
        T = int(seq_len[b])
        L = int(rng.integers(1, max(2, min(max_len, T) + 1)))
        while True:
            l = rng.integers(0, num_labels, L)
            for i in range(1, L):
                if rng.random() < repeat_p:
                    l[i] = l[i - 1]
            need = L + int(np.sum(l[1:] == l[:-1]))
            if need <= T:
                break
            L = max(1, L - 1)
        labels.append([int(v) for v in l])
    return labels


def cfg2_inputs(seed=1, T=64, B=256, C=63, ragged=True, relu=False, scale=1.0):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((T, B, C)) * scale).astype(np.float32)
    if relu:
        x = np.maximum(x, 0.0)
    seq_len = rng.integers(T // 2, T + 1, B).astype(np.int32) if ragged else np.full(B, T, np.int32)
    labels = make_labels(rng, B, seq_len, 16, C - 1)
    return x, labels, seq_len
