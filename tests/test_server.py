"""Bucketing contract of the recognition step (server.py:17-57,104-143): CPU tier with a stub recognizer,
GPU tier end to end."""
import time

import numpy as np
import pytest
import torch

from cnn_lstm_ctc_ocr_b200 import server


class _Stub:
    """Records the batches the server hands to the graph."""

    def __init__(self):
        self.calls = []

    def recognize(self, images, widths):
        self.calls.append((tuple(images.shape), images.dtype, widths.clone()))
        return ["w%d" % int(w) for w in widths]


def _crop(w, val=200):
    return np.full((32, w), val, np.uint8)


def test_bucket_boundaries_padding_and_fillers():
    stub = _Stub()
    srv = server.LocalServer(stub, bucket_size=4, device="cpu")
    assert len(srv.buckets) == 31 and srv.buckets[0].widthrange == (32, 64) and srv.buckets[-1].widthrange == (992, 1024)
    for i, w in enumerate([33, 64, 65, 1024, 40]):
        srv.submit("c", i, _crop(w))
    with pytest.raises(ValueError):
        srv.submit("c", 9, _crop(32))       # (32, 64]: width 32 matches no bucket (reference: assert at server.py:114)
    with pytest.raises(ValueError):
        srv.submit("c", 9, _crop(1025))
    srv.flush()
    shapes = sorted(c[0] for c in stub.calls)
    assert shapes == [(4, 32, 64, 1), (4, 32, 96, 1), (4, 32, 1024, 1)]     # padded to the bucket's upper width, fixed batch
    first = [c for c in stub.calls if c[0][2] == 64][0]
    assert first[1] == torch.uint8
    assert first[2].tolist() == [33, 64, 40, 33]                            # filler rows reuse widths[0]
    got = srv.take("c")
    assert got == {0: "w33", 1: "w64", 2: "w65", 3: "w1024", 4: "w40"}       # fillers dropped


def test_multi_gpu_dispatch_deals_whole_batches():
    """shard=(rank, world): every process sees the same crops and runs the batches (bucket k, batch j) with
    (k + j) % world == rank.  Union of the shares == the one-process result; the padded pixels add up to the one-process
    figure (fillers only in a bucket's last batch), and the shares are balanced."""
    rng = np.random.default_rng(3)
    widths = rng.integers(64, 1025, 900)
    crops = [_crop(int(w), i % 251) for i, w in enumerate(widths)]
    one = server.LocalServer(_Stub(), bucket_size=8, device="cpu")
    ref = server.BatchLinePredictor(one).predict_batch("s", crops)
    for world in (2, 4, 8):
        union, padded, real, calls = {}, 0, 0, []
        for rank in range(world):
            stub = _Stub()
            srv = server.LocalServer(stub, bucket_size=8, device="cpu", shard=(rank, world))
            got = server.BatchLinePredictor(srv).predict_batch("s", crops)
            assert not (set(got) & set(union))
            union.update(got)
            padded += srv.padded_pixels
            real += srv.real_pixels
            calls.append(len(stub.calls))
            assert all(c[0][0] == 8 for c in stub.calls)          # fixed batch, as on one GPU
        assert union == ref
        assert (padded, real) == (one.padded_pixels, one.real_pixels)
        assert max(calls) - min(calls) <= 4


def test_release_rule_is_strictly_more_than_batchsize_or_age():
    stub = _Stub()
    srv = server.LocalServer(stub, bucket_size=2, bucket_max_time=0.05, device="cpu")
    srv.submit("c", 0, _crop(50)); srv.submit("c", 1, _crop(50))
    srv.poll()
    assert stub.calls == []                    # len == batchsize is not enough (server.py:47, strict '>')
    srv.submit("c", 2, _crop(50))
    srv.poll()
    assert len(stub.calls) == 1 and stub.calls[0][2].tolist() == [50, 50]
    time.sleep(0.08)
    srv.poll()                                 # the leftover crop leaves when it is older than bucket_max_time
    assert len(stub.calls) == 2 and stub.calls[1][2].tolist() == [50, 50]   # 1 real + 1 filler


def test_padding_value_is_uint8_zero():
    stub = _Stub()
    b = server.Bucket(1.0, 4, (64, 96))
    assert b.addImgToBucket("c", 0, 0.0, _crop(70, 255))
    infos, batch, widths = b.getBatch(force=True)
    assert batch.shape == (1, 32, 96, 1) and (batch[0, :, :70, 0] == 255).all() and (batch[0, :, 70:, 0] == 0).all()


@pytest.mark.gpu
def test_batch_line_predictor_end_to_end():
    from cnn_lstm_ctc_ocr_b200 import model
    from oracle import model_oracle as mo
    rng = np.random.default_rng(0)
    m = model.Model(mo.init_params(0, "gru", (512, 256)), cell_type="gru", rnn_sizes=(512, 256))
    srv = server.LocalServer(m, bucket_size=8)
    crops = [rng.integers(0, 256, (32, int(w))).astype(np.uint8) for w in rng.integers(33, 400, 20)]
    out = server.BatchLinePredictor(srv).predict_batch("page0", crops)
    assert sorted(out) == list(range(20)) and all(isinstance(t, str) for t in out.values())
    # a crop recognised alone in its bucket gives the same text as inside the mixed submission
    solo = server.BatchLinePredictor(server.LocalServer(m, bucket_size=8)).predict_batch("x", [crops[3]])
    assert solo[0] == out[3]
    assert 0 < srv.real_pixels <= srv.padded_pixels
