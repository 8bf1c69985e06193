"""Training-side input format (src/weinman/mjsynth.py:28-194): TFRecord framing, tf.train.Example parsing, the word-record
feature map, train-side preprocessing and the width-bucketed batcher -- against the reference's own fixture records."""
import os

import numpy as np
import pytest
import torch

from cnn_lstm_ctc_ocr_b200 import mjsynth, model

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
SAMPLE = os.path.join(GOLD, "words-sample.tfrecord")
REF_DATA = "/root/reference/data"


def test_crc32c_known_answers():
    assert mjsynth.crc32c(b"123456789") == 0xE3069283          # the CRC-32C check value
    assert mjsynth.crc32c(b"") == 0
    assert mjsynth.crc32c(bytes(32)) == 0x8A9136AA             # RFC 3720 B.4: 32 bytes of zeros


def test_sample_records_frame_and_parse():
    recs = list(mjsynth.read_tfrecord(SAMPLE, verify=True))
    assert len(recs) == 48
    for p in recs:
        image, width, labels, length, text, filename = mjsynth.read_word_record(p)
        assert image.dtype == np.uint8 and image.shape == (31, int(width[0]), 1)
        assert int(length[0]) == len(text) == len(labels)
        assert model.get_string(labels) == text.decode("ascii")       # the label ids index the reference's out_charset
        assert filename.endswith(b".jpg")
    f = mjsynth.parse_example(recs[0])
    assert sorted(f) == ["image/encoded", "image/filename", "image/height", "image/labels", "image/width", "text/length", "text/string"]
    assert f["image/height"] == [31] and f["text/string"] == [b"slinking"] and f["image/width"] == [130]


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="the reference's fixtures only exist in the build container")
def test_all_reference_fixture_records():
    n = {}
    for split in ("test", "val"):
        cnt = 0
        for p in mjsynth.read_tfrecord(os.path.join(REF_DATA, split, "words-000.tfrecord")):
            f = mjsynth.parse_example(p)
            assert model.get_string(f["image/labels"]) == f["text/string"][0].decode("ascii")
            assert f["text/length"] == [len(f["image/labels"])]
            cnt += 1
        n[split] = cnt
    assert n == {"test": 892, "val": 803}                                 # SURVEY.md 8(c) "Fixtures"


def test_truncated_and_corrupted_records_raise(tmp_path):
    raw = open(SAMPLE, "rb").read()
    bad = tmp_path / "bad.tfrecord"
    bad.write_bytes(raw[:1000])
    with pytest.raises(ValueError, match="truncated"):
        list(mjsynth.read_tfrecord(str(bad)))
    flipped = bytearray(raw)
    flipped[40] ^= 1
    bad.write_bytes(bytes(flipped))
    with pytest.raises(ValueError, match="corrupted"):
        list(mjsynth.read_tfrecord(str(bad), verify=True))


def test_example_round_trip(tmp_path):
    ex = mjsynth.make_example({"image/labels": [3, 0, 94, 300, -1], "text/string": [b"abc"], "image/width": [77], "score": [0.5, -2.0]})
    f = mjsynth.parse_example(ex)
    assert f["image/labels"] == [3, 0, 94, 300, -1] and f["text/string"] == [b"abc"] and f["image/width"] == [77] and f["score"] == [0.5, -2.0]
    path = tmp_path / "w.tfrecord"
    mjsynth.write_tfrecord(str(path), [ex, b"", ex])
    assert list(mjsynth.read_tfrecord(str(path), verify=True)) == [ex, b"", ex]


def test_train_side_preprocessing():
    img = np.arange(31 * 5, dtype=np.uint8).reshape(31, 5, 1)
    out = mjsynth.preprocess_image(img)
    assert out.dtype == np.float32 and out.shape == (32, 5, 1)
    assert (out[0] == out[1]).all()                                         # copy of the first row on top (mjsynth.py:190-192)
    assert (out[1:] == img.astype(np.float32) * np.float32(1 / 255.0) - np.float32(0.5)).all()
    assert out.min() >= -0.5 and out.max() <= 0.5


def test_bucketed_pipeline_contract():
    bounds = (32, 64, 96, 128, 160, 192, 224, 256)
    batches = list(mjsynth.bucketed_input_pipeline(GOLD, ["words-sample.tfrecord"], batch_size=4, boundaries=bounds, num_epochs=1))
    seen = 0
    for image, width, label, length, text, filename in batches:
        B = len(width)
        seen += B
        assert 1 <= B <= 4 and image.dtype == np.float32 and image.shape == (B, 32, int(width.max()), 1)      # padded to the batch maximum
        ks = np.searchsorted(bounds, width, side="right")
        assert (ks == ks[0]).all()                                          # one bucket per batch: boundaries[k-1] <= w < boundaries[k]
        for b in range(B):
            assert (image[b, :, width[b]:] == 0.0).all()                    # dynamic_pad pads the PREPROCESSED image with 0.0
            rows = label.indices[:, 0] == b
            assert model.get_string(label.values[rows].tolist()) == text[b].decode("ascii")
        assert label.values.dtype == torch.int32 and label.indices.dtype == torch.int64
        assert label.dense_shape.tolist() == [B, int(length.max())] and length.shape == (B, 1)
    assert seen == 48                                                       # num_epochs=1: the smaller final batches leave too
    full = [len(b[1]) for b in batches]
    assert full.count(4) >= 6 and any(n < 4 for n in full)
    # thresholds drop records (mjsynth._get_input_filter)
    kept = sum(len(b[1]) for b in mjsynth.bucketed_input_pipeline(GOLD, ["words-sample.tfrecord"], batch_size=4, num_epochs=1,
                                                                  width_threshold=100, length_threshold=6))
    recs = [mjsynth.read_word_record(p) for p in mjsynth.read_tfrecord(SAMPLE)]
    assert kept == sum(1 for r in recs if r[1][0] <= 100 and r[3][0] <= 6) and 0 < kept < 48
    # uint8 hand-out for the device-side preprocessing: raw rows, zero right-padding
    image, width, *_ = next(mjsynth.bucketed_input_pipeline(GOLD, ["words-sample.tfrecord"], batch_size=4, num_epochs=1, as_uint8=True))
    assert image.dtype == np.uint8 and image.shape[1] == 31


def test_threaded_pipeline_keeps_record_order():
    texts = [t for b in mjsynth.threaded_input_pipeline(GOLD, ["words-sample.tfrecord"], batch_size=5, num_epochs=1) for t in b[4]]
    assert texts == [mjsynth.read_word_record(p)[4] for p in mjsynth.read_tfrecord(SAMPLE)]
    it = mjsynth.threaded_input_pipeline(GOLD, ["words-sample.tfrecord"], batch_size=48, num_epochs=None)    # endless epochs
    a, b = next(it), next(it)
    assert a[4] == b[4]


@pytest.mark.gpu
def test_device_preprocessing_equals_the_host_pipeline_bit_for_bit():
    from cnn_lstm_ctc_ocr_b200 import train
    tr = train.Trainer(model.init_params(0, "lstm", (32, 32)), rnn_sizes=(32, 32))
    f32 = mjsynth.bucketed_input_pipeline(GOLD, ["words-sample.tfrecord"], batch_size=4, num_epochs=1)
    u8 = mjsynth.bucketed_input_pipeline(GOLD, ["words-sample.tfrecord"], batch_size=4, num_epochs=1, as_uint8=True)
    for (image, width, *_), (raw, width8, *_) in zip(f32, u8):
        assert (width == width8).all()
        got = tr.preprocess_train(torch.from_numpy(raw).cuda(), torch.from_numpy(width8).cuda())
        assert torch.equal(got.cpu(), torch.from_numpy(image))


@pytest.mark.gpu
def test_trainer_consumes_bucketed_batches():
    """One step per form of the same batch: float32 from the host pipeline (the reference's contract) and raw uint8 rows
    preprocessed on the device give the same losses; the loss falls over a few steps on the fixture words."""
    from cnn_lstm_ctc_ocr_b200 import train
    p = model.init_params(1, "lstm", (32, 32))
    batches = [b for b in mjsynth.bucketed_input_pipeline(GOLD, ["words-sample.tfrecord"], batch_size=4, num_epochs=1) if len(b[1]) == 4]
    raws = [b for b in mjsynth.bucketed_input_pipeline(GOLD, ["words-sample.tfrecord"], batch_size=4, num_epochs=1, as_uint8=True) if len(b[1]) == 4]
    image, width, label, *_ = batches[0]
    ta, tb = train.Trainer(p, rnn_sizes=(32, 32)), train.Trainer(p, rnn_sizes=(32, 32))
    la = ta.forward_backward(torch.from_numpy(image).cuda(), width, label)
    lb = tb.forward_backward(torch.from_numpy(raws[0][0]).cuda(), raws[0][1], raws[0][2])
    assert torch.equal(la, lb) and torch.equal(ta.grad, tb.grad)
    tr = train.Trainer(p, rnn_sizes=(32, 32), learning_rate=1e-3)
    first = last = None
    for epoch in range(6):
        tot = 0.0
        for image, width, label, *_ in batches:
            tot += float(tr.train_step(torch.from_numpy(image).cuda(), width, label))
        first = tot if first is None else first
        last = tot
    assert np.isfinite(last) and last < first
