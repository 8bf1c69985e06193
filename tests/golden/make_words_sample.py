"""Regenerates tests/golden/words-sample.tfrecord: the first 48 records of the reference's own fixture
data/test/words-000.tfrecord, re-framed by mjsynth.write_tfrecord (payloads byte-identical).  /root/reference does not
exist on the GPU box; this slice travels instead."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cnn_lstm_ctc_ocr_b200 import mjsynth  # noqa: E402

if __name__ == "__main__":
    src = "/root/reference/data/test/words-000.tfrecord"
    recs = []
    for i, p in enumerate(mjsynth.read_tfrecord(src, verify=True)):
        if i == 48:
            break
        recs.append(p)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "words-sample.tfrecord")
    mjsynth.write_tfrecord(out, recs)
    # framing check: the slice is a byte prefix of the source file
    assert open(out, "rb").read() == open(src, "rb").read()[:os.path.getsize(out)]
    print("wrote", out, os.path.getsize(out), "bytes")
