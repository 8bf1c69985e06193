"""cnn_lstm_ctc_ocr_b200 -- B200-native hot path of tgialoimtr/cnn_lstm_ctc_ocr.

The weinman CNN -> BiLSTM/BiGRU -> CTC line recognizer (reference: src/weinman/model.py,
model_bu.py, validate.py, test.py, train.py, src/processing/server.py) behind the reference's own
Python names, computed by hand-written sm_100a CUDA kernels in libocr_b200.so (C ABI:
include/ocr_b200.h).  PyTorch is used for device memory, streams and torch.distributed only.
There is no CPU path: importing works anywhere, calling an op needs the built library and a GPU.
"""
from . import _lib  # noqa: F401
from . import ctc  # noqa: F401
from . import model  # noqa: F401

__version__ = "0.1.0"
