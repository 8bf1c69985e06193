"""The label alphabet is part of the interface (mjsynth.py:23-26, used by validate.py:126-129 and server.py:136-138):
it must equal the reference's constant byte for byte."""
import importlib.util
import os

from cnn_lstm_ctc_ocr_b200 import model

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/weinman/mjsynth.py"


def _golden():
    return open(os.path.join(HERE, "golden", "out_charset.txt"), encoding="utf-8").read()


def test_out_charset_equals_the_reference_constant():
    gold = _golden()
    if os.path.exists(REF):    # this container: read the constant out of the reference file itself
        spec = importlib.util.spec_from_file_location("make_out_charset", os.path.join(HERE, "golden", "make_out_charset.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        assert mod.read_reference_charset(REF) == gold
    assert model.out_charset == gold
    assert model.num_classes() == 95
    assert model.out_charset.index("A") == 0 and model.out_charset.index("a") == 26 and model.out_charset.index("0") == 52
    assert model.get_string([0, 26, 52]) == "Aa0"
    assert model.get_string(range(95)) == gold


def test_get_input_placeholder_contract():
    """validate._get_input (validate.py:71-78): uint8 [bucket_size,32,None,1] and int32 [bucket_size]; a feed of another dtype
    or static shape is refused the way TensorFlow refuses it (ValueError); the width axis is free."""
    import numpy as np
    import pytest
    import torch
    from cnn_lstm_ctc_ocr_b200 import model
    ph = model.get_input(4)
    assert ph.image_shape == (4, 32, None, 1) and ph.width_shape == (4,)
    for w in (40, 1024):
        img, wid = ph.feed(np.zeros((4, 32, w, 1), np.uint8), np.full(4, w, np.int32))
        assert img.dtype == torch.uint8 and wid.dtype == torch.int32 and tuple(img.shape) == (4, 32, w, 1)
    with pytest.raises(ValueError):
        ph.feed(np.zeros((4, 32, 40, 1), np.float32), np.full(4, 40, np.int32))      # preprocessed floats are not the placeholder's dtype
    with pytest.raises(ValueError):
        ph.feed(np.zeros((3, 32, 40, 1), np.uint8), np.full(3, 40, np.int32))        # wrong bucket size
    with pytest.raises(ValueError):
        ph.feed(np.zeros((4, 31, 40, 1), np.uint8), np.full(4, 40, np.int32))        # height is fixed at 32
    with pytest.raises(ValueError):
        ph.feed(np.zeros((4, 32, 40, 1), np.uint8), np.full((4, 1), 40, np.int32))   # width is a vector
