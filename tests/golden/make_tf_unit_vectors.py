"""Writes tests/golden/tf_unit_vectors.npz.

The reference repository holds no golden vectors for its CTC path, and
TensorFlow cannot be installed here.  The vectors below are the inputs and expected outputs of
upstream TensorFlow's own unit tests for the three ops the reference calls
(python/kernel_tests/ctc_loss_op_test.py::testBasic, ctc_decoder_ops_test.py::
testCTCGreedyDecoder and ::testCTCDecoderBeamSearch), transcribed from upstream.  They are
self-validating: the oracle reproduces every printed digit from the probability matrices, which a
mis-transcription could not do.
"""
import os

import numpy as np

loss_p0 = np.array(
    [[0.633766, 0.221185, 0.0917319, 0.0129757, 0.0142857, 0.0260553],
     [0.111121, 0.588392, 0.278779, 0.0055756, 0.00569609, 0.010436],
     [0.0357786, 0.633813, 0.321418, 0.00249248, 0.00272882, 0.0037688],
     [0.0663296, 0.643849, 0.280111, 0.00283995, 0.0035545, 0.00331533],
     [0.458235, 0.396634, 0.123377, 0.00648837, 0.00903441, 0.00623107]], np.float32)
loss_g0 = np.array(
    [[-0.366234, 0.221185, 0.0917319, 0.0129757, 0.0142857, 0.0260553],
     [0.111121, -0.411608, 0.278779, 0.0055756, 0.00569609, 0.010436],
     [0.0357786, 0.633813, -0.678582, 0.00249248, 0.00272882, 0.0037688],
     [0.0663296, -0.356151, 0.280111, 0.00283995, 0.0035545, 0.00331533],
     [-0.541765, 0.396634, 0.123377, 0.00648837, 0.00903441, 0.00623107]], np.float32)
loss_p1 = np.array(
    [[0.30176, 0.28562, 0.0831517, 0.0862751, 0.0816851, 0.161508],
     [0.24082, 0.397533, 0.0557226, 0.0546814, 0.0557528, 0.19549],
     [0.230246, 0.450868, 0.0389607, 0.038309, 0.0391602, 0.202456],
     [0.280884, 0.429522, 0.0326593, 0.0339046, 0.0326856, 0.190345],
     [0.423286, 0.315517, 0.0338439, 0.0393744, 0.0339315, 0.154046]], np.float32)
loss_g1 = np.array(
    [[-0.69824, 0.28562, 0.0831517, 0.0862751, 0.0816851, 0.161508],
     [0.24082, -0.602467, 0.0557226, 0.0546814, 0.0557528, 0.19549],
     [0.230246, 0.450868, 0.0389607, 0.038309, 0.0391602, -0.797544],
     [0.280884, -0.570478, 0.0326593, 0.0339046, 0.0326856, 0.190345],
     [-0.576714, 0.315517, 0.0338439, 0.0393744, 0.0339315, 0.154046]], np.float32)
greedy_p0 = np.array([[1.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.4, 0.6], [0.0, 0.0, 0.4, 0.6],
                      [0.0, 0.9, 0.1, 0.0], [0.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0]], np.float32)
greedy_p1 = np.array([[0.1, 0.9, 0.0, 0.0], [0.0, 0.9, 0.1, 0.0], [0.0, 0.0, 0.1, 0.9],
                      [0.0, 0.9, 0.1, 0.1], [0.9, 0.1, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0]], np.float32)
beam_p = np.array(
    [[0.30999, 0.309938, 0.0679938, 0.0673362, 0.0708352, 0.173908],
     [0.215136, 0.439699, 0.0370931, 0.0393967, 0.0381581, 0.230517],
     [0.199959, 0.489485, 0.0233221, 0.0251417, 0.0233289, 0.238763],
     [0.279611, 0.452966, 0.0204795, 0.0209126, 0.0194803, 0.20655],
     [0.51286, 0.288951, 0.0243026, 0.0220788, 0.0219297, 0.129878],
     [0.155251, 0.164444, 0.173517, 0.176138, 0.169979, 0.160671]], np.float32)

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tf_unit_vectors.npz")
np.savez(out,
         loss_p0=loss_p0, loss_g0=loss_g0, loss_targets0=np.array([0, 1, 2, 1, 0]), loss_value0=np.float32(3.34211),
         loss_p1=loss_p1, loss_g1=loss_g1, loss_targets1=np.array([0, 1, 1, 0]), loss_value1=np.float32(5.42262),
         greedy_p0=greedy_p0, greedy_p1=greedy_p1, greedy_seq_len=np.array([4, 5]),
         greedy_decode0=np.array([0, 1]), greedy_decode1=np.array([1, 1, 0]),
         beam_p=beam_p, beam_seq_len=np.array([5]), beam_offset=np.float32(2.0),
         beam_decode0=np.array([1, 0]), beam_decode1=np.array([0, 1, 0]),
         # TF <= 1.x output (per-frame max subtraction): "negative log probabilities" 0.584855, 0.389139
         beam_logprob_vmax=np.array([0.584855, 0.389139], np.float32))
print("wrote", out)
