"""ctypes binding of libocr_b200.so (the C ABI declared in include/ocr_b200.h).

There is no CPU fallback and no alternative backend: if the shared library is missing or a call
fails, this module raises.  Build it with `python -m cnn_lstm_ctc_ocr_b200.build`.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libocr_b200.so")

OCR_OK = 0
_c = ctypes
_vp, _i, _f, _sz, _ll = _c.c_void_p, _c.c_int, _c.c_float, _c.c_size_t, _c.c_longlong

# name -> (restype, argtypes); mirrors include/ocr_b200.h one to one
SIGNATURES = {
    "ocr_last_error": (_c.c_char_p, []),
    "ocr_version": (_c.c_char_p, []),
    "ocr_launch_count": (_c.c_uint64, []),
    "ocr_ctc_loss_workspace_bytes": (_i, [_i, _i, _i, _i, _c.POINTER(_sz)]),
    "ocr_ctc_loss": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _f, _vp, _sz, _vp]),
    "ocr_ctc_loss_set_path": (_i, [_i]),
    "ocr_debug_ctc_timeline": (_i, [_vp]),
    "ocr_debug_ctc_group": (_i, [_i]),
    "ocr_debug_ctc_prefetch": (_i, [_i]),
    "ocr_debug_ctc_pdl": (_i, [_i]),
    "ocr_debug_ctc_inline_redo": (_i, [_i]),
    "ocr_debug_ctc_speculate": (_i, [_i]),
    "ocr_debug_ctc_stream_nbuf": (_i, [_i]),
    "ocr_ctc_greedy_decode": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "ocr_ctc_beam_search_workspace_bytes": (_i, [_i, _i, _i, _i, _c.POINTER(_sz)]),
    "ocr_debug_beam_path": (_i, [_i]),
    "ocr_debug_beam_profile": (_i, [_vp, _i]),
    "ocr_ctc_beam_search": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ocr_gemm_tf32": (_i, [_vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ocr_float_to_half": (_i, [_vp, _vp, _ll, _vp]),
    "ocr_gemm_f16": (_i, [_vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ocr_debug_proj_f16": (_i, [_i]),
    "ocr_preprocess_train": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "ocr_conv1_3x3_valid": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "ocr_im2col3x3_same": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "ocr_rows_max_to_seq": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "ocr_conv3x3_same": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp]),
    "ocr_maxpool": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "ocr_conv_set_path": (_i, [_i]),
    "ocr_conv3x3_pool_fused": (_i, [_i, _i, _i, _i, _i, _i]),
    "ocr_conv3x3_same_pool": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "ocr_debug_conv_tma_store": (_i, [_i]),
    "ocr_birnn_workspace_bytes": (_i, [_i, _i, _i, _i, _c.POINTER(_sz)]),
    "ocr_birnn_set_path": (_i, [_i]),
    "ocr_debug_lstm_timeline": (_i, [_vp]),
    "ocr_lstm_prepare_wh": (_i, [_vp, _i, _vp, _vp]),
    "ocr_birnn_layer": (_i, [_i, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ocr_edit_distance": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    # training step
    "ocr_transpose": (_i, [_vp, _ll, _i, _i, _vp, _ll, _ll, _vp]),
    "ocr_planar_pad_pitch": (_i, [_i]),
    "ocr_nhwc_to_planar_pad": (_i, [_vp, _i, _i, _i, _i, _vp, _ll, _i, _ll, _vp]),
    "ocr_gemm_wgrad_scratch_bytes": (_i, [_i, _i, _ll, _i, _c.POINTER(_sz)]),
    "ocr_gemm_tf32_wgrad": (_i, [_vp, _ll, _vp, _ll, _vp, _i, _ll, _i, _i, _ll, _i, _c.POINTER(_c.c_int32), _c.POINTER(_c.c_int32), _ll, _vp, _sz, _vp]),
    "ocr_planar_pad_pitch32": (_i, [_i]),
    "ocr_nhwc_to_planar_blocked": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _vp]),
    "ocr_gemm_tf32_wgrad_blocked": (_i, [_vp, _ll, _vp, _ll, _vp, _i, _ll, _i, _i, _ll, _i, _c.POINTER(_c.c_int32), _c.POINTER(_c.c_int32), _vp, _sz, _vp]),
    "ocr_bn_batch_sums": (_i, [_vp, _ll, _i, _vp, _vp]),
    "ocr_bn_finalize": (_i, [_vp, _ll, _i, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    "ocr_bn_relu_apply": (_i, [_vp, _ll, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ocr_bn_relu_bwd_sums": (_i, [_vp, _vp, _ll, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ocr_bn_relu_apply_pool": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "ocr_bn_relu_apply_pool_arg": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "ocr_bn_relu_bwd_sums_pool": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ocr_bn_relu_bwd_apply_bias_pool": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _ll, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ocr_bn_relu_bwd_apply": (_i, [_vp, _vp, _ll, _ll, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ocr_bn_relu_bwd_apply_bias": (_i, [_vp, _vp, _ll, _ll, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ocr_copy_2d": (_i, [_vp, _ll, _vp, _ll, _ll, _ll, _vp]),
    "ocr_relu_bwd_bias": (_i, [_vp, _vp, _ll, _i, _vp, _vp, _vp, _vp]),
    "ocr_colsum": (_i, [_vp, _ll, _i, _i, _vp, _vp, _vp]),
    "ocr_relu_bwd": (_i, [_vp, _vp, _ll, _vp, _vp]),
    "ocr_maxpool_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "ocr_rows_max_to_seq_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "ocr_conv1_wgrad": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp]),
    "ocr_conv_filter_layouts": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "ocr_adam_step": (_i, [_vp, _vp, _vp, _vp, _ll, _f, _vp, _f, _f, _f, _f, _vp]),
    "ocr_debug_bptt_pdl": (_i, [_i]),
    "ocr_debug_gemm_tma_store": (_i, [_i]),
    "ocr_debug_bptt_copies": (_i, [_i]),
    "ocr_debug_lstm_operands": (_i, [_i]),
    "ocr_birnn_lstm_train_workspace_bytes": (_i, [_i, _i, _i, _c.POINTER(_sz)]),
    "ocr_birnn_lstm_train_fwd": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ocr_birnn_lstm_bwd": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ocr_birnn_gru_train_workspace_bytes": (_i, [_i, _i, _i, _c.POINTER(_sz)]),
    "ocr_birnn_gru_train_fwd": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ocr_birnn_gru_bwd": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
}

_lib = None


class OcrLibraryError(RuntimeError):
    pass


def load():
    """Load libocr_b200.so (no CUDA call is made by loading)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise OcrLibraryError(
            "libocr_b200.so is not built (%s). Run `python -m cnn_lstm_ctc_ocr_b200.build`. "
            "This package has no CPU or PyTorch fallback." % SO_PATH)
    lib = ctypes.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != OCR_OK:
        msg = load().ocr_last_error().decode("utf-8", "replace")
        raise OcrLibraryError("%s failed (code %d): %s" % (what, rc, msg))


def launch_count():
    return int(load().ocr_launch_count())


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream_handle():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise OcrLibraryError("cnn_lstm_ctc_ocr_b200 operates on CUDA tensors only (got a %s tensor); "
                                  "there is no CPU path" % t.device)
