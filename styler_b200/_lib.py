"""ctypes binding of libstyler_b200.so (C ABI declared in include/styler_b200.h).

There is NO CPU fallback: if the shared library is missing the import of any op fails loudly with the build
instruction.  The library is built in-tree by `__graft_entry__.build()` / `make -C styler_b200/csrc`.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstyler_b200.so")

F32, BF16, F16 = 0, 1, 2
IMPL_AUTO, IMPL_SIMT, IMPL_TC = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_TANH, ACT_LRELU = 0, 1, 2, 3

c_i32, c_i64, c_f32, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p


class Conv1dArgs(ctypes.Structure):
    _fields_ = [
        ("x", c_vp), ("x_bstride", c_i64), ("x_ld", c_i32),
        ("B", c_i32), ("T", c_i32), ("Cin", c_i32),
        ("w", c_vp), ("N", c_i32), ("KS", c_i32), ("pad", c_i32),
        ("bias", c_vp), ("act", c_i32),
        ("residual", c_vp), ("r_bstride", c_i64), ("r_ld", c_i32), ("residual_is_f32", c_i32),
        ("ln_gamma", c_vp), ("ln_beta", c_vp), ("ln_eps", c_f32), ("act2", c_i32),
        ("lens", c_vp),
        ("dot_w", c_vp), ("dot_b", c_f32), ("dot_out", c_vp),
        ("out", c_vp), ("o_bstride", c_i64), ("o_ld", c_i32),
        ("out_f32", c_vp), ("of_bstride", c_i64), ("of_ld", c_i32),
        ("vt", c_vp), ("vt_col0", c_i32), ("vt_bstride", c_i64), ("vt_ld", c_i32),
        ("dtype", c_i32), ("impl", c_i32),
        ("dilation", c_i32), ("act_slope", c_f32), ("residual_inv_lrelu", c_i32),
        ("out2_f32", c_vp), ("gn_partial", c_vp),
    ]


class FftWeights(ctypes.Structure):
    _fields_ = [
        ("d_model", c_i32), ("d_inner", c_i32), ("n_head", c_i32),
        ("wqkv", c_vp), ("bqkv", c_vp),
        ("wfc", c_vp), ("bfc", c_vp),
        ("ln1_gamma", c_vp), ("ln1_beta", c_vp),
        ("w1", c_vp), ("b1", c_vp), ("ks1", c_i32),
        ("w2", c_vp), ("b2", c_vp), ("ks2", c_i32),
        ("ln2_gamma", c_vp), ("ln2_beta", c_vp),
        ("ln_eps", c_f32),
    ]


class PredictorWeights(ctypes.Structure):
    _fields_ = [
        ("c_in", c_i32), ("channels", c_i32), ("ks", c_i32),
        ("w1", c_vp), ("b1", c_vp), ("ln1_gamma", c_vp), ("ln1_beta", c_vp),
        ("w2", c_vp), ("b2", c_vp), ("ln2_gamma", c_vp), ("ln2_beta", c_vp),
        ("lin_w", c_vp), ("lin_b", c_f32), ("ln_eps", c_f32),
    ]


class PostnetWeights(ctypes.Structure):
    _fields_ = [
        ("n_layers", c_i32), ("n_mel", c_i32), ("channels", c_i32), ("ks", c_i32),
        ("w", c_vp * 8), ("b", c_vp * 8),
    ]


class DecoderWeights(ctypes.Structure):
    _fields_ = [
        ("n_layers", c_i32), ("layers", ctypes.POINTER(FftWeights)),
        ("mel_w", c_vp), ("mel_b", c_vp), ("n_mel", c_i32),
        ("postnet", ctypes.POINTER(PostnetWeights)),
    ]


# name -> argtypes (restype int unless noted); mirrors include/styler_b200.h one to one
_SIGNATURES = {
    "styler_conv1d_fwd": [ctypes.POINTER(Conv1dArgs), c_vp],
    "styler_fftblock_fwd": [ctypes.POINTER(FftWeights), c_vp, c_i64, c_i32, c_vp, c_i64, c_i32, c_vp, c_i32, c_i32, c_i32,
                            c_i32, c_vp, c_i64, c_vp],
    "styler_fftblock_workspace_bytes": [c_i32, c_i32, c_i32, c_i32, c_i32],
    "styler_debug_ffn1_timing": [c_i32, c_i32],
    "styler_debug_ffn1_timing_read": [c_vp, c_vp, c_vp, c_vp],
    "styler_lrelu_mean_fwd": [c_vp, c_vp, c_vp, c_f32, c_f32, c_vp, c_i64, c_i32, c_vp],
    "styler_attention_fwd": [c_vp, c_i64, c_i32, c_vp, c_i64, c_i32, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32,
                             c_i32, c_i32, c_vp],
    "styler_embed_pos_fwd": [c_vp, c_vp, c_i32, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp],
    "styler_add_fwd": [c_vp, c_i64, c_i32, c_vp, c_i64, c_i32, c_vp, c_i32, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32,
                       c_i32, c_i32, c_vp],
    "styler_cast_fwd": [c_vp, c_vp, c_i64, c_i32, c_vp],
    "styler_quantize_index_fwd": [c_vp, c_vp, c_i64, c_vp],
    "styler_onehot_conv_fwd": [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp],
    "styler_groupnorm_relu_fwd": [c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_f32, c_i32, c_vp],
    "styler_groupnorm_relu_partial_fwd": [c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_i32, c_vp, c_i32, c_i32, c_i32, c_f32, c_i32, c_vp],
    "styler_mel_calibrator_fwd": [c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32,
                                  c_i32, c_vp],
    "styler_gn_calibrator_fwd": [c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32,
                                 c_i32, c_f32, c_i32, c_vp],
    "styler_bilstm_layer_fwd": [c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp],
    "styler_classifier_tail_fwd": [c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp],
    "styler_duration_round_fwd": [c_vp, c_vp, c_i64, c_f32, c_f32, c_vp],
    "styler_length_regulator_fwd": [c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_i32, c_i32,
                                    c_i32, c_i32, c_i32, c_vp],
    "styler_bucket_embed_sum_fwd": [c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_f32, c_f32, c_vp, c_vp, c_i32, c_vp,
                                    c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32,
                                    c_i32, c_vp],
    "styler_stft_mel_fwd": [c_vp, c_i32, c_i32, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp],
    "styler_stft_mel_ex_fwd": [c_vp, c_i32, c_i32, c_vp, c_i32, c_vp, c_vp, c_vp, c_f32, c_i32, c_vp, c_i32, c_vp, c_f32,
                               c_f32, c_vp, c_vp],
    "styler_f0_norm_fwd": [c_vp, c_vp, c_vp, c_i32, c_i32, c_vp],
    "styler_debug_set_phase_buffer": [c_vp, c_i32],
    "styler_debug_trace": [c_i32],
    "styler_debug_trace_dump": [ctypes.c_char_p, c_i64],
    "styler_set_tuning": [ctypes.c_char_p, c_i32],
    "styler_predictor_workspace_bytes": [c_i32, c_i32, c_i32, c_i32],
    "styler_predictor_fwd": [ctypes.POINTER(PredictorWeights), c_vp, c_i64, c_i32, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp,
                             c_i64, c_vp],
    "styler_postnet_workspace_bytes": [c_i32, c_i32, c_i32, c_i32],
    "styler_postnet_fwd": [ctypes.POINTER(PostnetWeights), c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_i64, c_vp],
    "styler_decoder_workspace_bytes": [ctypes.POINTER(DecoderWeights), c_i32, c_i32, c_i32],
    "styler_decoder_fwd": [ctypes.POINTER(DecoderWeights), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32,
                           c_vp, c_i64, c_vp],
    "styler_loss_workspace_bytes": [],
    "styler_loss_fwd": [c_vp] * 15 + [c_i32, c_i32, c_i32, c_i32, c_vp, c_i64, c_vp, c_vp],
    "styler_peer_alloc": [c_i64, ctypes.POINTER(c_vp), c_vp],
    "styler_peer_open": [c_vp, ctypes.POINTER(c_vp)],
    "styler_peer_close": [c_vp],
    "styler_peer_free": [c_vp],
    "styler_peer_signal": [c_vp, ctypes.c_uint64, c_vp],
    "styler_peer_wait": [c_vp, c_i32, c_i64, ctypes.c_uint64, c_vp],
}
EXPORTED_SYMBOLS = sorted(list(_SIGNATURES) + ["styler_version", "styler_last_error", "styler_launch_count"])

_lib = None


def lib():
    """Load (once) and return the ctypes handle.  Raises if the CUDA library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            "styler_b200: %s is missing -- there is no CPU fallback.  Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C styler_b200/csrc`." % LIB_PATH)
    h = ctypes.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(h, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    h.styler_version.restype = ctypes.c_int
    h.styler_last_error.restype = ctypes.c_char_p
    h.styler_launch_count.restype = ctypes.c_int64
    h.styler_fftblock_workspace_bytes.restype = ctypes.c_int64
    h.styler_debug_trace_dump.restype = ctypes.c_int64
    for name in ("styler_predictor_workspace_bytes", "styler_postnet_workspace_bytes", "styler_decoder_workspace_bytes",
                 "styler_loss_workspace_bytes"):
        getattr(h, name).restype = ctypes.c_int64
    _lib = h
    return h


def check(rc, what=""):
    if rc != 0:
        msg = lib().styler_last_error()
        raise RuntimeError("styler_b200 %s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else "?"))


def set_tuning(name, value):
    """Flip one of the library's A/B switches at run time (styler_set_tuning); value < 0 restores the default."""
    check(lib().styler_set_tuning(name.encode(), int(value)), "set_tuning")


def launch_count():
    return int(lib().styler_launch_count())


def dtype_code(dt):
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    if dt == torch.float16:
        return F16
    raise TypeError("styler_b200: unsupported activation dtype %s" % dt)


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("styler_b200: expected CUDA tensors (the product path has no CPU implementation)")
