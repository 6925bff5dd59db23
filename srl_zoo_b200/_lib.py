"""ctypes binding of libsrlz.so (include/srlz.h).  The product path has NO fallback: if the CUDA library is
missing or fails to load, importing this module raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libsrlz.so")

MAX_BATCH = 2048   # SRLZ_MAX_BATCH (include/srlz.h): images per model call

c_float_p = C.POINTER(C.c_float)
VP = C.c_void_p


class SrlzBn(C.Structure):
    _fields_ = [("weight", VP), ("bias", VP), ("running_mean", VP), ("running_var", VP), ("num_batches_tracked", VP)]


class SrlzNet(C.Structure):
    _fields_ = [("is_vae", C.c_int32), ("state_dim", C.c_int32),
                ("enc_w", VP * 3), ("enc_bn", SrlzBn * 3),
                ("dec_w", VP * 5), ("dec_b", VP * 5), ("dec_bn", SrlzBn * 4),
                ("fc_enc_w", VP * 2), ("fc_enc_b", VP * 2), ("fc_dec_w", VP), ("fc_dec_b", VP)]


class SrlzNetGrads(C.Structure):
    _fields_ = [("enc_w", VP * 3), ("enc_bn_w", VP * 3), ("enc_bn_b", VP * 3),
                ("dec_w", VP * 5), ("dec_b", VP * 5), ("dec_bn_w", VP * 4), ("dec_bn_b", VP * 4),
                ("fc_enc_w", VP * 2), ("fc_enc_b", VP * 2), ("fc_dec_w", VP), ("fc_dec_b", VP)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libsrlz.so not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "-- there is no CPU / PyTorch fallback for the hot path" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.srlz_version.restype = C.c_int
    lib.srlz_last_error.restype = C.c_char_p
    for name in ("srlz_pack_floats",):
        getattr(lib, name).restype = C.c_size_t
        getattr(lib, name).argtypes = [C.c_int, C.c_int]
    for name in ("srlz_saved_bytes", "srlz_workspace_bytes"):
        getattr(lib, name).restype = C.c_size_t
        getattr(lib, name).argtypes = [C.c_int, C.c_int, C.c_int]
    lib.srlz_saved_layout.restype = C.c_int
    lib.srlz_saved_layout.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.c_int]
    lib.srlz_saved_names.restype = C.c_char_p
    lib.srlz_pack_weights.argtypes = [C.POINTER(SrlzNet), VP, VP]
    lib.srlz_forward.argtypes = [C.POINTER(SrlzNet), VP, VP, VP, VP, C.c_int, C.c_int, VP, VP, VP, VP, VP, VP, VP, VP]
    lib.srlz_replay_running_stats.argtypes = [C.POINTER(SrlzNet), C.c_int, VP, VP]
    lib.srlz_backward.argtypes = [C.POINTER(SrlzNet), VP, C.POINTER(SrlzNetGrads), C.c_int, VP, VP, VP, C.c_int, C.c_int,
                                  C.c_int, VP, VP, VP, C.c_float, VP, VP, C.c_float, VP, VP, VP]
    lib.srlz_heads.argtypes = [VP, VP, VP, C.c_int, C.c_int, C.c_int, C.c_int, VP, VP, VP, VP, C.c_float, C.c_float, VP, VP,
                               VP, VP, VP, VP, VP, C.c_int, VP, VP]
    lib.srlz_heads_workspace_bytes.restype = C.c_size_t
    lib.srlz_heads_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.srlz_kl.argtypes = [VP, VP, C.c_int, VP, VP, VP]
    lib.srlz_kl_grad.argtypes = [VP, VP, C.c_int, C.c_float, VP, VP, VP]
    lib.srlz_cross_entropy.argtypes = [VP, VP, C.c_int, C.c_int, VP, VP, VP, VP]
    lib.srlz_launch_count.restype = C.c_longlong
    lib.srlz_launch_count.argtypes = []
    lib.srlz_prof_enable.argtypes = [C.c_int]
    lib.srlz_prof_enable.restype = None
    lib.srlz_prof_report.argtypes = [C.c_char_p, C.c_int]
    lib.srlz_prof_report.restype = C.c_int
    lib.srlz_preprocess_u8.argtypes = [VP, VP, C.c_int, VP]
    lib.srlz_eval_pack_floats.restype = C.c_size_t
    lib.srlz_eval_pack_floats.argtypes = [C.c_int]
    lib.srlz_eval_workspace_bytes.restype = C.c_size_t
    lib.srlz_eval_workspace_bytes.argtypes = [C.c_int, C.c_int]
    lib.srlz_eval_pack.argtypes = [C.POINTER(SrlzNet), VP, VP]
    lib.srlz_encode_eval.argtypes = [C.POINTER(SrlzNet), VP, VP, VP, C.c_int, VP, VP, VP]
    lib.srlz_decode.argtypes = [C.POINTER(SrlzNet), VP, VP, C.c_int, C.c_int, VP, VP, VP, VP]
    lib.srlz_decode_backward.argtypes = [C.POINTER(SrlzNet), VP, C.POINTER(SrlzNetGrads), C.c_int, C.c_int, C.c_int, VP, VP, VP, VP, VP]
    lib.srlz_relu.argtypes = [VP, VP, C.c_int64, VP]
    lib.srlz_relu_bwd.argtypes = [VP, VP, VP, C.c_int64, VP]
    lib.srlz_colmask.argtypes = [VP, VP, VP, C.c_int, C.c_int, VP]
    lib.srlz_cat_cols.argtypes = [VP, C.c_int, VP, C.c_int, VP, VP, C.c_int, VP]
    lib.srlz_reparam.argtypes = [VP, VP, VP, VP, C.c_int, VP]
    lib.srlz_reparam_bwd.argtypes = [VP, VP, VP, VP, VP, C.c_int, VP]
    lib.srlz_sse.argtypes = [VP, VP, C.c_int64, C.c_float, VP, VP, VP]
    lib.srlz_mse_grad.argtypes = [VP, VP, C.c_int64, C.c_float, VP, VP]
    lib.srlz_adam_step.argtypes = [VP, VP, VP, VP, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, VP]
    lib.srlz_op_conv64.argtypes = [VP, VP, VP, VP, VP, VP] + [C.c_int] * 9 + [VP, C.POINTER(C.c_int), VP]
    lib.srlz_op_wgrad64.argtypes = [VP, VP, VP, VP, VP] + [C.c_int] * 8 + [VP, VP]
    lib.srlz_op_wgrad64_workspace_bytes.restype = C.c_size_t
    lib.srlz_op_wgrad64_workspace_bytes.argtypes = [C.c_int] * 8
    lib.srlz_op_pack_conv_w_bf16.argtypes = [VP, VP, C.c_int, VP]
    lib.srlz_op_sgemm.argtypes = [VP, C.c_int64, C.c_int64, VP, C.c_int64, C.c_int64, VP, C.c_int64, C.c_int64, VP,
                                  C.c_int, C.c_int, C.c_int, C.c_int, VP]
    lib.srlz_op_pack_conv_w.argtypes = [VP, VP, VP, C.c_int, C.c_int, VP]
    lib.srlz_op_bn_relu_pool.argtypes = [VP, VP, VP, VP, VP] + [C.c_int] * 6 + [VP]
    lib.srlz_op_layer_workspace_bytes.restype = C.c_size_t
    lib.srlz_op_layer_workspace_bytes.argtypes = []
    lib.srlz_op_enc0_fwd.argtypes = [VP, VP, VP, VP, VP, C.POINTER(C.c_int), C.c_int, VP, VP]
    lib.srlz_op_enc0_wgrad.argtypes = [VP, VP, VP, VP, C.c_int, VP, VP]
    lib.srlz_op_dec12_fwd.argtypes = [VP, VP, VP, VP, VP, VP, VP, VP, C.c_int, VP, VP]
    lib.srlz_op_dec12_bwd.argtypes = [VP] * 9 + [C.c_float] + [VP] * 4 + [C.POINTER(C.c_int), C.c_int, VP, VP]
    for name in ("srlz_pack_weights", "srlz_forward", "srlz_replay_running_stats", "srlz_backward", "srlz_heads",
                 "srlz_sse", "srlz_mse_grad", "srlz_adam_step", "srlz_op_conv64", "srlz_op_wgrad64",
                 "srlz_op_pack_conv_w", "srlz_op_sgemm", "srlz_op_pack_conv_w_bf16", "srlz_kl", "srlz_kl_grad",
                 "srlz_cross_entropy", "srlz_preprocess_u8", "srlz_eval_pack", "srlz_encode_eval", "srlz_decode", "srlz_decode_backward", "srlz_relu", "srlz_relu_bwd",
                 "srlz_colmask", "srlz_cat_cols", "srlz_reparam", "srlz_reparam_bwd", "srlz_op_enc0_fwd", "srlz_op_enc0_wgrad", "srlz_op_dec12_fwd", "srlz_op_dec12_bwd", "srlz_op_bn_relu_pool"):
        getattr(lib, name).restype = C.c_int
    return lib


class _LazyLib:
    """`lib.srlz_xxx` loads libsrlz.so on first use (so the pure-host modules -- checkpoint, fold, parallel, occlusion --
    import on a machine without the built library); a missing library still raises, there is no fallback."""
    _real = None

    def __getattr__(self, name):
        real = _LazyLib._real
        if real is None:
            real = _LazyLib._real = _load()
        return getattr(real, name)


lib = _LazyLib()

EXPORTED = ["srlz_version", "srlz_last_error", "srlz_pack_floats", "srlz_saved_bytes", "srlz_workspace_bytes",
            "srlz_saved_layout", "srlz_saved_names", "srlz_pack_weights", "srlz_forward", "srlz_replay_running_stats",
            "srlz_backward", "srlz_heads", "srlz_heads_workspace_bytes", "srlz_sse", "srlz_mse_grad", "srlz_adam_step",
            "srlz_op_conv64", "srlz_op_wgrad64", "srlz_op_wgrad64_workspace_bytes", "srlz_op_pack_conv_w",
            "srlz_op_sgemm", "srlz_kl", "srlz_kl_grad", "srlz_cross_entropy", "srlz_prof_enable", "srlz_launch_count",
            "srlz_op_pack_conv_w_bf16", "srlz_prof_report", "srlz_op_layer_workspace_bytes", "srlz_op_enc0_fwd",
            "srlz_op_enc0_wgrad", "srlz_op_dec12_fwd", "srlz_op_dec12_bwd", "srlz_preprocess_u8", "srlz_decode",
            "srlz_decode_backward", "srlz_relu", "srlz_relu_bwd", "srlz_colmask", "srlz_cat_cols", "srlz_reparam", "srlz_reparam_bwd",
            "srlz_eval_pack_floats", "srlz_eval_workspace_bytes", "srlz_eval_pack", "srlz_encode_eval", "srlz_op_bn_relu_pool"]


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError("libsrlz %s failed (code %d): %s" % (what, rc, lib.srlz_last_error().decode()))


def ptr(t):
    """device pointer of a torch tensor (None -> NULL)"""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def prof_enable(on=True):
    lib.srlz_prof_enable(int(on))


def prof_report():
    """-> {tag: (launch groups, total ms)} for the call sites timed since prof_enable(True)"""
    buf = C.create_string_buffer(16384)
    check(lib.srlz_prof_report(buf, len(buf)), "prof_report")
    out = {}
    for line in buf.value.decode().splitlines():
        tag, cnt, ms = line.split()
        out[tag] = (int(cnt), float(ms))
    return out
