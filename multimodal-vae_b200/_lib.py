"""ctypes binding of libmvae_b200.so (the C ABI declared in include/mvae_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing it is built with nvcc, and
if that is impossible every op raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build

_lock = threading.Lock()
_lib = None

c_void_p, c_int, c_int64, c_float = C.c_void_p, C.c_int, C.c_int64, C.c_float


class MvaeError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("dtype", c_int), ("M", c_int), ("N", c_int), ("K", c_int),
        ("A", c_void_p), ("lda", c_int64), ("a_major", c_int),
        ("B", c_void_p), ("ldb", c_int64), ("b_major", c_int),
        ("C", c_void_p), ("ldc", c_int64), ("c_dtype", c_int),
        ("bias", c_void_p), ("accumulate", c_int),
        ("col_sum", c_void_p), ("col_sumsq", c_void_p), ("rows_per_group", c_int),
        ("block_n", c_int), ("split_k", c_int), ("stages", c_int),
        ("debug_times", c_void_p),
        ("x3_scratch", c_void_p), ("x3_scratch_bytes", c_int64),
        ("act", c_int), ("act_out", c_void_p), ("act_pre", c_void_p), ("ld_act_pre", c_int64),
        ("bce_target", c_void_p), ("ld_bce_target", c_int64), ("bce_target_rows", c_int), ("bce_scale", c_float * 4),
        ("bce_loss", c_void_p), ("bce_probs", c_void_p), ("bce_row_weight", c_void_p),
    ]


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Load (building first if needed) the shared library.  Raises if it cannot be had."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if _build.needs_build():
            _build.build()
        lib = C.CDLL(_build.LIB_PATH)
        lib.mvae_last_error.restype = C.c_char_p
        lib.mvae_launch_count.restype = C.c_longlong
        lib.mvae_poe_forward.argtypes = [c_int, c_int, c_float, c_int, c_int64, c_int] + [c_void_p] * 6
        lib.mvae_poe_backward.argtypes = [c_int, c_int, c_float, c_int, c_int64, c_int] + [c_void_p] * 8
        lib.mvae_mnist_workspace_offset.restype = C.c_longlong
        lib.mvae_mnist_workspace_offset.argtypes = [C.c_char_p, c_int, c_int, c_int]
        lib.mvae_adam_step.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float,
                                       c_float, c_float, c_void_p, c_float, c_int, c_void_p]
        lib.mvae_dp_reduce_adam.argtypes = [C.POINTER(DpReduceAdamArgs), c_void_p]
        _conv_argtypes(lib)
        _lib = lib
        return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().mvae_last_error().decode("utf-8", "replace")
        raise MvaeError("%s failed (rc=%d): %s" % (what or "mvae call", rc, msg))


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


c_uint64 = C.c_uint64


class TensorInfo(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("kind", c_int), ("ndim", c_int), ("shape", c_int64 * 2),
                ("offset", c_int64)]


class MnistSizeInfo(C.Structure):
    _fields_ = [("param_floats", c_int64), ("encoder_param_floats", c_int64), ("buffer_floats", c_int64),
                ("num_bn", c_int64), ("workspace_bytes", c_int64)]


class MnistStepArgs(C.Structure):
    _fields_ = [
        ("batch", c_int), ("n_latents", c_int), ("dtype", c_int),
        ("n_terms", c_int),
        ("term_type", c_int * 3),
        ("lambda_image", c_float * 3), ("lambda_text", c_float * 3), ("kl_weight", c_float * 3),
        ("poe_mode", c_int), ("prior_expert", c_int),
        ("poe_eps", c_float),
        ("image", c_void_p), ("text", c_void_p), ("eps", c_void_p),
        ("seed", c_uint64),
        ("params", c_void_p), ("params_bf16", c_void_p), ("buffers", c_void_p),
        ("num_batches_tracked", c_void_p), ("grads", c_void_p),
        ("do_backward", c_int), ("zero_grad", c_int), ("do_adam", c_int),
        ("adam_m", c_void_p), ("adam_v", c_void_p), ("adam_step", c_void_p),
        ("lr", c_float), ("beta1", c_float), ("beta2", c_float), ("adam_eps", c_float), ("grad_scale", c_float),
        ("workspace", c_void_p), ("workspace_bytes", c_int64),
        ("out_losses", c_void_p), ("out_recon_image", c_void_p), ("out_recon_text", c_void_p),
        ("out_mu", c_void_p), ("out_logvar", c_void_p),
        ("eval_mode", c_int), ("phase", c_int), ("z_in", c_void_p),
        ("d_recon_image", c_void_p), ("d_recon_text", c_void_p), ("d_mu", c_void_p), ("d_logvar", c_void_p),
        ("noise_step", c_void_p), ("advance_adam_step", c_int),
    ]


class ElboLossArgs(C.Structure):
    _fields_ = [
        ("image_dtype", c_int), ("batch", c_int64), ("n_pixels", c_int), ("n_classes", c_int), ("n_latents", c_int),
        ("recon_image", c_void_p), ("image", c_void_p), ("recon_text", c_void_p), ("text", c_void_p),
        ("mu", c_void_p), ("logvar", c_void_p),
        ("lambda_image", c_float), ("lambda_text", c_float), ("kl_weight", c_float),
    ]


DT_F32, DT_BF16, DT_F32X3 = 0, 1, 2
POE_REF, POE_PRECISION = 0, 1
TERM_JOINT, TERM_IMAGE, TERM_TEXT = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_SWISH = 0, 1, 2


# ---------------------------------------------------------------- operator-level ABI (conv models)
class ConvGeometry(C.Structure):
    _fields_ = [("batch", c_int), ("height", c_int), ("width", c_int), ("channels", c_int),
                ("kernel", c_int), ("stride", c_int), ("pad", c_int),
                ("stride_n", c_int64), ("stride_h", c_int64), ("stride_w", c_int64), ("stride_c", c_int64)]


class BnActArgs(C.Structure):
    _fields_ = [("dtype", c_int), ("rows", c_int64), ("channels", c_int), ("rows_per_group", c_int64),
                ("act", c_int), ("training", c_int),
                ("x", c_void_p), ("y", c_void_p), ("gamma", c_void_p), ("beta", c_void_p),
                ("sum", c_void_p), ("sumsq", c_void_p), ("stats_ready", c_int),
                ("save_mean", c_void_p), ("save_rstd", c_void_p),
                ("running_mean", c_void_p), ("running_var", c_void_p),
                ("updates_per_group", c_int), ("momentum", c_float), ("eps", c_float),
                ("dy", c_void_p), ("dx", c_void_p), ("s0", c_void_p), ("s1", c_void_p),
                ("dgamma", c_void_p), ("dbeta", c_void_p)]


class SigmoidBceArgs(C.Structure):
    _fields_ = [("rows", c_int64), ("cols", c_int), ("rows_per_group", c_int64),
                ("logit_dtype", c_int), ("logits", c_void_p), ("ld_logits", c_int64),
                ("target_dtype", c_int), ("target", c_void_p), ("ld_target", c_int64), ("target_rows", c_int64),
                ("grad_scale", c_float * 3), ("loss", c_void_p),
                ("prob_dtype", c_int), ("probs", c_void_p), ("ld_probs", c_int64),
                ("grad_dtype", c_int), ("dlogits", c_void_p), ("ld_dlogits", c_int64),
                ("dprobs", c_void_p), ("ld_dprobs", c_int64)]


class LatentArgs(C.Structure):
    _fields_ = [("batch", c_int64), ("n_latents", c_int), ("n_terms", c_int), ("term_type", c_int * 3),
                ("poe_mode", c_int), ("prior_expert", c_int), ("poe_eps", c_float),
                ("expert_a", c_void_p), ("ld_a", c_int64), ("expert_a_row0", c_int64 * 3),
                ("expert_b", c_void_p), ("ld_b", c_int64),
                ("eps", c_void_p), ("seed", c_uint64), ("step_counter", c_void_p), ("training", c_int),
                ("kl_weight", c_float * 3),
                ("z_dtype", c_int), ("z", c_void_p), ("ld_z", c_int64),
                ("mu", c_void_p), ("logvar", c_void_p), ("kl", c_void_p),
                ("dz_dtype", c_int), ("dz", c_void_p), ("ld_dz", c_int64),
                ("d_mu", c_void_p), ("d_logvar", c_void_p),
                ("d_dtype", c_int), ("d_expert_a", c_void_p), ("ld_da", c_int64),
                ("d_expert_b", c_void_p), ("ld_db", c_int64), ("row_weight", c_void_p)]


class GruCellArgs(C.Structure):
    _fields_ = [("rows", c_int64), ("hidden", c_int),
                ("gi", c_void_p), ("ld_gi", c_int64), ("gh", c_void_p), ("ld_gh", c_int64),
                ("h_dtype", c_int), ("h_prev", c_void_p), ("ld_h_prev", c_int64),
                ("addend", c_void_p), ("ld_addend", c_int64),
                ("h_out", c_void_p), ("ld_h_out", c_int64), ("h_out2", c_void_p), ("ld_h_out2", c_int64),
                ("saved", c_void_p),
                ("dh_a_dtype", c_int), ("dh_a", c_void_p), ("ld_dh_a", c_int64),
                ("dh_b_dtype", c_int), ("dh_b", c_void_p), ("ld_dh_b", c_int64),
                ("dg_dtype", c_int), ("dgi", c_void_p), ("dgh", c_void_p), ("ld_dg", c_int64),
                ("dh_prev", c_void_p), ("ld_dh_prev", c_int64)]


class LogSoftmaxNllArgs(C.Structure):
    _fields_ = [("rows", c_int64), ("classes", c_int), ("rows_per_group", c_int64),
                ("logits", c_void_p), ("ld_logits", c_int64),
                ("target", c_void_p), ("target_stride", c_int64), ("target_rows", c_int64),
                ("grad_scale", c_float * 3), ("loss", c_void_p),
                ("logp", c_void_p), ("ld_logp", c_int64), ("argmax", c_void_p),
                ("grad_dtype", c_int), ("dlogits", c_void_p), ("ld_dlogits", c_int64), ("row_weight", c_void_p)]


class ConvTClass(C.Structure):
    """mvae_convt_class: one output-parity class of a transposed convolution (include/mvae_b200.h)."""
    _fields_ = [
        ("batch", c_int), ("in_h", c_int), ("in_w", c_int), ("channels", c_int),
        ("out_h", c_int), ("out_w", c_int), ("out_channels", c_int),
        ("kernel", c_int), ("stride", c_int),
        ("a", c_int), ("b", c_int),
        ("count_h", c_int), ("count_w", c_int),
        ("taps_h", c_int), ("taps_w", c_int), ("pad_h", c_int), ("pad_w", c_int),
        ("kh", c_int * 8), ("kw", c_int * 8),
    ]


class DpReduceAdamArgs(C.Structure):
    """mvae_dp_reduce_adam_args (include/mvae_b200.h)."""
    _fields_ = [("world", c_int), ("rank", c_int), ("grads", c_void_p * 8), ("flags", c_void_p * 8),
                ("params", c_void_p), ("adam_m", c_void_p), ("adam_v", c_void_p), ("params_bf16", c_void_p),
                ("lo", c_int64), ("hi", c_int64), ("lr", c_float), ("beta1", c_float), ("beta2", c_float),
                ("eps", c_float), ("grad_scale", c_float), ("adam_step", c_void_p), ("blocks", c_int)]


def _conv_argtypes(lib) -> None:
    P = C.POINTER
    lib.mvae_convt_axis_classes.argtypes = [c_int] * 4 + [P(c_int)] * 4
    lib.mvae_convt_gemm.argtypes = [c_int] * 9 + [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p]
    lib.mvae_convt_class_gemm.argtypes = [P(ConvTClass), c_int, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p]
    lib.mvae_conv_out_size.argtypes = [c_int] * 4
    lib.mvae_im2col.argtypes = [P(ConvGeometry), c_int, c_void_p, c_int, c_void_p, c_int64, c_void_p]
    lib.mvae_col2im.argtypes = [P(ConvGeometry), c_int, c_void_p, c_int64, c_int, c_void_p, c_void_p]
    lib.mvae_col_stats.argtypes = [c_int, c_void_p, c_int64, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p]
    lib.mvae_bn_act_forward.argtypes = [P(BnActArgs), c_void_p]
    lib.mvae_bn_act_backward.argtypes = [P(BnActArgs), c_void_p]
    lib.mvae_act_forward.argtypes = [c_int, c_int, c_void_p, c_void_p, c_int64, c_int, c_int, c_float, c_uint64,
                                     c_void_p, c_void_p]
    lib.mvae_act_backward.argtypes = [c_int, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_float,
                                      c_uint64, c_void_p, c_void_p, c_void_p]
    lib.mvae_sigmoid_bce.argtypes = [P(SigmoidBceArgs), c_void_p]
    lib.mvae_latent_forward.argtypes = [P(LatentArgs), c_void_p]
    lib.mvae_latent_backward.argtypes = [P(LatentArgs), c_void_p]
    lib.mvae_cast_pad_2d.argtypes = [c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_int64, c_void_p]
    lib.mvae_step_begin.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p]
    lib.mvae_gemm.argtypes = [P(GemmArgs), c_void_p]
    lib.mvae_conv_gemm.argtypes = [P(GemmArgs), P(ConvGeometry), c_int, c_void_p]
    lib.mvae_embed_forward.argtypes = [c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int64, c_int64,
                                       c_void_p]
    lib.mvae_embed_backward.argtypes = [c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int64, c_int64,
                                        c_void_p, c_void_p]
    lib.mvae_gru_cell_forward.argtypes = [P(GruCellArgs), c_void_p]
    lib.mvae_gru_cell_backward.argtypes = [P(GruCellArgs), c_void_p]
    lib.mvae_logsoftmax_nll.argtypes = [P(LogSoftmaxNllArgs), c_void_p]
    lib.mvae_logsoftmax_backward.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_int64, c_void_p]
    lib.mvae_copy_2d.argtypes = [c_int, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_void_p,
                                 c_int64, c_void_p]
