"""Thin Python wrappers over the operator-level C ABI (include/mvae_b200.h): one function per entry point, torch
tensors in, nothing computed on the host.  Used by the conv-model hosts (celeba.py).  No fallback: every function
ends in a call into libmvae_b200.so and raises MvaeError on failure.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib

DT = {torch.float32: _lib.DT_F32, torch.bfloat16: _lib.DT_BF16}


def stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def gemm(A, B, Cout, M, N, K, lda, ldb, ldc, a_major=0, b_major=0, bias=None, accumulate=False, col_sum=None,
         col_sumsq=None, rows_per_group=0, patch=None, split_k=0, block_n=0, stages=0, act=0, act_out=None, act_pre=None,
         ld_act_pre=0, bce=None):
    """Cout[M,N] (+)= A[M,K] * B[N,K]^T (+ bias).  A and B share one storage dtype (bf16 -> kind::f16, fp32 -> tf32).
    `patch=(geometry, operand)`: implicit GEMM - operand 1: A is the channels-last image whose patch matrix is the real A
    (lda ignored); operand 2: likewise for B (b_major must be 1).  See mvae_conv_gemm.
    `act` (ACT_SWISH) fuses the activation into the epilogue: with `act_out` the forward (Cout = pre-activation or None,
    act_out = swish(pre)); with `act_pre` the backward (Cout = (A B^T) * swish'(act_pre), col_sum += its column sums)."""
    if A.dtype != B.dtype:
        raise TypeError("gemm operands differ in dtype: %s vs %s" % (A.dtype, B.dtype))
    a = _lib.GemmArgs()
    a.dtype = DT[A.dtype]
    a.M, a.N, a.K = int(M), int(N), int(K)
    a.A, a.lda, a.a_major = A.data_ptr(), int(lda), int(a_major)
    a.B, a.ldb, a.b_major = B.data_ptr(), int(ldb), int(b_major)
    cref = Cout if Cout is not None else act_out
    a.C, a.ldc, a.c_dtype = _p(Cout), int(ldc), DT[cref.dtype]
    a.act, a.act_out, a.act_pre, a.ld_act_pre = int(act), _p(act_out), _p(act_pre), int(ld_act_pre)
    if bce is not None:   # dict(target, ld_target, target_rows, scale=[per group], loss, probs=None): see mvae_gemm_args
        a.bce_target, a.ld_bce_target, a.bce_target_rows = bce["target"].data_ptr(), int(bce["ld_target"]), int(bce["target_rows"])
        for i, v in enumerate(bce["scale"]):
            a.bce_scale[i] = float(v)
        a.bce_loss, a.bce_probs, a.bce_row_weight = _p(bce.get("loss")), _p(bce.get("probs")), _p(bce.get("row_weight"))
    a.bias = _p(bias)
    a.accumulate = 1 if accumulate else 0
    a.col_sum, a.col_sumsq = _p(col_sum), _p(col_sumsq)
    a.rows_per_group = int(rows_per_group)
    a.split_k, a.block_n, a.stages = int(split_k), int(block_n), int(stages)
    if patch is not None:
        geom, operand = patch
        _lib.check(_lib.load().mvae_conv_gemm(C.byref(a), C.byref(geom), int(operand), stream()),
                   "mvae_conv_gemm %dx%dx%d" % (M, N, K))
        return
    _lib.check(_lib.load().mvae_gemm(C.byref(a), stream()), "mvae_gemm %dx%dx%d" % (M, N, K))


def geometry(batch, height, width, channels, kernel, stride, pad, strides=None) -> _lib.ConvGeometry:
    g = _lib.ConvGeometry()
    g.batch, g.height, g.width, g.channels = int(batch), int(height), int(width), int(channels)
    g.kernel, g.stride, g.pad = int(kernel), int(stride), int(pad)
    if strides is None:  # dense NHWC
        strides = (height * width * channels, width * channels, channels, 1)
    g.stride_n, g.stride_h, g.stride_w, g.stride_c = (int(s) for s in strides)
    return g


def nchw_strides(channels, height, width):
    return (channels * height * width, width, 1, height * width)


def out_size(n, kernel, stride, pad) -> int:
    return (n + 2 * pad - kernel) // stride + 1


def transposed_conv_classes(kernel: int, stride: int, pad: int, size_in: int):
    """Output-parity decomposition of ConvTranspose2d (celeba/model.py:142-152, multimnist/model.py:198-210) along one axis.

    out[o] = sum over (i, kh) with o = i*stride - pad + kh of x[i] * w[kh].  Outputs with o % stride == a only ever see the
    taps kh = r + stride*t (r = (a + pad) % stride), so each class is a STRIDE-1 convolution of x with a short kernel:
        out[stride*u + a] = sum_{t' < taps} x[u - pad_lo + t'] * w[kh[t']],   u in [0, count)
    Returns (size_out, [dict(a, count, pad_lo, kh=[...]) per class]).  This is the index map of the implicit (col2im-free)
    transposed convolution: a gather GEMM per class (patch matrix = A, K = taps_h*taps_w*Cin) writing interleaved rows."""
    size_out = (size_in - 1) * stride - 2 * pad + kernel
    classes = []
    for a in range(stride):
        r = (a + pad) % stride
        taps = (kernel - r + stride - 1) // stride if kernel > r else 0
        q = (a + pad) // stride
        count = (size_out - a + stride - 1) // stride if size_out > a else 0
        classes.append({"a": a, "count": count, "pad_lo": taps - 1 - q, "kh": [r + stride * (taps - 1 - t) for t in range(taps)]})
    return size_out, classes


def transposed_conv_implicit(x, weight, out, batch, size_in, channels, out_channels, kernel, stride, pad, ldc=None,
                             merged=False):
    """mvae_convt_class_gemm per class: ConvTranspose2d forward / Conv2d input gradient without the
    patch matrix - one gather GEMM per output-parity class writing the interleaved rows of `out` [batch, H_out, H_out, ldc].
    x: [batch, size_in, size_in, channels] bf16 channels-last; weight: [channels, kernel, kernel, out_channels] bf16.
    `merged` (draft): all classes in one launch (mvae_convt_gemm) when they have equal shape, else the per-class loop."""
    size_out, classes = transposed_conv_classes(kernel, stride, pad, size_in)
    lib = _lib.load()
    if merged:
        rc = lib.mvae_convt_gemm(DT[x.dtype], int(batch), int(size_in), int(size_in), int(channels), int(out_channels),
                                 int(kernel), int(stride), int(pad), x.data_ptr(), weight.data_ptr(), int(out_channels),
                                 out.data_ptr(), int(ldc or out_channels), DT[out.dtype], stream())
        if rc == 0:
            return size_out
        if rc != 4:   # 4 = classes of unequal shape: fall through to the per-class launches
            _lib.check(rc, "mvae_convt_gemm")
    for ca in classes:
        for cb in classes:
            if ca["count"] == 0 or cb["count"] == 0 or not ca["kh"] or not cb["kh"]:
                continue   # (a class without taps would be all zeros: not produced by the geometries of the reference)
            c = _lib.ConvTClass()
            c.batch, c.in_h, c.in_w, c.channels = int(batch), int(size_in), int(size_in), int(channels)
            c.out_h, c.out_w, c.out_channels = size_out, size_out, int(out_channels)
            c.kernel, c.stride, c.a, c.b = int(kernel), int(stride), ca["a"], cb["a"]
            c.count_h, c.count_w = ca["count"], cb["count"]
            c.taps_h, c.taps_w, c.pad_h, c.pad_w = len(ca["kh"]), len(cb["kh"]), ca["pad_lo"], cb["pad_lo"]
            for t, v in enumerate(ca["kh"]):
                c.kh[t] = v
            for t, v in enumerate(cb["kh"]):
                c.kw[t] = v
            _lib.check(lib.mvae_convt_class_gemm(C.byref(c), DT[x.dtype], x.data_ptr(), weight.data_ptr(), int(out_channels),
                                                 out.data_ptr(), int(ldc or out_channels), DT[out.dtype], stream()),
                       "mvae_convt_class_gemm")
    return size_out


def im2col(g, image, col, ldcol):
    _lib.check(_lib.load().mvae_im2col(C.byref(g), DT[image.dtype], image.data_ptr(), DT[col.dtype], col.data_ptr(),
                                       int(ldcol), stream()), "mvae_im2col")


def col2im(g, col, ldcol, image):
    _lib.check(_lib.load().mvae_col2im(C.byref(g), DT[col.dtype], col.data_ptr(), int(ldcol), DT[image.dtype],
                                       image.data_ptr(), stream()), "mvae_col2im")


def col_stats(x, rows, channels, sum_, sumsq=None, valid_channels=0, rows_per_group=0):
    _lib.check(_lib.load().mvae_col_stats(DT[x.dtype], x.data_ptr(), int(rows), int(channels), int(valid_channels),
                                          int(rows_per_group), _p(sum_), _p(sumsq), stream()), "mvae_col_stats")


def bn_args(x, rows, channels, rows_per_group, act, training, gamma, beta, sum_, sumsq, save_mean, save_rstd,
            running_mean=None, running_var=None, updates=1, stats_ready=False) -> _lib.BnActArgs:
    a = _lib.BnActArgs()
    a.dtype = DT[x.dtype]
    a.rows, a.channels, a.rows_per_group = int(rows), int(channels), int(rows_per_group)
    a.act, a.training = int(act), 1 if training else 0
    a.x = x.data_ptr()
    a.gamma, a.beta = gamma.data_ptr(), beta.data_ptr()
    a.sum, a.sumsq, a.stats_ready = _p(sum_), _p(sumsq), 1 if stats_ready else 0
    a.save_mean, a.save_rstd = _p(save_mean), _p(save_rstd)
    a.running_mean, a.running_var = _p(running_mean), _p(running_var)
    a.updates_per_group, a.momentum, a.eps = int(updates), 0.1, 1e-5
    return a


def bn_act_forward(a: _lib.BnActArgs, y):
    a.y = y.data_ptr()
    _lib.check(_lib.load().mvae_bn_act_forward(C.byref(a), stream()), "mvae_bn_act_forward")


def bn_act_backward(a: _lib.BnActArgs, dy, dx, s0, s1, dgamma=None, dbeta=None):
    a.dy, a.dx, a.s0, a.s1 = dy.data_ptr(), dx.data_ptr(), s0.data_ptr(), s1.data_ptr()
    a.dgamma, a.dbeta = _p(dgamma), _p(dbeta)
    _lib.check(_lib.load().mvae_bn_act_backward(C.byref(a), stream()), "mvae_bn_act_backward")


def act_forward(act, x, y, rows, channels, repeat=1, dropout_p=0.0, seed=0, step_counter=None):
    _lib.check(_lib.load().mvae_act_forward(DT[x.dtype], int(act), x.data_ptr(), y.data_ptr(), int(rows), int(channels),
                                            int(repeat), float(dropout_p), int(seed), _p(step_counter), stream()),
               "mvae_act_forward")


def act_backward(act, x, dy, dx, rows, channels, repeat=1, dropout_p=0.0, seed=0, step_counter=None, dbias=None):
    _lib.check(_lib.load().mvae_act_backward(DT[x.dtype], int(act), x.data_ptr(), dy.data_ptr(), dx.data_ptr(), int(rows),
                                             int(channels), int(repeat), float(dropout_p), int(seed), _p(step_counter),
                                             _p(dbias), stream()), "mvae_act_backward")


def sigmoid_bce(logits, ld_logits, rows, cols, rows_per_group=0, target=None, ld_target=0, target_rows=0,
                grad_scale=(0.0, 0.0, 0.0), loss=None, probs=None, ld_probs=0, dlogits=None, ld_dlogits=0, dprobs=None,
                ld_dprobs=0):
    a = _lib.SigmoidBceArgs()
    a.rows, a.cols, a.rows_per_group = int(rows), int(cols), int(rows_per_group)
    a.logit_dtype, a.logits, a.ld_logits = DT[logits.dtype], logits.data_ptr(), int(ld_logits)
    if target is not None:
        a.target_dtype, a.target, a.ld_target, a.target_rows = DT[target.dtype], target.data_ptr(), int(ld_target), int(target_rows)
    for i in range(3):
        a.grad_scale[i] = float(grad_scale[i]) if i < len(grad_scale) else 0.0
    a.loss = _p(loss)
    if probs is not None:
        a.prob_dtype, a.probs, a.ld_probs = DT[probs.dtype], probs.data_ptr(), int(ld_probs)
    if dlogits is not None:
        a.grad_dtype, a.dlogits, a.ld_dlogits = DT[dlogits.dtype], dlogits.data_ptr(), int(ld_dlogits)
    if dprobs is not None:
        a.dprobs, a.ld_dprobs = dprobs.data_ptr(), int(ld_dprobs)
    _lib.check(_lib.load().mvae_sigmoid_bce(C.byref(a), stream()), "mvae_sigmoid_bce")


def mask_weights(has_image, has_text, batch, term_types, weight, counts=None):
    """Per-(term, row) loss weights batch / count from uint8 presence masks, on the device (mvae_mask_weights)."""
    tt = (C.c_int * len(term_types))(*[int(t) for t in term_types])
    _lib.check(_lib.load().mvae_mask_weights(C.c_void_p(has_image.data_ptr()), C.c_void_p(has_text.data_ptr()), C.c_int64(int(batch)),
                                             len(term_types), tt, C.c_void_p(weight.data_ptr()), C.c_void_p(_p(counts) or 0), stream()),
               "mvae_mask_weights")


def cast_pad_2d(src, rows, cols, ld_src, dst, ld_dst):
    _lib.check(_lib.load().mvae_cast_pad_2d(src.data_ptr(), int(rows), int(cols), int(ld_src), DT[dst.dtype], dst.data_ptr(),
                                            int(ld_dst), stream()), "mvae_cast_pad_2d")


def step_begin(step_counter, zero_buf=None, counters=None, increments=None):
    n = 0 if counters is None else counters.numel()
    _lib.check(_lib.load().mvae_step_begin(_p(step_counter), _p(zero_buf), 0 if zero_buf is None else zero_buf.numel(),
                                           _p(counters), _p(increments), n, stream()), "mvae_step_begin")


def adam_step(params, grads, m, v, params_bf16, count, lr, beta1, beta2, eps, step_counter, grad_scale=1.0, zero_grad=True):
    _lib.check(_lib.load().mvae_adam_step(params.data_ptr(), grads.data_ptr(), m.data_ptr(), v.data_ptr(), _p(params_bf16),
                                          int(count), lr, beta1, beta2, eps, step_counter.data_ptr(), grad_scale,
                                          1 if zero_grad else 0, stream()), "mvae_adam_step")


def cast_f32_to_bf16(src, dst, count):
    _lib.check(_lib.load().mvae_cast_f32_to_bf16(C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()), C.c_int64(int(count)),
                                                 stream()), "mvae_cast_f32_to_bf16")


# ---------------------------------------------------------------- MultiMNIST text path
def view_ptr(t: torch.Tensor, elem_offset: int = 0) -> int:
    """Device address of element `elem_offset` of a flat tensor."""
    return t.data_ptr() + int(elem_offset) * t.element_size()


def embed_forward(indices, index_offset, index_stride, table, vocab, width, act, out, col_off, ld_out, rows):
    _lib.check(_lib.load().mvae_embed_forward(view_ptr(indices, index_offset), int(index_stride), table.data_ptr(), int(vocab),
                                              int(width), int(act), DT[out.dtype], view_ptr(out, col_off), int(ld_out), int(rows),
                                              stream()), "mvae_embed_forward")


def embed_backward(indices, index_offset, index_stride, table, vocab, width, act, dout, col_off, ld_dout, rows, dtable):
    _lib.check(_lib.load().mvae_embed_backward(view_ptr(indices, index_offset), int(index_stride), table.data_ptr(), int(vocab),
                                               int(width), int(act), DT[dout.dtype], view_ptr(dout, col_off), int(ld_dout),
                                               int(rows), dtable.data_ptr(), stream()), "mvae_embed_backward")


def gru_cell_forward(rows, hidden, gi, gh, h_prev, ld_h_prev, h_out, ld_h_out, saved, h_out2=None, ld_h_out2=0, addend=None,
                     ld_addend=0) -> _lib.GruCellArgs:
    a = _lib.GruCellArgs()
    a.rows, a.hidden = int(rows), int(hidden)
    a.gi, a.ld_gi, a.gh, a.ld_gh = gi.data_ptr(), 3 * hidden, gh.data_ptr(), 3 * hidden
    a.h_dtype = DT[h_out.dtype]
    a.h_prev, a.ld_h_prev = _p(h_prev), int(ld_h_prev)
    a.addend, a.ld_addend = _p(addend), int(ld_addend)
    a.h_out, a.ld_h_out = h_out.data_ptr(), int(ld_h_out)
    a.h_out2, a.ld_h_out2 = _p(h_out2), int(ld_h_out2)
    a.saved = saved.data_ptr()
    _lib.check(_lib.load().mvae_gru_cell_forward(C.byref(a), stream()), "mvae_gru_cell_forward")
    return a


def gru_cell_backward(a: _lib.GruCellArgs, dh_a, ld_dh_a, dh_b, ld_dh_b, dgi, dgh, ld_dg, dh_prev, ld_dh_prev):
    """`a` is the struct returned by gru_cell_forward for the same cell (gi / gh / h_prev / saved are reused)."""
    a.dh_a_dtype, a.dh_a, a.ld_dh_a = (DT[dh_a.dtype], dh_a.data_ptr(), int(ld_dh_a)) if dh_a is not None else (0, None, 0)
    a.dh_b_dtype, a.dh_b, a.ld_dh_b = (DT[dh_b.dtype], dh_b.data_ptr(), int(ld_dh_b)) if dh_b is not None else (0, None, 0)
    a.dg_dtype, a.dgi, a.dgh, a.ld_dg = DT[dgi.dtype], dgi.data_ptr(), dgh.data_ptr(), int(ld_dg)
    a.dh_prev, a.ld_dh_prev = _p(dh_prev), int(ld_dh_prev)
    _lib.check(_lib.load().mvae_gru_cell_backward(C.byref(a), stream()), "mvae_gru_cell_backward")


def logsoftmax_nll(logits, ld_logits, rows, classes, rows_per_group=0, target=None, target_offset=0, target_stride=1,
                   target_rows=0, grad_scale=(0.0, 0.0, 0.0), loss=None, logp=None, logp_offset=0, ld_logp=0, argmax=None,
                   dlogits=None, ld_dlogits=0, row_weight=None):
    a = _lib.LogSoftmaxNllArgs()
    a.rows, a.classes, a.rows_per_group = int(rows), int(classes), int(rows_per_group)
    a.logits, a.ld_logits = logits.data_ptr(), int(ld_logits)
    if target is not None:
        a.target, a.target_stride, a.target_rows = view_ptr(target, target_offset), int(target_stride), int(target_rows)
    for i in range(3):
        a.grad_scale[i] = float(grad_scale[i]) if i < len(grad_scale) else 0.0
    a.loss = _p(loss)
    if logp is not None:
        a.logp, a.ld_logp = view_ptr(logp, logp_offset), int(ld_logp)
    a.argmax = _p(argmax)
    if dlogits is not None:
        a.grad_dtype, a.dlogits, a.ld_dlogits = DT[dlogits.dtype], dlogits.data_ptr(), int(ld_dlogits)
    a.row_weight = _p(row_weight)
    _lib.check(_lib.load().mvae_logsoftmax_nll(C.byref(a), stream()), "mvae_logsoftmax_nll")


def logsoftmax_backward(logp, logp_off, ld_logp, dlogp, dlogp_off, ld_dlogp, rows, classes, dlogits, ld_dlogits):
    _lib.check(_lib.load().mvae_logsoftmax_backward(view_ptr(logp, logp_off), int(ld_logp), view_ptr(dlogp, dlogp_off), int(ld_dlogp),
                                                    int(rows), int(classes), DT[dlogits.dtype], dlogits.data_ptr(), int(ld_dlogits),
                                                    stream()), "mvae_logsoftmax_backward")


def copy_2d(src, src_off, ld_src, dst, dst_off, ld_dst, rows, cols, accumulate=False, src2=None, src2_off=0, ld_src2=0):
    _lib.check(_lib.load().mvae_copy_2d(DT[src.dtype], view_ptr(src, src_off), int(ld_src), DT[dst.dtype], view_ptr(dst, dst_off),
                                        int(ld_dst), int(rows), int(cols), 1 if accumulate else 0,
                                        DT[src2.dtype] if src2 is not None else 0,
                                        view_ptr(src2, src2_off) if src2 is not None else None, int(ld_src2), stream()),
               "mvae_copy_2d")
