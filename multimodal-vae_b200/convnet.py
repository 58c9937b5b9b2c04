"""Shared host code of the convolutional MVAEs (CelebA celeba/model.py, MultiMNIST multimnist/model.py).

ConvMVAEBase owns the flat parameter / gradient buffers (reference state_dict keys, internal GEMM-friendly layouts),
the convolutional image encoder (`features`) and decoder (`hallucinate`) stacks composed from the operator-level C ABI
(mvae_im2col / mvae_col2im / mvae_gemm / mvae_bn_act_* / mvae_act_*), the Linear helpers, and the latent path
(mvae_latent_*).  ConvMVAETrainer owns the step: zero accumulators, forward of all ELBO terms, backward, the
data-parallel gradient all-reduce (decoder bucket overlapped with the encoder backward) and fused Adam, optionally as
one CUDA graph.  Subclasses provide the model spec and the second modality's networks.

Every Conv2d / ConvTranspose2d / Linear is a tcgen05 GEMM over NHWC activation matrices [batch*H*W, C].  The image
encoder runs ONCE per step (the reference runs it twice on identical inputs; only Dropout differs, so everything up
to the first Dropout is shared and the masks are applied to replicated rows); the decoders run once on the stacked
[terms*B] latents with per-term BatchNorm statistics.  There is no PyTorch fallback.
"""
from __future__ import annotations

import os

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib, _ops

_DTYPES = {"tf32": torch.float32, "fp32": torch.float32, "bf16": torch.bfloat16}
SWISH = _lib.ACT_SWISH


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class Layout:
    """How one reference parameter tensor is held inside the flat buffer.
    kind: conv [Co,Ci,kh,kw]->[Co,kh,kw,Ci] | convT [Ci,Co,kh,kw]->[Ci,kh,kw,Co] | fc_in (columns (c,hw)->(hw,c)) |
          fc_out (rows (c,hw)->(hw,c)) | fc_out_bias | plain"""

    def __init__(self, key, ref_shape, kind, flat_c=256, flat_hw=1):
        self.key, self.ref_shape, self.kind = key, tuple(ref_shape), kind
        self.c, self.hw = flat_c, flat_hw
        self.numel = 1
        for s in ref_shape:
            self.numel *= s
        self.offset = -1

    def to_internal(self, t: torch.Tensor) -> torch.Tensor:
        k, c, hw = self.kind, self.c, self.hw
        if k in ("conv", "convT"):
            return t.permute(0, 2, 3, 1).contiguous()
        if k == "fc_in":
            o = t.shape[0]
            return t.reshape(o, c, hw).permute(0, 2, 1).contiguous().reshape(o, c * hw)
        if k == "fc_out":
            n = t.shape[1]
            return t.reshape(c, hw, n).permute(1, 0, 2).contiguous().reshape(c * hw, n)
        if k == "fc_out_bias":
            return t.reshape(c, hw).t().contiguous().reshape(c * hw)
        return t.contiguous()

    def to_reference(self, t: torch.Tensor) -> torch.Tensor:
        k, c, hw = self.kind, self.c, self.hw
        if k in ("conv", "convT"):
            a, b, kh, kw = self.ref_shape
            return t.reshape(a, kh, kw, b).permute(0, 3, 1, 2).contiguous()
        if k == "fc_in":
            o = self.ref_shape[0]
            return t.reshape(o, hw, c).permute(0, 2, 1).contiguous().reshape(o, c * hw)
        if k == "fc_out":
            n = self.ref_shape[1]
            return t.reshape(hw, c, n).permute(1, 0, 2).contiguous().reshape(c * hw, n)
        if k == "fc_out_bias":
            return t.reshape(hw, c).t().contiguous().reshape(c * hw)
        return t.reshape(self.ref_shape).clone()


class ConvMVAEBase:
    """Parameter store + conv stacks.  Subclass attributes:
        IMG_C, IMG_H           image channels / side
        ENC_CONVS              [(key, Cin, Cout, k, stride, pad, H_in, bn_key | None)]
        DEC_CONVS              [(key, Cin, Cout, k, stride, pad, H_out, bn_key | None)]   (last one produces the logits)
        FLAT_C, FLAT_HW        channels / pixels of the bottleneck feature map
        BN_LAYERS              {bn_key: channels} in a fixed order (num_batches_tracked index)
    and reference_keys(n_latents) -> [(key, reference shape, kind)] in state_dict order."""

    IMG_C = 3
    IMG_H = 64
    ENC_CONVS: Tuple = ()
    DEC_CONVS: Tuple = ()
    FLAT_C = 256
    FLAT_HW = 25
    BN_LAYERS: Dict[str, int] = {}

    def __init__(self, n_latents: int, precision: str, dropout_p: float, device, seed: int):
        if precision not in _DTYPES:
            raise ValueError("precision must be one of %s" % sorted(_DTYPES))
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if dev.type != "cuda":
            raise RuntimeError("mvae_b200 has no CPU path: a CUDA (sm_100) device is required")
        _lib.check(_lib.load().mvae_device_check(dev.index or 0), "mvae_device_check")
        self.device = dev
        self.n_latents = int(n_latents)
        self.precision = precision
        self.act_dtype = _DTYPES[precision]
        self.vec = 8 if self.act_dtype == torch.bfloat16 else 4
        self.dropout_p = float(dropout_p)
        self.noise_seed = int(seed)
        self.training = True
        self.poe_mode, self.prior_expert, self.poe_eps = _lib.POE_REF, 0, 1e-8
        self.flat_feat = self.FLAT_C * self.FLAT_HW
        self.n_pixels = self.IMG_C * self.IMG_H * self.IMG_H
        # ---- flat parameter buffer, [encoders | decoders] (the two all-reduce buckets of data-parallel training)
        self.layouts: Dict[str, Layout] = {}
        self.state_keys = self.reference_keys(self.n_latents)
        params = [Layout(k, s, kind, self.FLAT_C, self.FLAT_HW) for k, s, kind in self.state_keys
                  if kind not in ("rm", "rv", "nbt")]
        enc = [l for l in params if "encoder" in l.key]
        dec = [l for l in params if "encoder" not in l.key]
        off = 0
        for l in enc + dec:
            l.offset = off
            off += round_up(l.numel, 64)
            self.layouts[l.key] = l
            if l is enc[-1]:
                self.encoder_param_floats = off
        self.param_floats = off
        self.flat_params = torch.zeros(off, device=dev, dtype=torch.float32)
        self.flat_grads = torch.zeros(off, device=dev, dtype=torch.float32)
        self.flat_params_bf16 = torch.zeros(off, device=dev, dtype=torch.bfloat16) if self.act_dtype == torch.bfloat16 else None
        # the autograd / torch.optim view of the parameters: ONE leaf sharing the flat buffer's storage
        self.param = torch.nn.Parameter(self.flat_params)
        self.param.grad = self.flat_grads
        # ---- BatchNorm buffers
        self.bn_names = list(self.BN_LAYERS)
        self.bn_index = {p: i for i, p in enumerate(self.bn_names)}
        self.bn_off: Dict[str, int] = {}
        boff = 0
        for p in self.bn_names:
            self.bn_off[p] = boff
            boff += 2 * self.BN_LAYERS[p]
        self.flat_buffers = torch.zeros(max(boff, 4), device=dev, dtype=torch.float32)
        self.flat_nbt = torch.zeros(max(len(self.bn_names), 1), device=dev, dtype=torch.int64)
        # [0] noise / dropout (Philox) counter, ticked by every forward; [1] Adam's bias-correction step
        self._counters = torch.zeros(2, device=dev, dtype=torch.int32)
        self._step_counter = self._counters[0:1]
        self._adam_counter = self._counters[1:2]
        self._ws: Dict[Tuple[int, int], object] = {}
        self._pad: Dict[str, Tuple[torch.Tensor, int]] = {}
        self._fresh = set()
        # weight-gradient GEMMs (small split-K grids) run on a side stream under the bandwidth-bound dgrad / col2im / BatchNorm
        # kernels of the main stream; fork / join are stream-ordered events (valid under CUDA-graph capture)
        self.side_stream = torch.cuda.Stream(device=dev)
        self.use_side_stream = True
        # implicit GEMM: patch matrices of layers with >= 8 bf16 channels are never materialised.  The im2col side
        # (Conv2d forward, ConvTranspose2d input gradient, both weight gradients) gathers its operand inside mvae_conv_gemm;
        # the col2im side (ConvTranspose2d forward, Conv2d input gradient) runs as stride^2 output-parity classes in ONE
        # launch of mvae_convt_gemm (classes of unequal shape fall back to one launch per class).  tf32 keeps the explicit
        # im2col / col2im kernels (the gather writes bf16 operand tiles).  Measured on a B200, B = 256 (profiles/r02_conv_*):
        # CelebA 121.7 k -> 132.0 k samples/s, MultiMNIST 152.6 k -> 160.2 k with the col2im side implicit.
        self.implicit_conv = self.act_dtype == torch.bfloat16
        self.implicit_col2im = self.implicit_conv
        self.convt_merged = True
        # BatchNorm batch statistics of the encoder's conv layers from the conv GEMM's epilogue (col_sum / col_sumsq of
        # mvae_gemm) instead of a reduction pass over the stored pre-activations
        self.epilogue_bn_stats = os.environ.get("MVAE_CONV_EPI_STATS", "1") != "0"
        # the second modality's networks (attribute MLPs / GRU text encoder + decoder: many latency-sized launches) run on their
        # own stream beside the image networks
        self.mod_stream = torch.cuda.Stream(device=dev)
        self.reset_parameters()

    # ------------------------------------------------------------------ parameters
    def reference_keys(self, n_latents: int):
        raise NotImplementedError

    def P(self, key: str) -> torch.Tensor:
        l = self.layouts[key]
        return self.flat_params[l.offset:l.offset + l.numel]

    def G(self, key: str) -> torch.Tensor:
        l = self.layouts[key]
        return self.flat_grads[l.offset:l.offset + l.numel]

    def W(self, key: str) -> torch.Tensor:
        """GEMM-operand view of a weight in the activation dtype (bf16 mirror / fp32 master)."""
        if self.act_dtype == torch.float32:
            return self.P(key)
        l = self.layouts[key]
        return self.flat_params_bf16[l.offset:l.offset + l.numel]

    def running(self, bn: str) -> Tuple[torch.Tensor, torch.Tensor]:
        o, c = self.bn_off[bn], self.BN_LAYERS[bn]
        return self.flat_buffers[o:o + c], self.flat_buffers[o + c:o + 2 * c]

    def _init_tensor(self, key, shape, kind, g, sd):
        """PyTorch-default-like initialisation of one tensor (subclasses override for GRU / Embedding)."""
        if kind == "nbt":
            return torch.zeros((), dtype=torch.int64)
        if kind == "rm":
            return torch.zeros(shape)
        if kind == "rv":
            return torch.ones(shape)
        if key.rsplit(".", 1)[0] in self.BN_LAYERS:
            return torch.ones(shape) if key.endswith("weight") else torch.zeros(shape)
        if len(shape) == 4:
            fan_in = (shape[1] if kind == "conv" else shape[0]) * shape[2] * shape[3]
            return (torch.rand(shape, generator=g) * 2 - 1) / fan_in ** 0.5
        if len(shape) == 2:
            return (torch.rand(shape, generator=g) * 2 - 1) / shape[1] ** 0.5
        fan_in = sd[key[:-4] + "weight"].shape[1]
        return (torch.rand(shape, generator=g) * 2 - 1) / fan_in ** 0.5

    def reset_parameters(self, seed: int = 1234) -> None:
        g = torch.Generator().manual_seed(seed)
        sd: Dict[str, torch.Tensor] = {}
        for k, shp, kind in self.state_keys:
            sd[k] = self._init_tensor(k, shp, kind, g, sd)
        self.load_state_dict(sd)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """Reference-shaped copies under the reference's keys."""
        out: Dict[str, torch.Tensor] = {}
        for k, shp, kind in self.state_keys:
            bn = k.rsplit(".", 1)[0]
            if kind == "rm":
                out[k] = self.running(bn)[0].clone()
            elif kind == "rv":
                out[k] = self.running(bn)[1].clone()
            elif kind == "nbt":
                out[k] = self.flat_nbt[self.bn_index[bn]].clone()
            else:
                out[k] = self.layouts[k].to_reference(self.P(k))
        return out

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True) -> None:
        """Loads a reference checkpoint's state_dict (reference shapes; converted to the internal layouts)."""
        missing = [k for k, _, _ in self.state_keys if k not in sd]
        if strict and missing:
            raise KeyError("missing keys in state_dict: %s" % missing[:4])
        for k, shp, kind in self.state_keys:
            if k not in sd:
                continue
            t = sd[k].detach()
            if tuple(t.shape) != tuple(shp):
                raise ValueError("%s: shape %s, expected %s" % (k, tuple(t.shape), shp))
            bn = k.rsplit(".", 1)[0]
            if kind == "rm":
                self.running(bn)[0].copy_(t)
            elif kind == "rv":
                self.running(bn)[1].copy_(t)
            elif kind == "nbt":
                self.flat_nbt[self.bn_index[bn]] = int(t)
            else:
                self.P(k).copy_(self.layouts[k].to_internal(t.to(torch.float32)).reshape(-1))
        self.sync_operands()

    def grads_reference(self) -> Dict[str, torch.Tensor]:
        """Gradients in the reference's layout (tests / interop)."""
        return {k: l.to_reference(self.G(k)) for k, l in self.layouts.items()}

    def sync_operands(self) -> None:
        """Refresh the bf16 mirror after the fp32 master changed outside the fused Adam kernel."""
        if self.flat_params_bf16 is not None:
            _ops.cast_f32_to_bf16(self.flat_params, self.flat_params_bf16, self.param_floats)

    def train(self, mode: bool = True):
        self.training = bool(mode)
        return self

    def eval(self):
        return self.train(False)

    def cuda(self, *a, **k):
        return self

    def parameters(self):
        """One flat leaf (elementwise optimisers such as torch.optim.Adam do not care about the layout)."""
        return [self.param]

    def zero_grad(self, set_to_none: bool = False):
        self.flat_grads.zero_()
        self.param.grad = self.flat_grads

    # ------------------------------------------------------------------ module surface with autograd (train mode)
    def _autograd_forward(self, image, other, term, eps):
        """vae(image, other) usable in the reference's own training loop: outputs carry a grad_fn whose backward runs the
        C-ABI backward kernels and delivers ONE gradient for the flat parameter leaf."""
        return _ConvForwardFn.apply(self, image, other, term, eps, self.param)

    def module_outputs(self, ws):
        raise NotImplementedError

    def module_backward(self, ws, g_image, g_other, g_mu, g_logvar):
        raise NotImplementedError

    # ------------------------------------------------------------------ Linear helpers
    def operand(self, key: str, rows: int, cols: int) -> Tuple[torch.Tensor, int]:
        """GEMM operand of weight `key` [rows, cols] with a TMA-legal row stride: the bf16 mirror / fp32 master itself,
        or (row length not a 16-byte multiple) a zero-padded copy refreshed once per forward (begin_forward())."""
        ld = round_up(cols, self.vec)
        if ld == cols:
            return self.W(key), cols
        if key not in self._pad:
            self._pad[key] = (torch.zeros(rows * ld, device=self.device, dtype=self.act_dtype), ld)
        buf, _ = self._pad[key]
        if key not in self._fresh:
            _ops.cast_pad_2d(self.P(key), rows, cols, cols, buf, ld)
            self._fresh.add(key)
        return buf, ld

    def begin_forward(self) -> None:
        """The parameters may have changed since the last forward: padded operand copies must be rebuilt."""
        self._fresh.clear()

    def linear_fwd(self, x, ldx, M, prefix, n_out, n_in, out, ldo, col_off: int = 0, refresh: bool = True):
        """out[:, col_off:col_off+n_out] = x[M, n_in] W^T + b  (nn.Linear `prefix`)."""
        w, ldw = self.operand(prefix + ".weight", n_out, n_in) if refresh else self._operand_cached(prefix + ".weight", n_in)
        dst = out if col_off == 0 else out[col_off:]
        _ops.gemm(x, w, dst, M, n_out, n_in, ldx, ldw, ldo, bias=self.P(prefix + ".bias"))

    def _operand_cached(self, key: str, cols: int) -> Tuple[torch.Tensor, int]:
        if key in self._pad:
            return self._pad[key]
        return self.W(key), cols

    def linear_bwd(self, x, ldx, dy, lddy, M, prefix, n_out, n_in, dx=None, lddx=0, accumulate_dx=False, bias=True):
        """dW += dy^T x, db += colsum(dy), dx (=|+=) dy W.  dy is [M, n_out] in the activation dtype with a row stride that
        is a multiple of the vector width (columns >= n_out zero)."""
        _ops.gemm(dy, x, self.G(prefix + ".weight"), n_out, n_in, M, lddy, ldx, n_in, a_major=1, b_major=1, accumulate=True)
        if bias:
            _ops.col_stats(dy, M, lddy, self.G(prefix + ".bias"), valid_channels=n_out)
        if dx is not None:
            w, ldw = self._operand_cached(prefix + ".weight", n_in)
            _ops.gemm(dy, w, dx, M, n_in, n_out, lddy, ldw, lddx, b_major=1, accumulate=accumulate_dx)

    def on_mod_stream(self, fn) -> None:
        """Fork: run `fn` on the second-modality stream after everything enqueued so far on the current stream."""
        if not self.use_side_stream:
            fn()
            return
        self.mod_stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.mod_stream):
            fn()
        self._mod_forked = True

    def join_mod_stream(self) -> None:
        if self.use_side_stream and getattr(self, "_mod_forked", False):
            torch.cuda.current_stream(self.device).wait_stream(self.mod_stream)
            self._mod_forked = False

    def _wgrad_aside(self, fn) -> None:
        """Run `fn` (weight-gradient launches) on the side stream after everything enqueued so far on the current stream."""
        if not self.use_side_stream:
            fn()
            return
        self.side_stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.side_stream):
            fn()
        self._side_forked = True

    def _join_side(self) -> None:
        # only after a fork: waiting on a side stream that holds no work of this step would, under CUDA-graph capture,
        # create a dependency on uncaptured work
        if self.use_side_stream and getattr(self, "_side_forked", False):
            torch.cuda.current_stream(self.device).wait_stream(self.side_stream)
            self._side_forked = False

    # ------------------------------------------------------------------ conv stacks
    def _implicit(self, channels: int, channels_last: bool, pixels: int, elements: int) -> bool:
        """Gather the patch matrix inside the GEMM?  (bf16 storage, channels-last source whose channel count is a multiple
        of 8; the encoder's first layer reads the caller's NCHW fp32 image and keeps the explicit im2col.)  `pixels` =
        rows of the patch matrix, `elements` = size of the gathered tensor: the kernel's 32-bit offsets and multiply-shift
        division cover pixels < 2^24 and elements < 2^31 (mvae_conv_gemm refuses more) - beyond that the explicit path runs."""
        return (self.implicit_conv and channels_last and channels % 8 == 0 and pixels < (1 << 24) and elements < (1 << 31))

    def _implicit_t(self, channels: int, out_channels: int, exact: bool, pixels: int, elements: int) -> bool:
        """Transposed convolution / Conv2d input gradient without the patch matrix?  (channels % 64: a K block of the class
        GEMM never straddles a tap; `exact`: the transposed output covers the whole image.)"""
        return (self.implicit_col2im and exact and channels % 64 == 0 and out_channels % 8 == 0 and pixels < (1 << 24)
                and elements < (1 << 31))

    def features_fwd(self, ws, image, B, training: bool, updates: int) -> None:
        """The image encoder's conv stack (celeba/model.py:101-113, multimnist/model.py:159-171): ws.enc_act[-1] is the
        NHWC bottleneck [B, FLAT_HW * FLAT_C]."""
        src = image
        epi_stats = training and self.epilogue_bn_stats
        if epi_stats:
            ws.enc_stats.zero_()
        for li, (pre, ci, co, k, s, p, hin, bn) in enumerate(self.ENC_CONVS):
            ho = _ops.out_size(hin, k, s, p)
            K = k * k * ci
            ldk = ws.enc_ldk[li]
            strides = _ops.nchw_strides(self.IMG_C, self.IMG_H, self.IMG_H) if li == 0 else None
            g = _ops.geometry(B, hin, hin, ci, k, s, p, strides)
            rows = B * ho * ho
            w, ldw = self.operand(pre + ".weight", co, K)
            # BatchNorm batch statistics (celeba/model.py:104-111) = column sums of the conv GEMM's output: from its epilogue
            st = dict(col_sum=ws.enc_sum[li], col_sumsq=ws.enc_sumsq[li], rows_per_group=rows) if (bn and epi_stats) else {}
            if self._implicit(ci, li > 0, rows, B * hin * hin * ci):
                _ops.gemm(src, w, ws.enc_pre[li], rows, co, K, 0, ldw, co, patch=(g, 1), **st)
            else:
                _ops.im2col(g, src, ws.enc_col[li], ldk)
                _ops.gemm(ws.enc_col[li], w, ws.enc_pre[li], rows, co, K, ldk, ldw, co, **st)
            if bn:
                rm, rv = self.running(bn)
                a = _ops.bn_args(ws.enc_pre[li], rows, co, rows, SWISH, training, self.P(bn + ".weight"), self.P(bn + ".bias"),
                                 ws.enc_sum[li], ws.enc_sumsq[li], ws.enc_mean[li], ws.enc_rstd[li], rm, rv, updates=updates,
                                 stats_ready=bool(st))
                _ops.bn_act_forward(a, ws.enc_act[li])
            else:
                _ops.act_forward(SWISH, ws.enc_pre[li], ws.enc_act[li], rows, co)
            src = ws.enc_act[li]

    def features_bwd(self, ws, B) -> None:
        """Backward of features_fwd from ws.enc_dact[-1] (gradient at the bottleneck)."""
        for li in range(len(self.ENC_CONVS) - 1, -1, -1):
            pre, ci, co, k, s, p, hin, bn = self.ENC_CONVS[li]
            ho = _ops.out_size(hin, k, s, p)
            K = k * k * ci
            ldk = ws.enc_ldk[li]
            rows = B * ho * ho
            if bn:
                a = _ops.bn_args(ws.enc_pre[li], rows, co, rows, SWISH, True, self.P(bn + ".weight"), self.P(bn + ".bias"),
                                 None, None, ws.enc_mean[li], ws.enc_rstd[li])
                _ops.bn_act_backward(a, ws.enc_dact[li], ws.enc_dpre[li], ws.enc_s0[li], ws.enc_s1[li], self.G(bn + ".weight"),
                                     self.G(bn + ".bias"))
            else:
                _ops.act_backward(SWISH, ws.enc_pre[li], ws.enc_dact[li], ws.enc_dpre[li], rows, co)
            # dW'[co, K] += dpre^T col   (side stream: overlaps the dgrad GEMM / col2im / BatchNorm backward of the next layer)
            if self._implicit(ci, li > 0, rows, B * hin * hin * ci):
                gi = _ops.geometry(B, hin, hin, ci, k, s, p)
                self._wgrad_aside(lambda li=li, pre=pre, co=co, K=K, rows=rows, gi=gi: _ops.gemm(
                    ws.enc_dpre[li], ws.enc_act[li - 1], self.G(pre + ".weight"), co, K, rows, co, 0, K, a_major=1, b_major=1,
                    accumulate=True, patch=(gi, 2)))
            else:
                self._wgrad_aside(lambda li=li, pre=pre, co=co, K=K, rows=rows, ldk=ldk: _ops.gemm(
                    ws.enc_dpre[li], ws.enc_col[li], self.G(pre + ".weight"), co, K, rows, co, ldk, K, a_major=1, b_major=1,
                    accumulate=True))
            if li > 0:
                w, ldw = self._operand_cached(pre + ".weight", K)
                if ldw == K and self._implicit_t(co, ci, (hin + 2 * p - k) % s == 0, rows, rows * co):
                    # the input gradient IS a transposed convolution of dpre with the same weights [co, kh, kw, ci]
                    _ops.transposed_conv_implicit(ws.enc_dpre[li], w, ws.enc_dact[li - 1], B, ho, co, ci, k, s, p,
                                                  merged=self.convt_merged)
                else:
                    _ops.gemm(ws.enc_dpre[li], w, ws.colbuf, rows, K, co, co, ldw, ldk, b_major=1)   # dcol = dpre W'
                    g = _ops.geometry(B, hin, hin, ci, k, s, p)
                    _ops.col2im(g, ws.colbuf, ldk, ws.enc_dact[li - 1])
        self._join_side()

    def hallucinate_fwd(self, ws, M3, rows_per_term, training: bool) -> None:
        """The image decoder's transposed-conv stack on ws.u1 [M3, FLAT_HW*FLAT_C] -> ws.logits (NCHW fp32)."""
        src = ws.u1
        last = len(self.DEC_CONVS) - 1
        for li, (pre, ci, co, k, s, p, hout, bn) in enumerate(self.DEC_CONVS):
            hin = _ops.out_size(hout, k, s, p)
            K = k * k * co
            ldk = ws.dec_ldk[li]
            rows_in = M3 * hin * hin
            # col[M_in, (kh,kw,co)] = X[M_in, ci] * W'[ci, (kh,kw,co)]
            w, ldw = self.operand(pre + ".weight", ci, K)
            direct = li < last and ldw == K and self._implicit_t(ci, co, True, rows_in, rows_in * ci)
            if direct:
                _ops.transposed_conv_implicit(src, w, ws.dec_pre[li], M3, hin, ci, co, k, s, p, merged=self.convt_merged)
            else:
                _ops.gemm(src, w, ws.colbuf, rows_in, K, ci, ci, ldw, ldk, b_major=1)
            if li < last:
                if not direct:
                    g = _ops.geometry(M3, hout, hout, co, k, s, p)
                    _ops.col2im(g, ws.colbuf, ldk, ws.dec_pre[li])
                rows = M3 * hout * hout
                rm, rv = self.running(bn)
                a = _ops.bn_args(ws.dec_pre[li], rows, co, rows_per_term * hout * hout, SWISH, training, self.P(bn + ".weight"),
                                 self.P(bn + ".bias"), ws.dec_sum[li], ws.dec_sumsq[li], ws.dec_mean[li], ws.dec_rstd[li],
                                 rm, rv, updates=1)
                ws.dec_bn[li] = a
                _ops.bn_act_forward(a, ws.dec_act[li])
                src = ws.dec_act[li]
            else:
                g = _ops.geometry(M3, hout, hout, co, k, s, p, _ops.nchw_strides(self.IMG_C, self.IMG_H, self.IMG_H))
                _ops.col2im(g, ws.colbuf, ldk, ws.logits)

    def hallucinate_bwd(self, ws, M3) -> None:
        """Backward of hallucinate_fwd: the gradient at the logits sits in ws.logits (NCHW fp32); leaves ws.du1."""
        dsrc = ws.logits
        last = len(self.DEC_CONVS) - 1
        for li in range(last, -1, -1):
            pre, ci, co, k, s, p, hout, bn = self.DEC_CONVS[li]
            hin = _ops.out_size(hout, k, s, p)
            K = k * k * co
            ldk = ws.dec_ldk[li]
            rows_in = M3 * hin * hin
            strides = _ops.nchw_strides(self.IMG_C, self.IMG_H, self.IMG_H) if li == last else None
            g = _ops.geometry(M3, hout, hout, co, k, s, p, strides)
            if li < last:
                _ops.bn_act_backward(ws.dec_bn[li], dsrc, ws.dec_dpre[li], ws.dec_s0[li], ws.dec_s1[li], self.G(bn + ".weight"),
                                     self.G(bn + ".bias"))
                dsrc = ws.dec_dpre[li]
            x_in = ws.dec_act[li - 1] if li > 0 else ws.u1             # the layer's input [M_in, ci]
            dx = ws.dec_dact[li - 1] if li > 0 else ws.du1
            w, ldw = self._operand_cached(pre + ".weight", K)
            if self._implicit(co, li < last, rows_in, M3 * hout * hout * co):
                # dcol = im2col(dOut) is never written: gathered as the B operand of the weight gradient and as the A
                # operand of the input gradient
                self._wgrad_aside(lambda x_in=x_in, pre=pre, ci=ci, K=K, rows_in=rows_in, g=g, dsrc=dsrc: _ops.gemm(
                    x_in, dsrc, self.G(pre + ".weight"), ci, K, rows_in, ci, 0, K, a_major=1, b_major=1, accumulate=True,
                    patch=(g, 2)))
                _ops.gemm(dsrc, w, dx, rows_in, ci, K, 0, ldw, ci, patch=(g, 1))
            else:
                self._join_side()                                      # the previous layer's weight gradient still reads colbuf
                _ops.im2col(g, dsrc, ws.colbuf, ldk)                   # dcol [M_in, (kh,kw,co)]
                self._wgrad_aside(lambda x_in=x_in, pre=pre, ci=ci, K=K, rows_in=rows_in, ldk=ldk: _ops.gemm(
                    x_in, ws.colbuf, self.G(pre + ".weight"), ci, K, rows_in, ci, ldk, K, a_major=1, b_major=1, accumulate=True))
                _ops.gemm(ws.colbuf, w, dx, rows_in, ci, K, ldk, ldw, ci)
            dsrc = dx
        self._join_side()

    def alloc_conv_buffers(self, ws, B, G) -> None:
        """Activation / gradient buffers of the two conv stacks for batch B and G stacked terms."""
        dev, T = self.device, self.act_dtype
        f32 = torch.float32
        M3 = G * B

        def buf(*shape, dtype=T):
            return torch.zeros(*shape, device=dev, dtype=dtype)

        ws.buf = buf
        ws.enc_col, ws.enc_pre, ws.enc_act, ws.enc_dact, ws.enc_dpre, ws.enc_ldk = [], [], [], [], [], []
        ws.enc_sum, ws.enc_sumsq, ws.enc_mean, ws.enc_rstd, ws.enc_s0, ws.enc_s1 = [], [], [], [], [], []
        colmax = 0
        for li, (pre, ci, co, k, s, p, hin, bn) in enumerate(self.ENC_CONVS):
            ho = _ops.out_size(hin, k, s, p)
            rows = B * ho * ho
            ldk = round_up(k * k * ci, self.vec)
            ws.enc_ldk.append(ldk)
            ws.enc_col.append(None if self._implicit(ci, li > 0, rows, B * hin * hin * ci) else buf(rows * ldk))
            ws.enc_pre.append(buf(rows * co))
            ws.enc_act.append(buf(rows * co))
            ws.enc_dact.append(buf(rows * co))
            ws.enc_dpre.append(buf(rows * co))
            for lst in (ws.enc_sum, ws.enc_sumsq, ws.enc_mean, ws.enc_rstd, ws.enc_s0, ws.enc_s1):
                lst.append(buf(co, dtype=f32))
            if li > 0:
                colmax = max(colmax, rows * ldk)
        # the forward statistics of the encoder's BatchNorm layers come out of the conv GEMMs' epilogues (column sums added
        # with atomics): their accumulators live in ONE buffer that a single memset clears
        tot = sum(co for (_, _, co, *_r) in self.ENC_CONVS)
        ws.enc_stats = buf(2 * tot, dtype=f32)
        off = 0
        for li, (pre, ci, co, *_r) in enumerate(self.ENC_CONVS):
            ws.enc_sum[li] = ws.enc_stats[off:off + co]
            ws.enc_sumsq[li] = ws.enc_stats[tot + off:tot + off + co]
            off += co
        F = self.flat_feat
        ws.u1pre, ws.u1, ws.du1, ws.du1pre = buf(M3 * F), buf(M3 * F), buf(M3 * F), buf(M3 * F)
        ws.dec_pre, ws.dec_act, ws.dec_dact, ws.dec_dpre, ws.dec_ldk = [], [], [], [], []
        ws.dec_sum, ws.dec_sumsq, ws.dec_mean, ws.dec_rstd, ws.dec_s0, ws.dec_s1 = [], [], [], [], [], []
        ws.dec_bn = [None] * len(self.DEC_CONVS)
        for li, (pre, ci, co, k, s, p, hout, bn) in enumerate(self.DEC_CONVS):
            hin = _ops.out_size(hout, k, s, p)
            ldk = round_up(k * k * co, self.vec)
            ws.dec_ldk.append(ldk)
            colmax = max(colmax, M3 * hin * hin * ldk)
            rows = M3 * hout * hout
            if li < len(self.DEC_CONVS) - 1:
                ws.dec_pre.append(buf(rows * co))
                ws.dec_act.append(buf(rows * co))
                ws.dec_dact.append(buf(rows * co))
                ws.dec_dpre.append(buf(rows * co))
            for lst in (ws.dec_sum, ws.dec_sumsq, ws.dec_mean, ws.dec_rstd, ws.dec_s0, ws.dec_s1):
                lst.append(buf(G, co, dtype=f32))
        ws.colbuf = buf(colmax)
        ws.logits = buf(M3 * self.n_pixels, dtype=f32)
        ws.probs_image = buf(M3 * self.n_pixels, dtype=f32)
        # latent
        n = self.n_latents
        ws.ld_z = round_up(n, self.vec)
        ws.ld_enc = round_up(2 * n, self.vec)
        ws.z = buf(M3 * ws.ld_z)
        ws.dz = buf(M3 * n, dtype=f32)
        ws.mu, ws.logvar = buf(M3 * n, dtype=f32), buf(M3 * n, dtype=f32)
        ws.acc = buf(3, 4, dtype=f32)   # rows: image BCE sums, second-modality loss sums, weighted KL; columns: term

    # ------------------------------------------------------------------ latent path
    def latent_forward(self, ws, term_types, kl_weights, eps, training, enc_a, enc_b, R) -> None:
        B, n, G = ws.B, self.n_latents, len(term_types)
        la = _lib.LatentArgs()
        la.batch, la.n_latents, la.n_terms = B, n, G
        img_seen = 0
        for gi, t in enumerate(term_types):
            la.term_type[gi] = t
            la.kl_weight[gi] = float(kl_weights[gi])
            la.expert_a_row0[gi] = 0
            if t != _lib.TERM_TEXT:
                la.expert_a_row0[gi] = (img_seen * B) if R > 1 else 0
                img_seen += 1
        la.poe_mode, la.prior_expert, la.poe_eps = self.poe_mode, self.prior_expert, self.poe_eps
        if enc_a is not None:
            la.expert_a, la.ld_a = enc_a.data_ptr(), 2 * n
        if enc_b is not None:
            la.expert_b, la.ld_b = enc_b.data_ptr(), 2 * n
        ws.keep_eps = eps                          # the backward re-reads the noise through this raw pointer
        la.eps = None if eps is None else eps.data_ptr()
        la.seed, la.step_counter = self.noise_seed, self._step_counter.data_ptr()
        la.training = 1 if training else 0
        rw = getattr(ws, "row_weight", None)                 # per-(term, row) weights of a masked step, or None
        la.row_weight = None if rw is None else rw.data_ptr()
        la.z_dtype, la.z, la.ld_z = _ops.DT[self.act_dtype], ws.z.data_ptr(), ws.ld_z
        la.mu, la.logvar, la.kl = ws.mu.data_ptr(), ws.logvar.data_ptr(), ws.acc[2].data_ptr()
        ws.latent = la
        _lib.check(_lib.load().mvae_latent_forward(C.byref(la), _ops.stream()), "mvae_latent_forward")

    def latent_backward(self, ws, denc_a, denc_b, d_mu=None, d_logvar=None) -> None:
        la = ws.latent
        ws.keep_up = (d_mu, d_logvar)            # raw pointers below: keep the tensors alive
        la.d_mu = None if d_mu is None else d_mu.data_ptr()
        la.d_logvar = None if d_logvar is None else d_logvar.data_ptr()
        la.dz_dtype, la.dz, la.ld_dz = _lib.DT_F32, ws.dz.data_ptr(), self.n_latents
        la.d_dtype = _ops.DT[self.act_dtype]
        la.d_expert_a, la.ld_da = (None if denc_a is None else denc_a.data_ptr()), ws.ld_enc
        la.d_expert_b, la.ld_db = (None if denc_b is None else denc_b.data_ptr()), ws.ld_enc
        _lib.check(_lib.load().mvae_latent_backward(C.byref(la), _ops.stream()), "mvae_latent_backward")

    def workspace(self, batch: int, n_terms: int):
        key = (int(batch), int(n_terms))
        if key not in self._ws:
            self._ws[key] = self._make_workspace(int(batch), int(n_terms))
        return self._ws[key]

    # subclass hooks
    def _make_workspace(self, B, G):
        raise NotImplementedError

    def run_forward(self, ws, image, other, term_types, eps, training, lambdas, kl_weights, want_probs, with_loss):
        raise NotImplementedError

    def backward_decoders(self, ws):
        """Decoder-side backward + latent backward: fills every decoder gradient (the first all-reduce bucket)."""
        raise NotImplementedError

    def backward_encoders(self, ws):
        raise NotImplementedError

    def bn_increments(self, term_types) -> List[int]:
        raise NotImplementedError


class Workspace:
    pass


class _ConvForwardFn(torch.autograd.Function):
    """Autograd bridge of MultimodalVAE.forward for the conv models: forward = the forward kernels on a PRIVATE workspace
    (the reference's loop keeps three forwards alive before one backward), backward = the backward kernels fed with the
    upstream gradients of (image_recon, second recon, mu, logvar); the gradient of the flat parameter leaf is returned."""

    @staticmethod
    def forward(ctx, model, image, other, term, eps, param):
        B = (image if image is not None else other).shape[0]
        model.sync_operands()                      # the parameters may have been stepped by a torch optimiser
        inc = torch.tensor(model.bn_increments((term,)), dtype=torch.int64, device=model.device)
        _ops.step_begin(model._step_counter, None, model.flat_nbt, inc)   # noise / dropout counter, num_batches_tracked
        ws = model._make_workspace(B, 1)
        ws.keep_inc = inc
        model.run_forward(ws, image, other, (term,), eps, True, ((0.0, 0.0),), (0.0,), True, False)
        ctx.model, ctx.ws = model, ws
        return model.module_outputs(ws)

    @staticmethod
    def backward(ctx, g_image, g_other, g_mu, g_logvar):
        m, ws = ctx.model, ctx.ws

        def c(t):
            return None if t is None else t.to(torch.float32).contiguous()

        g = m.param.grad
        if g is not None and g.data_ptr() == m.flat_grads.data_ptr():
            # param.grad IS the library's flat gradient buffer (model.zero_grad()): every kernel accumulates (+=) into it
            # directly - autograd's AccumulateGrad would otherwise replace the aliased tensor by an out-of-place sum
            m.module_backward(ws, c(g_image), c(g_other), c(g_mu), c(g_logvar))
            return None, None, None, None, None, None
        grads = torch.zeros_like(m.flat_params)     # e.g. after optimizer.zero_grad(set_to_none=True)
        saved = m.flat_grads
        m.flat_grads = grads
        try:
            m.module_backward(ws, c(g_image), c(g_other), c(g_mu), c(g_logvar))
        finally:
            m.flat_grads = saved
        return None, None, None, None, None, grads


class ConvMVAETrainer:
    """zero_grad, the three forwards, the three loss_function calls, backward, Adam (celeba/train.py:132-157,
    multimnist/train.py:148-175) as one stream of kernels; data-parallel when torch.distributed is initialised:
    the decoder gradient bucket is all-reduced on a side stream while the encoder backward runs."""

    def __init__(self, model: ConvMVAEBase, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, kl_lambda: float = 1e-3,
                 use_cuda_graph: bool = False, group=None, overlap: bool = True):
        self.model = model
        self.lr, self.betas, self.eps, self.kl_lambda = float(lr), betas, float(eps), float(kl_lambda)
        self.adam_m = torch.zeros_like(model.flat_params)
        self.adam_v = torch.zeros_like(model.flat_params)
        self.use_cuda_graph = use_cuda_graph
        self._graphs: Dict[Tuple, Tuple] = {}
        self._inc_cache: Dict[Tuple, torch.Tensor] = {}
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.overlap = overlap
        self.comm_stream = torch.cuda.Stream(device=model.device) if self.world > 1 else None
        if self.world > 1:
            dist.broadcast(model.flat_params, 0, group=group)
            dist.broadcast(model.flat_buffers, 0, group=group)
            dist.broadcast(model.flat_nbt, 0, group=group)
            model.sync_operands()

    def _increments(self, term_types) -> torch.Tensor:
        key = tuple(term_types)
        if key not in self._inc_cache:
            self._inc_cache[key] = torch.tensor(self.model.bn_increments(term_types), dtype=torch.int64, device=self.model.device)
        return self._inc_cache[key]

    def _enqueue(self, ws, image, other, term_types, lambdas, eps, adam: bool) -> None:
        m = self.model
        B = ws.B
        # everything that accumulates over the step is zeroed here: the loss sums (ws.acc) or a workspace-defined region around them
        _ops.step_begin(m._step_counter, getattr(ws, "zero_region", ws.acc.view(-1)), m.flat_nbt, self._increments(term_types))
        klw = [self.kl_lambda / B] * len(term_types)
        m.run_forward(ws, image, other, term_types, eps, True, lambdas, klw, False, True)
        m.backward_decoders(ws)
        split = m.encoder_param_floats
        if getattr(self, "_defer_reduce", False):
            m.backward_encoders(ws)       # accumulated step (step_masked): the caller reduces once, after the last class
        elif self.world > 1 and self.overlap:
            main = torch.cuda.current_stream(m.device)
            self.comm_stream.wait_stream(main)
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(m.flat_grads[split:], op=dist.ReduceOp.SUM, group=self.group)
            m.backward_encoders(ws)
            dist.all_reduce(m.flat_grads[:split], op=dist.ReduceOp.SUM, group=self.group)
            main.wait_stream(self.comm_stream)
        else:
            m.backward_encoders(ws)
            if self.world > 1:
                dist.all_reduce(m.flat_grads, op=dist.ReduceOp.SUM, group=self.group)
        if adam:
            _ops.step_begin(m._adam_counter)     # one optimizer step = one tick of Adam's clock (forwards never tick it)
            _ops.adam_step(m.flat_params, m.flat_grads, self.adam_m, self.adam_v, m.flat_params_bf16, m.param_floats, self.lr,
                           self.betas[0], self.betas[1], self.eps, m._adam_counter, 1.0 / self.world, True)

    def _prepare(self, image, other):
        m = self.model
        return image.to(m.device, torch.float32).contiguous(), other

    def step(self, image: torch.Tensor, other: torch.Tensor, terms: Sequence[str], lambdas, eps: Optional[torch.Tensor] = None,
             adam: bool = True):
        m = self.model
        tt = tuple(m.TERMS[t] for t in terms)
        lambdas = tuple(tuple(float(v) for v in l) for l in lambdas)
        B = image.shape[0]
        ws = m.workspace(B, len(tt))
        image, other = self._prepare(image, other)
        if eps is not None:
            eps = eps.to(m.device, torch.float32).contiguous()
        self._last = (ws, tt, lambdas)
        if not self.use_cuda_graph:
            self._enqueue(ws, image, other, tt, lambdas, eps, adam)
            return ws.acc
        # every scalar baked into the captured kernel arguments is part of the key (anneal_kl / adjust_learning_rate of
        # multimnist/train.py:227-241 mutate kl_lambda / lr between steps); the cache is bounded
        key = (B, tt, lambdas, eps is not None, adam, float(self.kl_lambda), float(self.lr), tuple(map(float, self.betas)),
               float(self.eps), getattr(self, "_key_extra", None))
        if key not in self._graphs:
            while len(self._graphs) >= 16:
                self._graphs.pop(next(iter(self._graphs)))
            st_img, st_oth = image.clone(), other.clone()
            st_eps = None if eps is None else eps.clone()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                # warm-up outside capture (allocates padded operand copies, sets up NCCL), state restored afterwards
                state = (m.flat_params, m.flat_buffers, m.flat_nbt, m._counters, self.adam_m, self.adam_v, m.flat_grads)
                snap = [t.clone() for t in state]
                self._enqueue(ws, st_img, st_oth, tt, lambdas, st_eps, adam)
                for dst, src in zip(state, snap):
                    dst.copy_(src)
                m.sync_operands()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            lib = _lib.load()
            before = lib.mvae_launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue(ws, st_img, st_oth, tt, lambdas, st_eps, adam)
            self._graphs[key] = (g, st_img, st_oth, st_eps, int(lib.mvae_launch_count() - before))
        g, st_img, st_oth, st_eps, self.last_graph_launches = self._graphs[key]
        st_img.copy_(image, non_blocking=True)
        st_oth.copy_(other, non_blocking=True)
        if st_eps is not None:
            st_eps.copy_(eps, non_blocking=True)
        g.replay()
        return ws.acc

    # (class, term indices into list(model.TERMS), lambdas): celeba has no paired_weak script; these are the per-class loss calls
    # of multimnist/paired_weak.py:85-110 (paired rows: three terms; unpaired rows: each modality's own term)
    _MASK_CLASSES = (("paired", (0, 1, 2), ((1.0, 1.0), (1.0, 1.0), (0.0, 1.0))),
                     ("image_only", (1,), ((1.0, 0.0),)),
                     ("other_only", (2,), ((0.0, 1.0),)))

    def step_masked(self, image, other, has_image, has_other, eps: Optional[torch.Tensor] = None, update: bool = True):
        """One optimizer step over a batch that MIXES paired and unpaired samples (the reference flips a coin per
        batch, multimnist/paired_weak.py:85-110).  `has_image`, `has_other`: [B] bool.  Rows are compacted by presence class on
        the device and each class runs the step on its own rows - paired rows: joint + image + second-modality terms, image-only
        rows: the image term, other-only rows: the second modality's term; rows with neither are skipped - so every loss is a mean
        over its own rows and BatchNorm only ever sees present rows, exactly what the reference computes when it is fed the three
        subsets as three batches; the gradients accumulate, the data-parallel all-reduce and ONE Adam update follow.  Eager (the
        class sizes are data dependent and train-mode BatchNorm statistics need compacted rows: one host sync per class); classes
        with fewer than two rows are skipped (BatchNorm).  `eps`: optional [3, B, n].  Returns {class: losses()}."""
        m = self.model
        names = list(m.TERMS)
        dev = m.device
        hi = torch.as_tensor(has_image).to(dev).bool().reshape(-1)
        ho = torch.as_tensor(has_other).to(dev).bool().reshape(-1)
        image, other = self._prepare(image, other)
        if hi.numel() != image.shape[0] or ho.numel() != image.shape[0]:
            raise ValueError("has_image / has_other must have one entry per sample")
        if eps is not None:
            eps = eps.to(dev, torch.float32)
        rows = {"paired": hi & ho, "image_only": hi & ~ho, "other_only": ~hi & ho}
        out = {}
        m.flat_grads.zero_()
        self._defer_reduce = True
        try:
            for cname, term_idx, lambdas in self._MASK_CLASSES:
                idx = torch.nonzero(rows[cname]).reshape(-1)
                if idx.numel() < 2:
                    continue
                e = None if eps is None else torch.stack([eps[t].index_select(0, idx) for t in term_idx]).contiguous()
                tt = tuple(m.TERMS[names[t]] for t in term_idx)
                ws = m._make_workspace(int(idx.numel()), len(tt))   # private: the class sizes change from step to step
                self._last = (ws, tt, lambdas)
                self._enqueue(ws, image.index_select(0, idx).contiguous(), other.index_select(0, idx).contiguous(), tt, lambdas, e,
                              adam=False)
                out[cname] = self.losses()
        finally:
            self._defer_reduce = False
        if self.world > 1:
            dist.all_reduce(m.flat_grads, op=dist.ReduceOp.SUM, group=self.group)
        if update and out:
            _ops.step_begin(m._adam_counter)
            _ops.adam_step(m.flat_params, m.flat_grads, self.adam_m, self.adam_v, m.flat_params_bf16, m.param_floats, self.lr,
                           self.betas[0], self.betas[1], self.eps, m._adam_counter, 1.0 / self.world, True)
        return out

    def teardown(self) -> None:
        self._graphs.clear()
