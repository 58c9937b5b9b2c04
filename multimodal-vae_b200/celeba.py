"""CelebA MVAE (64x64 conv image encoder / decoder + 18 binary attributes) on the B200-native library.

Reference surface kept (celeba/model.py:13-58, celeba/train.py:60-81,132-157):
    MultimodalVAE(n_latents=20, use_cuda=False).forward(image=None, attrs=None)
        -> (image_recon [B,3,64,64] probs, attrs_recon [B,18] probs, mu, logvar)
    state_dict()/load_state_dict() with the reference's keys and shapes (image_encoder.features.0.weight, ...)
    CelebATrainer.step(image, attrs): zero_grad + vae(image, attrs) + vae(image) + vae(attrs) + three
        loss_function calls + backward + Adam, as one stream of hand-written sm_100a kernels.

How it runs: every Conv2d / ConvTranspose2d / Linear is a tcgen05 GEMM (mvae_gemm) over NHWC activations; patch
matrices come from mvae_im2col / mvae_col2im; BatchNorm+Swish, Swish(+Dropout), sigmoid+BCE and the PoE /
reparametrize / KL latent path are fused kernels (csrc/conv_ops.cu).  The image encoder and the attribute encoder
run ONCE per step (the reference runs each twice on identical inputs; only Dropout differs between its two image
passes, so everything up to the Dropout is shared and the two masks are applied to replicated rows); the decoders
run once on the stacked [3B] latents with per-term BatchNorm statistics.

Internal parameter layout: conv weights [Cout, kh, kw, Cin], transposed-conv weights [Cin, kh, kw, Cout], and the
two Linear layers adjacent to the NCHW flatten (classifier.0, upsample.0) permuted to the NHWC feature order.
state_dict()/load_state_dict() convert to / from the reference layout.  There is no PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib, _ops

N_ATTRS = 18  # celeba/datasets.py:27
TERMS = {"joint": _lib.TERM_JOINT, "image": _lib.TERM_IMAGE, "attrs": _lib.TERM_TEXT}
_DTYPES = {"tf32": torch.float32, "fp32": torch.float32, "bf16": torch.bfloat16}
SWISH = _lib.ACT_SWISH

# (key prefix, Cin, Cout, stride, pad, input H, BatchNorm prefix or None)  celeba/model.py:101-113
ENC_CONVS = (("image_encoder.features.0", 3, 32, 2, 1, 64, None),
             ("image_encoder.features.2", 32, 64, 2, 1, 32, "image_encoder.features.3"),
             ("image_encoder.features.5", 64, 128, 2, 1, 16, "image_encoder.features.6"),
             ("image_encoder.features.8", 128, 256, 1, 0, 8, "image_encoder.features.9"))
# (key prefix, Cin, Cout, stride, pad, OUTPUT H, BatchNorm prefix or None)  celeba/model.py:142-152
DEC_CONVS = (("image_decoder.hallucinate.0", 256, 128, 1, 0, 8, "image_decoder.hallucinate.1"),
             ("image_decoder.hallucinate.3", 128, 64, 2, 1, 16, "image_decoder.hallucinate.4"),
             ("image_decoder.hallucinate.6", 64, 32, 2, 1, 32, "image_decoder.hallucinate.7"),
             ("image_decoder.hallucinate.9", 32, 3, 2, 1, 64, None))
BN_LAYERS = ("image_encoder.features.3", "image_encoder.features.6", "image_encoder.features.9",
             "image_decoder.hallucinate.1", "image_decoder.hallucinate.4", "image_decoder.hallucinate.7",
             "attrs_encoder.net.1", "attrs_decoder.net.1")
_BN_CH = {"image_encoder.features.3": 64, "image_encoder.features.6": 128, "image_encoder.features.9": 256,
          "image_decoder.hallucinate.1": 128, "image_decoder.hallucinate.4": 64, "image_decoder.hallucinate.7": 32,
          "attrs_encoder.net.1": 64, "attrs_decoder.net.1": 64}


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# ---------------------------------------------------------------------------------------------- layouts
class _Layout:
    """How one reference tensor is held inside the flat buffer."""

    def __init__(self, key, ref_shape, kind):
        self.key, self.ref_shape, self.kind = key, tuple(ref_shape), kind  # kind: conv | convT | fc_in | fc_out | fc_out_bias | plain
        self.numel = 1
        for s in ref_shape:
            self.numel *= s
        self.offset = -1

    def to_internal(self, t: torch.Tensor) -> torch.Tensor:
        k = self.kind
        if k == "conv" or k == "convT":     # [a, b, kh, kw] -> [a, kh, kw, b]
            return t.permute(0, 2, 3, 1).contiguous()
        if k == "fc_in":                    # Linear(256*5*5, 1024): columns (c, hw) -> (hw, c)
            o = t.shape[0]
            return t.reshape(o, 256, 25).permute(0, 2, 1).contiguous().reshape(o, 6400)
        if k == "fc_out":                   # Linear(n, 256*5*5): rows (c, hw) -> (hw, c)
            n = t.shape[1]
            return t.reshape(256, 25, n).permute(1, 0, 2).contiguous().reshape(6400, n)
        if k == "fc_out_bias":
            return t.reshape(256, 25).t().contiguous().reshape(6400)
        return t.contiguous()

    def to_reference(self, t: torch.Tensor) -> torch.Tensor:
        k = self.kind
        if k == "conv" or k == "convT":
            a, b, kh, kw = self.ref_shape
            return t.reshape(a, kh, kw, b).permute(0, 3, 1, 2).contiguous()
        if k == "fc_in":
            o = self.ref_shape[0]
            return t.reshape(o, 25, 256).permute(0, 2, 1).contiguous().reshape(o, 6400)
        if k == "fc_out":
            n = self.ref_shape[1]
            return t.reshape(25, 256, n).permute(1, 0, 2).contiguous().reshape(6400, n)
        if k == "fc_out_bias":
            return t.reshape(25, 256).t().contiguous().reshape(6400)
        return t.reshape(self.ref_shape).clone()


def reference_keys(n_latents: int) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(key, reference shape, kind) in the reference's state_dict order (celeba/model.py:16-21)."""
    n = n_latents
    out: List[Tuple[str, Tuple[int, ...], str]] = []

    def bn(p, c):
        out.extend([(p + ".weight", (c,), "plain"), (p + ".bias", (c,), "plain"), (p + ".running_mean", (c,), "rm"),
                    (p + ".running_var", (c,), "rv"), (p + ".num_batches_tracked", (), "nbt")])

    def lin(p, o, i, wkind="plain", bkind="plain"):
        out.extend([(p + ".weight", (o, i), wkind), (p + ".bias", (o,), bkind)])

    for pre, ci, co, _, _, _, b in ENC_CONVS:
        out.append((pre + ".weight", (co, ci, 4, 4), "conv"))
        if b:
            bn(b, co)
    lin("image_encoder.classifier.0", 1024, 6400, "fc_in")
    lin("image_encoder.classifier.3", 2 * n, 1024)
    lin("image_decoder.upsample.0", 6400, n, "fc_out", "fc_out_bias")
    for pre, ci, co, _, _, _, b in DEC_CONVS:
        out.append((pre + ".weight", (ci, co, 4, 4), "convT"))
        if b:
            bn(b, co)
    lin("attrs_encoder.net.0", 64, N_ATTRS)
    bn("attrs_encoder.net.1", 64)
    lin("attrs_encoder.net.3", 2 * n, 64)
    lin("attrs_decoder.net.0", 64, n)
    bn("attrs_decoder.net.1", 64)
    lin("attrs_decoder.net.3", N_ATTRS, 64)
    return out


class MultimodalVAE:
    """Drop-in for celeba/model.py:13-58.  `precision`: "bf16" (default) or "tf32" (fp32 storage; the parity path)."""

    def __init__(self, n_latents: int = 20, use_cuda: bool = True, precision: str = "bf16", dropout_p: float = 0.1,
                 device: Optional[torch.device] = None, seed: int = 0):
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
        # ---- flat parameter buffer, [encoders | decoders] (the two all-reduce buckets of data-parallel training)
        self.layouts: Dict[str, _Layout] = {}
        self.state_keys = reference_keys(self.n_latents)
        params = [_Layout(k, s, kind) for k, s, kind in self.state_keys if kind not in ("rm", "rv", "nbt")]
        enc = [l for l in params if "encoder" in l.key]
        dec = [l for l in params if "decoder" in l.key]
        off = 0
        for l in enc + dec:
            l.offset = off
            off += _round_up(l.numel, 64)
            self.layouts[l.key] = l
            if l is enc[-1]:
                self.encoder_param_floats = off
        self.param_floats = off
        self.flat_params = torch.zeros(off, device=dev, dtype=torch.float32)
        self.flat_grads = torch.zeros(off, device=dev, dtype=torch.float32)
        self.flat_params_bf16 = torch.zeros(off, device=dev, dtype=torch.bfloat16) if self.act_dtype == torch.bfloat16 else None
        # ---- BatchNorm buffers
        self.bn_index = {p: i for i, p in enumerate(BN_LAYERS)}
        self.bn_off: Dict[str, int] = {}
        boff = 0
        for p in BN_LAYERS:
            self.bn_off[p] = boff
            boff += 2 * _BN_CH[p]
        self.flat_buffers = torch.zeros(boff, device=dev, dtype=torch.float32)
        self.flat_nbt = torch.zeros(len(BN_LAYERS), device=dev, dtype=torch.int64)
        self._step_counter = torch.zeros(1, device=dev, dtype=torch.int32)
        self._ws: Dict[Tuple[int, int], "_Workspace"] = {}
        self._pad: Dict[str, Tuple[torch.Tensor, int]] = {}
        self.reset_parameters()

    # ------------------------------------------------------------------ parameters
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
        o, c = self.bn_off[bn], _BN_CH[bn]
        return self.flat_buffers[o:o + c], self.flat_buffers[o + c:o + 2 * c]

    def reset_parameters(self, seed: int = 1234) -> None:
        """PyTorch-default-like initialisation (the reference's weight_init is a no-op, celeba/model.py:121-123)."""
        g = torch.Generator().manual_seed(seed)
        sd = {}
        for k, shp, kind in self.state_keys:
            if kind == "nbt":
                sd[k] = torch.zeros((), dtype=torch.int64)
            elif kind == "rm":
                sd[k] = torch.zeros(shp)
            elif kind == "rv":
                sd[k] = torch.ones(shp)
            elif k.rsplit(".", 1)[0] in BN_LAYERS:
                sd[k] = torch.ones(shp) if k.endswith("weight") else torch.zeros(shp)
            elif len(shp) == 4:
                fan_in = (shp[1] if kind == "conv" else shp[0]) * 16
                sd[k] = (torch.rand(shp, generator=g) * 2 - 1) / fan_in ** 0.5
            elif len(shp) == 2:
                sd[k] = (torch.rand(shp, generator=g) * 2 - 1) / shp[1] ** 0.5
            else:
                fan_in = sd[k[:-4] + "weight"].shape[1]
                sd[k] = (torch.rand(shp, generator=g) * 2 - 1) / fan_in ** 0.5
        self.load_state_dict(sd)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """Reference-shaped copies under the reference's keys (celeba/train.py:215)."""
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
        """Loads a reference checkpoint's state_dict (celeba/train.py:48-56)."""
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
        return [self.flat_params]

    # ------------------------------------------------------------------ operand copies with TMA-legal strides
    def _padded(self, key: str, rows: int, cols: int) -> Tuple[torch.Tensor, int]:
        """Weights whose row length is not a 16-byte multiple get a zero-padded operand copy (refreshed per forward)."""
        ld = _round_up(cols, self.vec)
        if ld == cols:
            return self.W(key), cols
        if key not in self._pad:
            self._pad[key] = (torch.zeros(rows * ld, device=self.device, dtype=self.act_dtype), ld)
        buf, _ = self._pad[key]
        _ops.cast_pad_2d(self.P(key), rows, cols, cols, buf, ld)
        return buf, ld

    # ------------------------------------------------------------------ forward / backward
    def workspace(self, batch: int, n_terms: int) -> "_Workspace":
        key = (int(batch), int(n_terms))
        if key not in self._ws:
            self._ws[key] = _Workspace(self, batch, n_terms)
        return self._ws[key]

    def _run_forward(self, ws: "_Workspace", image, attrs, term_types: Sequence[int], eps, training: bool,
                     lambdas, kl_weights, want_probs: bool, with_loss: bool) -> None:
        B, G, n, T = ws.B, ws.G, self.n_latents, self.act_dtype
        dev = self.device
        use_img = any(t != _lib.TERM_TEXT for t in term_types)
        use_att = any(t != _lib.TERM_IMAGE for t in term_types)
        n_img_terms = sum(1 for t in term_types if t != _lib.TERM_TEXT)
        n_att_terms = sum(1 for t in term_types if t != _lib.TERM_IMAGE)
        R = n_img_terms if (training and self.dropout_p > 0 and n_img_terms > 1) else 1
        ws.R, ws.term_types, ws.training = R, tuple(term_types), training
        ws.use_img, ws.use_att = use_img, use_att
        ws.image, ws.attrs = image, attrs
        # ---------------- image encoder (once)
        if use_img:
            src = image
            for li, (pre, ci, co, s, p, hin, bn) in enumerate(ENC_CONVS):
                ho = _ops.out_size(hin, 4, s, p)
                K = 16 * ci
                strides = _ops.nchw_strides(3, 64, 64) if li == 0 else None
                g = _ops.geometry(B, hin, hin, ci, 4, s, p, strides)
                _ops.im2col(g, src, ws.enc_col[li], K)
                rows = B * ho * ho
                _ops.gemm(ws.enc_col[li], self.W(pre + ".weight"), ws.enc_pre[li], rows, co, K, K, K, co)
                if bn:
                    rm, rv = self.running(bn)
                    a = _ops.bn_args(ws.enc_pre[li], rows, co, rows, SWISH, training, self.P(bn + ".weight"), self.P(bn + ".bias"),
                                     ws.enc_sum[li], ws.enc_sumsq[li], ws.enc_mean[li], ws.enc_rstd[li], rm, rv,
                                     updates=n_img_terms)
                    _ops.bn_act_forward(a, ws.enc_act[li])
                else:
                    _ops.act_forward(SWISH, ws.enc_pre[li], ws.enc_act[li], rows, co)
                src = ws.enc_act[li]
            # classifier: Linear(6400, 1024) + Swish + Dropout + Linear(1024, 2n)   celeba/model.py:114-119
            _ops.gemm(ws.enc_act[3], self.W("image_encoder.classifier.0.weight"), ws.f1pre, B, 1024, 6400, 6400, 6400, 1024,
                      bias=self.P("image_encoder.classifier.0.bias"))
            _ops.act_forward(SWISH, ws.f1pre, ws.f1, B, 1024, repeat=R, dropout_p=self.dropout_p if (training and R >= 1) else 0.0,
                             seed=self.noise_seed, step_counter=self._step_counter)
            _ops.gemm(ws.f1, self.W("image_encoder.classifier.3.weight"), ws.encA, R * B, 2 * n, 1024, 1024, 1024, 2 * n,
                      bias=self.P("image_encoder.classifier.3.bias"))
        # ---------------- attribute encoder (once)   celeba/model.py:170-176
        if use_att:
            _ops.cast_pad_2d(attrs, B, N_ATTRS, N_ATTRS, ws.attrs_pad, ws.ld_attr)
            w0, ldw0 = self._padded("attrs_encoder.net.0.weight", 64, N_ATTRS)
            bn = "attrs_encoder.net.1"
            _ops.gemm(ws.attrs_pad, w0, ws.t1pre, B, 64, N_ATTRS, ws.ld_attr, ldw0, 64, bias=self.P("attrs_encoder.net.0.bias"))
            rm, rv = self.running(bn)
            ws.ae_bn = _ops.bn_args(ws.t1pre, B, 64, B, SWISH, training, self.P(bn + ".weight"), self.P(bn + ".bias"),
                                    ws.ae_sum[0], ws.ae_sum[1], ws.ae_mean, ws.ae_rstd, rm, rv, updates=n_att_terms)
            _ops.bn_act_forward(ws.ae_bn, ws.t1)
            _ops.gemm(ws.t1, self.W("attrs_encoder.net.3.weight"), ws.encB, B, 2 * n, 64, 64, 64, 2 * n,
                      bias=self.P("attrs_encoder.net.3.bias"))
        # ---------------- latent path: PoE, reparametrize, KL for all terms
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
        if use_img:
            la.expert_a, la.ld_a = ws.encA.data_ptr(), 2 * n
        if use_att:
            la.expert_b, la.ld_b = ws.encB.data_ptr(), 2 * n
        la.eps = None if eps is None else eps.data_ptr()
        la.seed, la.step_counter = self.noise_seed, self._step_counter.data_ptr()
        la.training = 1 if training else 0
        la.z_dtype, la.z, la.ld_z = _ops.DT[T], ws.z.data_ptr(), ws.ld_z
        la.mu, la.logvar, la.kl = ws.mu.data_ptr(), ws.logvar.data_ptr(), ws.acc[2].data_ptr()
        ws.latent = la
        _lib.check(_lib.load().mvae_latent_forward(C.byref(la), _ops.stream()), "mvae_latent_forward")
        self._decode(ws, training, lambdas, want_probs, with_loss)

    def _decode(self, ws: "_Workspace", training: bool, lambdas, want_probs: bool, with_loss: bool) -> None:
        """Image and attribute decoders on the stacked [G*B] latents (celeba/model.py:134-163, 185-200) + BCE terms."""
        B, G, n = ws.B, ws.G, self.n_latents
        M3 = G * B
        wup, ldup = self._padded("image_decoder.upsample.0.weight", 6400, n)
        ws.wup, ws.ldup = wup, ldup
        _ops.gemm(ws.z, wup, ws.u1pre, M3, 6400, n, ws.ld_z, ldup, 6400, bias=self.P("image_decoder.upsample.0.bias"))
        _ops.act_forward(SWISH, ws.u1pre, ws.u1, M3, 6400)
        src = ws.u1
        for li, (pre, ci, co, s, p, hout, bn) in enumerate(DEC_CONVS):
            hin = _ops.out_size(hout, 4, s, p)
            K = 16 * co
            rows_in = M3 * hin * hin
            # col[M_in, (kh,kw,co)] = X[M_in, ci] * W'[ci, (kh,kw,co)]
            _ops.gemm(src, self.W(pre + ".weight"), ws.colbuf, rows_in, K, ci, ci, K, K, b_major=1)
            if bn:
                g = _ops.geometry(M3, hout, hout, co, 4, s, p)
                _ops.col2im(g, ws.colbuf, K, ws.dec_pre[li])
                rows = M3 * hout * hout
                rm, rv = self.running(bn)
                a = _ops.bn_args(ws.dec_pre[li], rows, co, B * hout * hout, SWISH, training, self.P(bn + ".weight"),
                                 self.P(bn + ".bias"), ws.dec_sum[li], ws.dec_sumsq[li], ws.dec_mean[li], ws.dec_rstd[li],
                                 rm, rv, updates=1)
                ws.dec_bn[li] = a
                _ops.bn_act_forward(a, ws.dec_act[li])
                src = ws.dec_act[li]
            else:
                g = _ops.geometry(M3, 64, 64, 3, 4, s, p, _ops.nchw_strides(3, 64, 64))
                _ops.col2im(g, ws.colbuf, K, ws.logits)
        sx = [float(lambdas[g][0]) / (B * 12288) for g in range(G)]
        sy = [float(lambdas[g][1]) / (B * N_ATTRS) for g in range(G)]
        _ops.sigmoid_bce(ws.logits, 12288, M3, 12288, rows_per_group=B,
                         target=ws.image if with_loss else None, ld_target=12288, target_rows=B, grad_scale=sx,
                         loss=ws.acc[0] if with_loss else None, probs=ws.probs_image if want_probs else None, ld_probs=12288,
                         dlogits=ws.logits if with_loss else None, ld_dlogits=12288)
        # attribute decoder
        wad, ldad = self._padded("attrs_decoder.net.0.weight", 64, n)
        ws.wad, ws.ldad = wad, ldad
        bn = "attrs_decoder.net.1"
        _ops.gemm(ws.z, wad, ws.s1pre, M3, 64, n, ws.ld_z, ldad, 64, bias=self.P("attrs_decoder.net.0.bias"))
        rm, rv = self.running(bn)
        ws.ad_bn = _ops.bn_args(ws.s1pre, M3, 64, B, SWISH, training, self.P(bn + ".weight"), self.P(bn + ".bias"),
                                ws.ad_sum[0], ws.ad_sum[1], ws.ad_mean, ws.ad_rstd, rm, rv, updates=1)
        _ops.bn_act_forward(ws.ad_bn, ws.s1)
        _ops.gemm(ws.s1, self.W("attrs_decoder.net.3.weight"), ws.alogits, M3, N_ATTRS, 64, 64, 64, N_ATTRS,
                  bias=self.P("attrs_decoder.net.3.bias"))
        _ops.sigmoid_bce(ws.alogits, N_ATTRS, M3, N_ATTRS, rows_per_group=B,
                         target=ws.attrs if with_loss else None, ld_target=N_ATTRS, target_rows=B, grad_scale=sy,
                         loss=ws.acc[1] if with_loss else None, probs=ws.probs_attrs if want_probs else None, ld_probs=N_ATTRS,
                         dlogits=ws.dalog if with_loss else None, ld_dlogits=ws.ld_dalog)

    def _run_backward(self, ws: "_Workspace") -> None:
        """Backward of the summed ELBO terms into flat_grads (celeba/train.py:151-152)."""
        B, G, n, R = ws.B, ws.G, self.n_latents, ws.R
        M3 = G * B
        Gd = self.G
        # ---------------- attribute decoder
        _ops.gemm(ws.dalog, ws.s1, Gd("attrs_decoder.net.3.weight"), N_ATTRS, 64, M3, ws.ld_dalog, 64, 64, a_major=1, b_major=1,
                  accumulate=True)
        _ops.col_stats(ws.dalog, M3, ws.ld_dalog, Gd("attrs_decoder.net.3.bias"), valid_channels=N_ATTRS)
        _ops.gemm(ws.dalog, self.W("attrs_decoder.net.3.weight"), ws.ds1, M3, 64, N_ATTRS, ws.ld_dalog, 64, 64, b_major=1)
        bn = "attrs_decoder.net.1"
        _ops.bn_act_backward(ws.ad_bn, ws.ds1, ws.ds1pre, ws.ad_s[0], ws.ad_s[1], Gd(bn + ".weight"), Gd(bn + ".bias"))
        _ops.gemm(ws.ds1pre, ws.z, Gd("attrs_decoder.net.0.weight"), 64, n, M3, 64, ws.ld_z, n, a_major=1, b_major=1, accumulate=True)
        _ops.gemm(ws.ds1pre, ws.wad, ws.dz, M3, n, 64, 64, ws.ldad, n, b_major=1)
        # ---------------- image decoder (gradient of the BCE sits in ws.logits, NCHW fp32)
        dsrc = ws.logits
        for li in (3, 2, 1, 0):
            pre, ci, co, s, p, hout, bn = DEC_CONVS[li]
            hin = _ops.out_size(hout, 4, s, p)
            K = 16 * co
            rows_in = M3 * hin * hin
            strides = _ops.nchw_strides(3, 64, 64) if li == 3 else None
            g = _ops.geometry(M3, hout, hout, co, 4, s, p, strides)
            if bn:
                _ops.bn_act_backward(ws.dec_bn[li], dsrc, ws.dec_dpre[li], ws.dec_s0[li], ws.dec_s1[li], Gd(bn + ".weight"),
                                     Gd(bn + ".bias"))
                dsrc = ws.dec_dpre[li]
            _ops.im2col(g, dsrc, ws.colbuf, K)                        # dcol [M_in, (kh,kw,co)]
            x_in = ws.dec_act[li - 1] if li > 0 else ws.u1            # the layer's input [M_in, ci]
            _ops.gemm(x_in, ws.colbuf, Gd(pre + ".weight"), ci, K, rows_in, ci, K, K, a_major=1, b_major=1, accumulate=True)
            dx = ws.dec_dact[li - 1] if li > 0 else ws.du1
            _ops.gemm(ws.colbuf, self.W(pre + ".weight"), dx, rows_in, ci, K, K, K, ci)
            dsrc = dx
        _ops.act_backward(SWISH, ws.u1pre, ws.du1, ws.du1pre, M3, 6400, dbias=Gd("image_decoder.upsample.0.bias"))
        _ops.gemm(ws.du1pre, ws.z, Gd("image_decoder.upsample.0.weight"), 6400, n, M3, 6400, ws.ld_z, n, a_major=1, b_major=1,
                  accumulate=True)
        _ops.gemm(ws.du1pre, ws.wup, ws.dz, M3, n, 6400, 6400, ws.ldup, n, b_major=1, accumulate=True)
        # ---------------- latent path
        la = ws.latent
        la.dz_dtype, la.dz, la.ld_dz = _lib.DT_F32, ws.dz.data_ptr(), n
        la.d_dtype = _ops.DT[self.act_dtype]
        la.d_expert_a, la.ld_da = (ws.dencA.data_ptr() if ws.use_img else None), ws.ld_enc
        la.d_expert_b, la.ld_db = (ws.dencB.data_ptr() if ws.use_att else None), ws.ld_enc
        _lib.check(_lib.load().mvae_latent_backward(C.byref(la), _ops.stream()), "mvae_latent_backward")
        # ---------------- attribute encoder
        if ws.use_att:
            _ops.gemm(ws.dencB, ws.t1, Gd("attrs_encoder.net.3.weight"), 2 * n, 64, B, ws.ld_enc, 64, 64, a_major=1, b_major=1,
                      accumulate=True)
            _ops.col_stats(ws.dencB, B, ws.ld_enc, Gd("attrs_encoder.net.3.bias"), valid_channels=2 * n)
            _ops.gemm(ws.dencB, self.W("attrs_encoder.net.3.weight"), ws.dt1, B, 64, 2 * n, ws.ld_enc, 64, 64, b_major=1)
            bn = "attrs_encoder.net.1"
            _ops.bn_act_backward(ws.ae_bn, ws.dt1, ws.dt1pre, ws.ae_s[0], ws.ae_s[1], Gd(bn + ".weight"), Gd(bn + ".bias"))
            _ops.gemm(ws.dt1pre, ws.attrs_pad, Gd("attrs_encoder.net.0.weight"), 64, N_ATTRS, B, 64, ws.ld_attr, N_ATTRS,
                      a_major=1, b_major=1, accumulate=True)
        # ---------------- image encoder
        if ws.use_img:
            RB = R * B
            _ops.gemm(ws.dencA, ws.f1, Gd("image_encoder.classifier.3.weight"), 2 * n, 1024, RB, ws.ld_enc, 1024, 1024,
                      a_major=1, b_major=1, accumulate=True)
            _ops.col_stats(ws.dencA, RB, ws.ld_enc, Gd("image_encoder.classifier.3.bias"), valid_channels=2 * n)
            _ops.gemm(ws.dencA, self.W("image_encoder.classifier.3.weight"), ws.df1, RB, 1024, 2 * n, ws.ld_enc, 1024, 1024, b_major=1)
            _ops.act_backward(SWISH, ws.f1pre, ws.df1, ws.df1pre, B, 1024, repeat=R,
                              dropout_p=self.dropout_p if ws.training else 0.0, seed=self.noise_seed,
                              step_counter=self._step_counter, dbias=Gd("image_encoder.classifier.0.bias"))
            _ops.gemm(ws.df1pre, ws.enc_act[3], Gd("image_encoder.classifier.0.weight"), 1024, 6400, B, 1024, 6400, 6400,
                      a_major=1, b_major=1, accumulate=True)
            _ops.gemm(ws.df1pre, self.W("image_encoder.classifier.0.weight"), ws.enc_dact[3], B, 6400, 1024, 1024, 6400, 6400,
                      b_major=1)
            for li in (3, 2, 1, 0):
                pre, ci, co, s, p, hin, bn = ENC_CONVS[li]
                ho = _ops.out_size(hin, 4, s, p)
                K = 16 * ci
                rows = B * ho * ho
                if bn:
                    a = _ops.bn_args(ws.enc_pre[li], rows, co, rows, SWISH, True, self.P(bn + ".weight"), self.P(bn + ".bias"),
                                     None, None, ws.enc_mean[li], ws.enc_rstd[li])
                    _ops.bn_act_backward(a, ws.enc_dact[li], ws.enc_dpre[li], ws.enc_s0[li], ws.enc_s1[li], Gd(bn + ".weight"),
                                         Gd(bn + ".bias"))
                else:
                    _ops.act_backward(SWISH, ws.enc_pre[li], ws.enc_dact[li], ws.enc_dpre[li], rows, co)
                _ops.gemm(ws.enc_dpre[li], ws.enc_col[li], Gd(pre + ".weight"), co, K, rows, co, K, K, a_major=1, b_major=1,
                          accumulate=True)
                if li > 0:
                    _ops.gemm(ws.enc_dpre[li], self.W(pre + ".weight"), ws.colbuf, rows, K, co, co, K, K, b_major=1)
                    g = _ops.geometry(B, hin, hin, ci, 4, s, p)
                    _ops.col2im(g, ws.colbuf, K, ws.enc_dact[li - 1])

    # ------------------------------------------------------------------ module surface
    def forward(self, image: Optional[torch.Tensor] = None, attrs: Optional[torch.Tensor] = None, eps: Optional[torch.Tensor] = None):
        """celeba/model.py:36-58: returns (image_recon, attrs_recon, mu, logvar).  Forward only (use CelebATrainer.step
        for training); train mode draws the reparametrize noise in-kernel unless `eps` [B, n] is given."""
        assert image is not None or attrs is not None
        t = _lib.TERM_JOINT if (image is not None and attrs is not None) else (_lib.TERM_IMAGE if image is not None else _lib.TERM_TEXT)
        B = (image if image is not None else attrs).shape[0]
        ws = self.workspace(B, 1)
        image = None if image is None else image.to(self.device, torch.float32).contiguous()
        attrs = None if attrs is None else attrs.to(self.device, torch.float32).contiguous()
        if eps is not None:
            eps = eps.to(self.device, torch.float32).contiguous()
        ws.acc.zero_()
        self._run_forward(ws, image, attrs, (t,), eps, self.training, ((0.0, 0.0),), (0.0,), True, False)
        n = self.n_latents
        return (ws.probs_image.view(B, 3, 64, 64).clone(), ws.probs_attrs.view(B, N_ATTRS).clone(),
                ws.mu.view(1, B, n)[0].clone(), ws.logvar.view(1, B, n)[0].clone())

    __call__ = forward

    def image_decoder(self, z: torch.Tensor) -> torch.Tensor:
        return self._decode_only(z)[0]

    def attrs_decoder(self, z: torch.Tensor) -> torch.Tensor:
        return self._decode_only(z)[1]

    def _decode_only(self, z: torch.Tensor):
        B = z.shape[0]
        ws = self.workspace(B, 1)
        ws.z.view(B, ws.ld_z)[:, :self.n_latents].copy_(z.to(self.device))
        self._decode(ws, self.training, ((0.0, 0.0),), True, False)
        return ws.probs_image.view(B, 3, 64, 64).clone(), ws.probs_attrs.view(B, N_ATTRS).clone()


class _Workspace:
    """Activation / gradient buffers of one (batch, n_terms) configuration (allocated once, reused every step)."""

    def __init__(self, m: MultimodalVAE, B: int, G: int):
        dev, T, n = m.device, m.act_dtype, m.n_latents
        self.B, self.G = B, G
        self.R = 1
        M3 = G * B
        Rmax = 2

        def buf(*shape, dtype=T):
            return torch.zeros(*shape, device=dev, dtype=dtype)

        f32 = torch.float32
        # encoder
        self.enc_col, self.enc_pre, self.enc_act, self.enc_dact, self.enc_dpre = [], [], [], [], []
        self.enc_sum, self.enc_sumsq, self.enc_mean, self.enc_rstd, self.enc_s0, self.enc_s1 = [], [], [], [], [], []
        for pre, ci, co, s, p, hin, bn in ENC_CONVS:
            ho = _ops.out_size(hin, 4, s, p)
            rows = B * ho * ho
            self.enc_col.append(buf(rows * 16 * ci))
            self.enc_pre.append(buf(rows * co))
            self.enc_act.append(buf(rows * co))
            self.enc_dact.append(buf(rows * co))
            self.enc_dpre.append(buf(rows * co))
            for lst in (self.enc_sum, self.enc_sumsq, self.enc_mean, self.enc_rstd, self.enc_s0, self.enc_s1):
                lst.append(buf(co, dtype=f32))
        self.f1pre = buf(B * 1024)
        self.f1 = buf(Rmax * B * 1024)
        self.df1 = buf(Rmax * B * 1024)
        self.df1pre = buf(B * 1024)
        self.ld_enc = _round_up(2 * n, m.vec)
        self.encA = buf(Rmax * B * 2 * n, dtype=f32)
        self.encB = buf(B * 2 * n, dtype=f32)
        self.dencA = buf(Rmax * B * self.ld_enc)
        self.dencB = buf(B * self.ld_enc)
        # attribute encoder
        self.ld_attr = _round_up(N_ATTRS, m.vec)
        self.attrs_pad = buf(B * self.ld_attr)
        self.t1pre, self.t1, self.dt1, self.dt1pre = buf(B * 64), buf(B * 64), buf(B * 64), buf(B * 64)
        self.ae_sum = buf(2, 64, dtype=f32)
        self.ae_mean, self.ae_rstd = buf(64, dtype=f32), buf(64, dtype=f32)
        self.ae_s = buf(2, 64, dtype=f32)
        # latent
        self.ld_z = _round_up(n, m.vec)
        self.z = buf(M3 * self.ld_z)
        self.dz = buf(M3 * n, dtype=f32)
        self.mu, self.logvar = buf(M3 * n, dtype=f32), buf(M3 * n, dtype=f32)
        self.acc = buf(3, 4, dtype=f32)   # rows: image BCE sums, attrs BCE sums, weighted KL; columns: term
        # image decoder
        self.u1pre, self.u1, self.du1, self.du1pre = buf(M3 * 6400), buf(M3 * 6400), buf(M3 * 6400), buf(M3 * 6400)
        self.dec_pre, self.dec_act, self.dec_dact, self.dec_dpre = [], [], [], []
        self.dec_sum, self.dec_sumsq, self.dec_mean, self.dec_rstd, self.dec_s0, self.dec_s1 = [], [], [], [], [], []
        self.dec_bn: List[Optional[_lib.BnActArgs]] = [None] * 4
        colmax = 0
        for pre, ci, co, s, p, hout, bn in DEC_CONVS:
            hin = _ops.out_size(hout, 4, s, p)
            colmax = max(colmax, M3 * hin * hin * 16 * co)
            rows = M3 * hout * hout
            if bn:
                self.dec_pre.append(buf(rows * co))
                self.dec_act.append(buf(rows * co))
                self.dec_dact.append(buf(rows * co))
                self.dec_dpre.append(buf(rows * co))
            for lst in (self.dec_sum, self.dec_sumsq, self.dec_mean, self.dec_rstd, self.dec_s0, self.dec_s1):
                lst.append(buf(G, co, dtype=f32))
        for pre, ci, co, s, p, hin, bn in ENC_CONVS[1:]:
            ho = _ops.out_size(hin, 4, s, p)
            colmax = max(colmax, B * ho * ho * 16 * ci)
        self.colbuf = buf(colmax)
        self.logits = buf(M3 * 12288, dtype=f32)
        self.probs_image = buf(M3 * 12288, dtype=f32)
        # attribute decoder
        self.s1pre, self.s1, self.ds1, self.ds1pre = buf(M3 * 64), buf(M3 * 64), buf(M3 * 64), buf(M3 * 64)
        self.ad_sum = buf(2, G, 64, dtype=f32)
        self.ad_mean, self.ad_rstd = buf(G, 64, dtype=f32), buf(G, 64, dtype=f32)
        self.ad_s = buf(2, G, 64, dtype=f32)
        self.alogits = buf(M3 * N_ATTRS, dtype=f32)
        self.probs_attrs = buf(M3 * N_ATTRS, dtype=f32)
        self.ld_dalog = _round_up(N_ATTRS, m.vec)
        self.dalog = buf(M3 * self.ld_dalog)


class CelebATrainer:
    """celeba/train.py:132-157: zero_grad, three forwards, three loss_function calls, backward, Adam."""

    def __init__(self, model: MultimodalVAE, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, kl_lambda: float = 1e-3,
                 use_cuda_graph: bool = False):
        self.model = model
        self.lr, self.betas, self.eps, self.kl_lambda = float(lr), betas, float(eps), float(kl_lambda)
        self.adam_m = torch.zeros_like(model.flat_params)
        self.adam_v = torch.zeros_like(model.flat_params)
        self.use_cuda_graph = use_cuda_graph
        self._graphs: Dict[Tuple, Tuple] = {}
        self._inc_cache: Dict[Tuple, torch.Tensor] = {}

    def _increments(self, term_types) -> torch.Tensor:
        key = tuple(term_types)
        if key not in self._inc_cache:
            ni = sum(1 for t in term_types if t != _lib.TERM_TEXT)
            na = sum(1 for t in term_types if t != _lib.TERM_IMAGE)
            g = len(term_types)
            inc = [ni, ni, ni, g, g, g, na, g]
            self._inc_cache[key] = torch.tensor(inc, dtype=torch.int64, device=self.model.device)
        return self._inc_cache[key]

    def _enqueue(self, ws, image, attrs, term_types, lambdas, eps, adam: bool, grad_scale: float = 1.0) -> None:
        m = self.model
        B = ws.B
        _ops.step_begin(m._step_counter, ws.acc.view(-1), m.flat_nbt, self._increments(term_types))
        klw = [self.kl_lambda / B] * len(term_types)
        m._run_forward(ws, image, attrs, term_types, eps, True, lambdas, klw, False, True)
        m._run_backward(ws)
        if adam:
            _ops.adam_step(m.flat_params, m.flat_grads, self.adam_m, self.adam_v, m.flat_params_bf16, m.param_floats, self.lr,
                           self.betas[0], self.betas[1], self.eps, m._step_counter, grad_scale, True)

    def step(self, image: torch.Tensor, attrs: torch.Tensor, terms: Sequence[str] = ("joint", "image", "attrs"),
             lambdas: Sequence[Tuple[float, float]] = ((1.0, 1.0),) * 3, eps: Optional[torch.Tensor] = None, adam: bool = True):
        """One training step.  Returns the device tensor of raw accumulators; `losses()` turns it into the per-term
        (total, image BCE, attrs BCE, KL) values of celeba/train.py:60-81.  eps: optional [n_terms, B, n] injected noise."""
        m = self.model
        tt = tuple(TERMS[t] for t in terms)
        B = image.shape[0]
        ws = m.workspace(B, len(tt))
        image = image.to(m.device, torch.float32).contiguous()
        attrs = attrs.to(m.device, torch.float32).contiguous()
        if eps is not None:
            eps = eps.to(m.device, torch.float32).contiguous()
        self._last = (ws, tt, tuple(lambdas))
        if not self.use_cuda_graph:
            self._enqueue(ws, image, attrs, tt, lambdas, eps, adam)
            return ws.acc
        key = (B, tt, tuple(lambdas), eps is not None, adam)
        if key not in self._graphs:
            st_img, st_att = image.clone(), attrs.clone()
            st_eps = None if eps is None else eps.clone()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                # warm-up outside capture (allocates the padded operand copies), state restored afterwards
                snap = (m.flat_params.clone(), m.flat_buffers.clone(), m.flat_nbt.clone(), m._step_counter.clone(),
                        self.adam_m.clone(), self.adam_v.clone(), m.flat_grads.clone())
                self._enqueue(ws, st_img, st_att, tt, lambdas, st_eps, adam)
                for dst, src in zip((m.flat_params, m.flat_buffers, m.flat_nbt, m._step_counter, self.adam_m, self.adam_v,
                                     m.flat_grads), snap):
                    dst.copy_(src)
                m.sync_operands()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue(ws, st_img, st_att, tt, lambdas, st_eps, adam)
            self._graphs[key] = (g, st_img, st_att, st_eps)
        g, st_img, st_att, st_eps = self._graphs[key]
        st_img.copy_(image, non_blocking=True)
        st_att.copy_(attrs, non_blocking=True)
        if st_eps is not None:
            st_eps.copy_(eps, non_blocking=True)
        g.replay()
        return ws.acc

    def losses(self) -> List[Tuple[float, float, float, float]]:
        """Per-term (total, image BCE term, attrs BCE term, KL term) of the last step (one small D2H copy)."""
        ws, tt, lambdas = self._last
        acc = ws.acc.cpu()
        B = ws.B
        out = []
        for g in range(len(tt)):
            x = float(acc[0, g]) * lambdas[g][0] / (B * 12288)
            y = float(acc[1, g]) * lambdas[g][1] / (B * N_ATTRS)
            k = float(acc[2, g])
            out.append((x + y + k, x, y, k))
        return out
