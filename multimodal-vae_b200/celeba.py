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
from .convnet import ConvMVAEBase, ConvMVAETrainer, Workspace, SWISH, round_up

N_ATTRS = 18  # celeba/datasets.py:27


class MultimodalVAE(ConvMVAEBase):
    """Drop-in for celeba/model.py:13-58.  `precision`: "bf16" (default) or "tf32" (fp32 storage; the parity path)."""

    TERMS = {"joint": _lib.TERM_JOINT, "image": _lib.TERM_IMAGE, "attrs": _lib.TERM_TEXT}
    IMG_C, IMG_H = 3, 64
    # (key, Cin, Cout, k, stride, pad, input H, BatchNorm key)  celeba/model.py:101-113
    ENC_CONVS = (("image_encoder.features.0", 3, 32, 4, 2, 1, 64, None),
                 ("image_encoder.features.2", 32, 64, 4, 2, 1, 32, "image_encoder.features.3"),
                 ("image_encoder.features.5", 64, 128, 4, 2, 1, 16, "image_encoder.features.6"),
                 ("image_encoder.features.8", 128, 256, 4, 1, 0, 8, "image_encoder.features.9"))
    # (key, Cin, Cout, k, stride, pad, OUTPUT H, BatchNorm key)  celeba/model.py:142-152
    DEC_CONVS = (("image_decoder.hallucinate.0", 256, 128, 4, 1, 0, 8, "image_decoder.hallucinate.1"),
                 ("image_decoder.hallucinate.3", 128, 64, 4, 2, 1, 16, "image_decoder.hallucinate.4"),
                 ("image_decoder.hallucinate.6", 64, 32, 4, 2, 1, 32, "image_decoder.hallucinate.7"),
                 ("image_decoder.hallucinate.9", 32, 3, 4, 2, 1, 64, None))
    FLAT_C, FLAT_HW = 256, 25
    BN_LAYERS = {"image_encoder.features.3": 64, "image_encoder.features.6": 128, "image_encoder.features.9": 256,
                 "image_decoder.hallucinate.1": 128, "image_decoder.hallucinate.4": 64, "image_decoder.hallucinate.7": 32,
                 "attrs_encoder.net.1": 64, "attrs_decoder.net.1": 64}

    def __init__(self, n_latents: int = 20, use_cuda: bool = True, precision: str = "bf16", dropout_p: float = 0.1,
                 device: Optional[torch.device] = None, seed: int = 0):
        super().__init__(n_latents, precision, dropout_p, device, seed)

    def reference_keys(self, n_latents: int) -> List[Tuple[str, Tuple[int, ...], str]]:
        """(key, reference shape, kind) in the reference's state_dict order (celeba/model.py:16-21)."""
        n = n_latents
        out: List[Tuple[str, Tuple[int, ...], str]] = []

        def bn(p, c):
            out.extend([(p + ".weight", (c,), "plain"), (p + ".bias", (c,), "plain"), (p + ".running_mean", (c,), "rm"),
                        (p + ".running_var", (c,), "rv"), (p + ".num_batches_tracked", (), "nbt")])

        def lin(p, o, i, wkind="plain", bkind="plain"):
            out.extend([(p + ".weight", (o, i), wkind), (p + ".bias", (o,), bkind)])

        for pre, ci, co, k, _, _, _, b in self.ENC_CONVS:
            out.append((pre + ".weight", (co, ci, k, k), "conv"))
            if b:
                bn(b, co)
        lin("image_encoder.classifier.0", 1024, 6400, "fc_in")
        lin("image_encoder.classifier.3", 2 * n, 1024)
        lin("image_decoder.upsample.0", 6400, n, "fc_out", "fc_out_bias")
        for pre, ci, co, k, _, _, _, b in self.DEC_CONVS:
            out.append((pre + ".weight", (ci, co, k, k), "convT"))
            if b:
                bn(b, co)
        lin("attrs_encoder.net.0", 64, N_ATTRS)
        bn("attrs_encoder.net.1", 64)
        lin("attrs_encoder.net.3", 2 * n, 64)
        lin("attrs_decoder.net.0", 64, n)
        bn("attrs_decoder.net.1", 64)
        lin("attrs_decoder.net.3", N_ATTRS, 64)
        return out

    def linear_shapes(self, n_terms: int, n_img_terms: int):
        """(out, in, passes) of every Linear, for the algorithmic FLOP count of bench_conv.py."""
        n = self.n_latents
        return [(1024, 6400, 1), (2 * n, 1024, n_img_terms), (6400, n, n_terms), (64, N_ATTRS, 1), (2 * n, 64, 1),
                (64, n, n_terms), (N_ATTRS, 64, n_terms)]

    def bn_increments(self, term_types) -> List[int]:
        ni = sum(1 for t in term_types if t != _lib.TERM_TEXT)
        na = sum(1 for t in term_types if t != _lib.TERM_IMAGE)
        g = len(term_types)
        return [ni, ni, ni, g, g, g, na, g]

    # ------------------------------------------------------------------ workspace
    def _make_workspace(self, B: int, G: int) -> Workspace:
        ws = Workspace()
        ws.B, ws.G, ws.R = B, G, 1
        self.alloc_conv_buffers(ws, B, G)
        buf, n, f32 = ws.buf, self.n_latents, torch.float32
        M3, Rmax = G * B, 2
        ws.f1pre, ws.f1, ws.df1, ws.df1pre = buf(B * 1024), buf(Rmax * B * 1024), buf(Rmax * B * 1024), buf(B * 1024)
        ws.encA, ws.encB = buf(Rmax * B * 2 * n, dtype=f32), buf(B * 2 * n, dtype=f32)
        ws.dencA, ws.dencB = buf(Rmax * B * ws.ld_enc), buf(B * ws.ld_enc)
        ws.ld_attr = round_up(N_ATTRS, self.vec)
        ws.attrs_pad = buf(B * ws.ld_attr)
        ws.t1pre, ws.t1, ws.dt1, ws.dt1pre = buf(B * 64), buf(B * 64), buf(B * 64), buf(B * 64)
        ws.ae_sum, ws.ae_s = buf(2, 64, dtype=f32), buf(2, 64, dtype=f32)
        ws.ae_mean, ws.ae_rstd = buf(64, dtype=f32), buf(64, dtype=f32)
        ws.s1pre, ws.s1, ws.ds1, ws.ds1pre = buf(M3 * 64), buf(M3 * 64), buf(M3 * 64), buf(M3 * 64)
        ws.ad_sum, ws.ad_s = buf(2, G, 64, dtype=f32), buf(2, G, 64, dtype=f32)
        ws.ad_mean, ws.ad_rstd = buf(G, 64, dtype=f32), buf(G, 64, dtype=f32)
        ws.alogits, ws.probs_attrs = buf(M3 * N_ATTRS, dtype=f32), buf(M3 * N_ATTRS, dtype=f32)
        ws.ld_dalog = round_up(N_ATTRS, self.vec)
        ws.dalog = buf(M3 * ws.ld_dalog)
        ws.dz_attr = buf(M3 * n, dtype=f32)
        return ws

    # ------------------------------------------------------------------ forward
    def run_forward(self, ws, image, attrs, term_types: Sequence[int], eps, training: bool, lambdas, kl_weights,
                    want_probs: bool, with_loss: bool) -> None:
        B, n = ws.B, self.n_latents
        self.begin_forward()
        use_img = any(t != _lib.TERM_TEXT for t in term_types)
        use_att = any(t != _lib.TERM_IMAGE for t in term_types)
        n_img = sum(1 for t in term_types if t != _lib.TERM_TEXT)
        n_att = sum(1 for t in term_types if t != _lib.TERM_IMAGE)
        R = n_img if (training and self.dropout_p > 0 and n_img > 1) else 1
        ws.R, ws.training, ws.use_img, ws.use_att = R, training, use_img, use_att
        ws.image, ws.attrs = image, attrs
        if use_att:
            self.on_mod_stream(lambda: self._attrs_encoder_fwd(ws, attrs, training, n_att))    # beside the image encoder
        if use_img:
            self.features_fwd(ws, image, B, training, n_img)
            # classifier: Linear(6400, 1024) + Swish + Dropout + Linear(1024, 2n)   celeba/model.py:114-119
            self.linear_fwd(ws.enc_act[3], 6400, B, "image_encoder.classifier.0", 1024, 6400, ws.f1pre, 1024)
            _ops.act_forward(SWISH, ws.f1pre, ws.f1, B, 1024, repeat=R, dropout_p=self.dropout_p if training else 0.0,
                             seed=self.noise_seed, step_counter=self._step_counter)
            self.linear_fwd(ws.f1, 1024, R * B, "image_encoder.classifier.3", 2 * n, 1024, ws.encA, 2 * n)
        self.join_mod_stream()
        self.latent_forward(ws, term_types, kl_weights, eps, training, ws.encA if use_img else None,
                            ws.encB if use_att else None, R)
        self.decode(ws, training, lambdas, want_probs, with_loss)

    def _attrs_encoder_fwd(self, ws, attrs, training: bool, n_att: int) -> None:
        """celeba/model.py:170-176."""
        B, n = ws.B, self.n_latents
        _ops.cast_pad_2d(attrs, B, N_ATTRS, N_ATTRS, ws.attrs_pad, ws.ld_attr)
        bn = "attrs_encoder.net.1"
        self.linear_fwd(ws.attrs_pad, ws.ld_attr, B, "attrs_encoder.net.0", 64, N_ATTRS, ws.t1pre, 64)
        rm, rv = self.running(bn)
        ws.ae_bn = _ops.bn_args(ws.t1pre, B, 64, B, SWISH, training, self.P(bn + ".weight"), self.P(bn + ".bias"),
                                ws.ae_sum[0], ws.ae_sum[1], ws.ae_mean, ws.ae_rstd, rm, rv, updates=n_att)
        _ops.bn_act_forward(ws.ae_bn, ws.t1)
        self.linear_fwd(ws.t1, 64, B, "attrs_encoder.net.3", 2 * n, 64, ws.encB, 2 * n)

    def decode(self, ws, training: bool, lambdas, want_probs: bool, with_loss: bool) -> None:
        """Image and attribute decoders on the stacked [G*B] latents (celeba/model.py:134-163, 185-200) + BCE terms."""
        B, G, n = ws.B, ws.G, self.n_latents
        M3 = G * B
        self.on_mod_stream(lambda: self._attrs_decoder_fwd(ws, training, lambdas, want_probs, with_loss))   # beside the image decoder
        self.linear_fwd(ws.z, ws.ld_z, M3, "image_decoder.upsample.0", 6400, n, ws.u1pre, 6400)
        _ops.act_forward(SWISH, ws.u1pre, ws.u1, M3, 6400)
        self.hallucinate_fwd(ws, M3, B, training)
        sx = [float(lambdas[g][0]) / (B * 12288) for g in range(G)]
        _ops.sigmoid_bce(ws.logits, 12288, M3, 12288, rows_per_group=B,
                         target=ws.image if with_loss else None, ld_target=12288, target_rows=B, grad_scale=sx,
                         loss=ws.acc[0] if with_loss else None, probs=ws.probs_image if want_probs else None, ld_probs=12288,
                         dlogits=ws.logits if with_loss else None, ld_dlogits=12288)
        self.join_mod_stream()

    def _attrs_decoder_fwd(self, ws, training: bool, lambdas, want_probs: bool, with_loss: bool) -> None:
        """celeba/model.py:185-200 + the attribute BCE of celeba/train.py:69-74."""
        B, G, n = ws.B, ws.G, self.n_latents
        M3 = G * B
        sy = [float(lambdas[g][1]) / (B * N_ATTRS) for g in range(G)]
        bn = "attrs_decoder.net.1"
        self.linear_fwd(ws.z, ws.ld_z, M3, "attrs_decoder.net.0", 64, n, ws.s1pre, 64)
        rm, rv = self.running(bn)
        ws.ad_bn = _ops.bn_args(ws.s1pre, M3, 64, B, SWISH, training, self.P(bn + ".weight"), self.P(bn + ".bias"),
                                ws.ad_sum[0], ws.ad_sum[1], ws.ad_mean, ws.ad_rstd, rm, rv, updates=1)
        _ops.bn_act_forward(ws.ad_bn, ws.s1)
        self.linear_fwd(ws.s1, 64, M3, "attrs_decoder.net.3", N_ATTRS, 64, ws.alogits, N_ATTRS)
        _ops.sigmoid_bce(ws.alogits, N_ATTRS, M3, N_ATTRS, rows_per_group=B,
                         target=ws.attrs if with_loss else None, ld_target=N_ATTRS, target_rows=B, grad_scale=sy,
                         loss=ws.acc[1] if with_loss else None, probs=ws.probs_attrs if want_probs else None, ld_probs=N_ATTRS,
                         dlogits=ws.dalog if with_loss else None, ld_dlogits=ws.ld_dalog)

    # ------------------------------------------------------------------ backward (celeba/train.py:151-152)
    def backward_decoders(self, ws) -> None:
        B, G, n = ws.B, ws.G, self.n_latents
        M3 = G * B
        Gd = self.G
        # attribute decoder beside the image decoder; its latent gradient goes to dz_attr and is added after the join
        self.on_mod_stream(lambda: self._attrs_decoder_bwd(ws))
        # image decoder (the gradient of the BCE sits in ws.logits)
        self.hallucinate_bwd(ws, M3)
        _ops.act_backward(SWISH, ws.u1pre, ws.du1, ws.du1pre, M3, 6400, dbias=Gd("image_decoder.upsample.0.bias"))
        self.linear_bwd(ws.z, ws.ld_z, ws.du1pre, 6400, M3, "image_decoder.upsample.0", 6400, n, dx=ws.dz, lddx=n, bias=False)
        self.join_mod_stream()
        _ops.copy_2d(ws.dz_attr, 0, n, ws.dz, 0, n, M3, n, accumulate=True)
        self.latent_backward(ws, ws.dencA if ws.use_img else None, ws.dencB if ws.use_att else None, *getattr(ws, "upstream", (None, None)))

    def _attrs_decoder_bwd(self, ws) -> None:
        B, G, n = ws.B, ws.G, self.n_latents
        M3 = G * B
        Gd = self.G
        self.linear_bwd(ws.s1, 64, ws.dalog, ws.ld_dalog, M3, "attrs_decoder.net.3", N_ATTRS, 64, dx=ws.ds1, lddx=64)
        bn = "attrs_decoder.net.1"
        _ops.bn_act_backward(ws.ad_bn, ws.ds1, ws.ds1pre, ws.ad_s[0], ws.ad_s[1], Gd(bn + ".weight"), Gd(bn + ".bias"))
        self.linear_bwd(ws.z, ws.ld_z, ws.ds1pre, 64, M3, "attrs_decoder.net.0", 64, n, dx=ws.dz_attr, lddx=n, bias=False)

    def module_outputs(self, ws):
        B, n = ws.B, self.n_latents
        return (ws.probs_image.view(B, 3, 64, 64).clone(), ws.probs_attrs.view(B, N_ATTRS).clone(),
                ws.mu.view(1, B, n)[0].clone(), ws.logvar.view(1, B, n)[0].clone())

    def module_backward(self, ws, g_image, g_attrs, g_mu, g_logvar) -> None:
        """Backward from the gradients of (image_recon probs, attrs_recon probs, mu, logvar) - the reference's
        loss.backward() (celeba/train.py:151) when the loss was built by the caller from forward()'s outputs."""
        B = ws.B
        if g_image is None:
            ws.logits.zero_()
        else:   # d logits = d probs * p * (1 - p), in place over the logits
            _ops.sigmoid_bce(ws.logits, 12288, B, 12288, dprobs=g_image.reshape(B, 12288), ld_dprobs=12288, dlogits=ws.logits,
                             ld_dlogits=12288)
        if g_attrs is None:
            ws.dalog.zero_()
        else:
            _ops.sigmoid_bce(ws.alogits, N_ATTRS, B, N_ATTRS, dprobs=g_attrs.reshape(B, N_ATTRS), ld_dprobs=N_ATTRS, dlogits=ws.dalog,
                             ld_dlogits=ws.ld_dalog)
        if g_mu is not None and g_logvar is None:
            g_logvar = torch.zeros_like(g_mu)
        if g_logvar is not None and g_mu is None:
            g_mu = torch.zeros_like(g_logvar)
        ws.upstream = (g_mu, g_logvar)
        self.backward_decoders(ws)
        self.backward_encoders(ws)

    def backward_encoders(self, ws) -> None:
        B, n, R = ws.B, self.n_latents, ws.R
        Gd = self.G
        def attrs_encoder_bwd():
            self.linear_bwd(ws.t1, 64, ws.dencB, ws.ld_enc, B, "attrs_encoder.net.3", 2 * n, 64, dx=ws.dt1, lddx=64)
            bn = "attrs_encoder.net.1"
            _ops.bn_act_backward(ws.ae_bn, ws.dt1, ws.dt1pre, ws.ae_s[0], ws.ae_s[1], Gd(bn + ".weight"), Gd(bn + ".bias"))
            self.linear_bwd(ws.attrs_pad, ws.ld_attr, ws.dt1pre, 64, B, "attrs_encoder.net.0", 64, N_ATTRS, bias=False)

        if ws.use_att:
            self.on_mod_stream(attrs_encoder_bwd)              # beside the image encoder's backward
        if ws.use_img:
            RB = R * B
            self.linear_bwd(ws.f1, 1024, ws.dencA, ws.ld_enc, RB, "image_encoder.classifier.3", 2 * n, 1024, dx=ws.df1, lddx=1024)
            _ops.act_backward(SWISH, ws.f1pre, ws.df1, ws.df1pre, B, 1024, repeat=R,
                              dropout_p=self.dropout_p if ws.training else 0.0, seed=self.noise_seed,
                              step_counter=self._step_counter, dbias=Gd("image_encoder.classifier.0.bias"))
            self.linear_bwd(ws.enc_act[3], 6400, ws.df1pre, 1024, B, "image_encoder.classifier.0", 1024, 6400,
                            dx=ws.enc_dact[3], lddx=6400, bias=False)
            self.features_bwd(ws, B)
        self.join_mod_stream()

    # ------------------------------------------------------------------ module surface
    def forward(self, image: Optional[torch.Tensor] = None, attrs: Optional[torch.Tensor] = None, eps: Optional[torch.Tensor] = None):
        """celeba/model.py:36-58: returns (image_recon, attrs_recon, mu, logvar).  In train mode with autograd enabled the
        outputs are differentiable (the reference's loop `loss_function(...).backward(); optimizer.step()` works unchanged;
        CelebATrainer.step is the fused fast path); the reparametrize noise is drawn in-kernel unless `eps` [B, n] is given."""
        assert image is not None or attrs is not None
        t = _lib.TERM_JOINT if (image is not None and attrs is not None) else (_lib.TERM_IMAGE if image is not None else _lib.TERM_TEXT)
        B = (image if image is not None else attrs).shape[0]
        image = None if image is None else image.detach().to(self.device, torch.float32).contiguous()
        attrs = None if attrs is None else attrs.detach().to(self.device, torch.float32).contiguous()
        if eps is not None:
            eps = eps.to(self.device, torch.float32).contiguous()
        if self.training and torch.is_grad_enabled():
            return self._autograd_forward(image, attrs, t, eps)
        ws = self.workspace(B, 1)
        self.run_forward(ws, image, attrs, (t,), eps, self.training, ((0.0, 0.0),), (0.0,), True, False)
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
        self.begin_forward()
        self.decode(ws, self.training, ((0.0, 0.0),), True, False)
        return ws.probs_image.view(B, 3, 64, 64).clone(), ws.probs_attrs.view(B, N_ATTRS).clone()


class CelebATrainer(ConvMVAETrainer):
    """celeba/train.py:132-157.  step(image [B,3,64,64], attrs [B,18]) with the reference's defaults (all lambdas 1,
    kl_lambda 1e-3); `eps` optionally injects the reparametrize noise [n_terms, B, n]."""

    def _prepare(self, image, attrs):
        m = self.model
        return image.to(m.device, torch.float32).contiguous(), attrs.to(m.device, torch.float32).contiguous()

    def step(self, image, attrs, terms: Sequence[str] = ("joint", "image", "attrs"),
             lambdas: Sequence[Tuple[float, float]] = ((1.0, 1.0),) * 3, eps: Optional[torch.Tensor] = None, adam: bool = True):
        return super().step(image, attrs, terms, lambdas, eps, adam)

    def losses(self) -> List[Tuple[float, float, float, float]]:
        """Per-term (total, image BCE term, attrs BCE term, KL term) of the last step (one small D2H copy);
        celeba/train.py:60-81."""
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


def loss_function(mu, logvar, recon_x=None, x=None, recon_y=None, y=None, kl_lambda=1e-3, lambda_x=1.0, lambda_y=1.0):
    """celeba/train.py:60-81 on the module outputs (probabilities), differentiable; the arithmetic is the library's loss
    kernels (mvae_elbo_loss_forward / _backward), not torch ops."""
    from .functional import _ElboFn
    B = mu.shape[0]
    total = _ElboFn.apply(mu, logvar, recon_x, x, None, None, float(lambda_x), 0.0, float(kl_lambda) / B)
    if recon_y is not None and y is not None:
        z = torch.zeros(B, 1, device=mu.device)
        total = total + _ElboFn.apply(z, z, recon_y, y, None, None, float(lambda_y), 0.0, 0.0)
    return total
