"""North-star MLP instantiation of the MNIST MVAE: Linear + Swish stacks (784 -> hidden -> hidden -> 2 * n_latents, no
normalisation), precision-weighted ProductOfExperts with the N(0, 1) prior expert, the three subsampled ELBO terms batched
into one pass, elbo_loss(lambda_image, lambda_text, annealing_factor).

The mounted reference has no such model (its MNIST model is 784 -> 400 -> 200 with BatchNorm + ReLU,
mnist/model.py:99-170, and its prior expert is commented out, mnist/model.py:44-51,70-75); this is the configuration
BASELINE.json::north_star describes, built from the reference's blocks: the Sequential skeleton of mnist/model.py:99-170,
Swish of multimnist/model.py:379-381, the product of paper/draft.tex:88, reparametrize mnist/model.py:24-30, the loss terms
of mnist/train.py:64-81.  Its oracle is oracle/mlp_oracle.py ("parity unpinned": there are no reference outputs to pin).

How it runs (every arrow is one launch of the tcgen05 GEMM, csrc/gemm.cu, through mvae_gemm):
  * forward Linear + bias + Swish is ONE kernel (epilogue EPI_STORE_ACT writes the pre-activation for the backward and the
    activation for the next layer); the last image-decoder Linear + sigmoid + BCE + dlogits + bias gradient is ONE kernel
    (EPI_BCE: the logits never reach memory);
  * backward: the input gradient through a Swish is ONE kernel (EPI_DGRAD_ACT: dgrad GEMM, times swish'(pre), column sums =
    the previous Linear's bias gradient); weight gradients are split-K GEMMs with both operands MN-major on a side stream;
  * each encoder runs once for all terms, the decoders once on the stacked [terms * B] latents; the text encoder is evaluated
    per LABEL (ten rows) and gathered by label (mvae_embed_forward), its backward scatter-adds by label first;
  * PoE + prior + reparametrize + KL for all terms: mvae_latent_forward / _backward.
There is no CPU or PyTorch fallback.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib, _ops
from .convnet import ConvMVAEBase, ConvMVAETrainer, Workspace, SWISH, round_up

N_CLASSES = 10
N_PIXELS = 784
NONE = _lib.ACT_NONE


class MVAE(ConvMVAEBase):
    """MVAE(n_latents).forward(image, text) -> (recon_image [B, 784] probabilities, recon_text [B, 10] log-probabilities, mu,
    logvar).  `precision`: "bf16" (default) or "tf32" (fp32 storage)."""

    TERMS = {"joint": _lib.TERM_JOINT, "image": _lib.TERM_IMAGE, "text": _lib.TERM_TEXT}
    IMG_C, IMG_H = 1, 28
    FLAT_C, FLAT_HW = 1, 1
    BN_LAYERS: dict = {}

    def __init__(self, n_latents: int = 64, hidden: int = 512, precision: str = "bf16", prior_expert: bool = True,
                 device: Optional[torch.device] = None, seed: int = 0):
        if hidden % 8 or n_latents % 8:
            raise ValueError("hidden and n_latents must be multiples of 8 (16-byte operand rows)")
        self.hidden = int(hidden)
        super().__init__(n_latents, precision, 0.0, device, seed)
        self.poe_mode, self.prior_expert = _lib.POE_PRECISION, 1 if prior_expert else 0
        self._labels10 = torch.arange(N_CLASSES, device=self.device, dtype=torch.int64)

    def reference_keys(self, n_latents: int):
        n, h = n_latents, self.hidden
        out: List[Tuple[str, Tuple[int, ...], str]] = []

        def lin(p, o, i):
            out.extend([(p + ".weight", (o, i), "plain"), (p + ".bias", (o,), "plain")])

        lin("image_encoder.fc1", h, N_PIXELS); lin("image_encoder.fc2", h, h); lin("image_encoder.fc3", 2 * n, h)
        lin("image_decoder.fc1", h, n); lin("image_decoder.fc2", h, h); lin("image_decoder.fc3", N_PIXELS, h)
        out.append(("text_encoder.embed.weight", (N_CLASSES, h), "plain"))
        lin("text_encoder.fc2", h, h); lin("text_encoder.fc3", 2 * n, h)
        lin("text_decoder.fc1", h, n); lin("text_decoder.fc2", h, h); lin("text_decoder.fc3", N_CLASSES, h)
        return out

    def bn_increments(self, term_types) -> List[int]:
        return [0]

    def linear_shapes(self, n_terms: int, n_img_terms: int):
        """(out, in, rows per sample) of every Linear whose work scales with the batch (algorithmic FLOP count of bench.py)."""
        n, h = self.n_latents, self.hidden
        return [(h, N_PIXELS, 1), (h, h, 1), (2 * n, h, 1), (h, n, n_terms), (h, h, n_terms), (N_PIXELS, h, n_terms),
                (h, n, n_terms), (h, h, n_terms), (N_CLASSES, h, n_terms)]

    # ------------------------------------------------------------------ workspace
    def _make_workspace(self, B: int, G: int) -> Workspace:
        ws = Workspace()
        dev, T, f32 = self.device, self.act_dtype, torch.float32
        n, h, M3 = self.n_latents, self.hidden, G * B

        def buf(*shape, dtype=T):
            return torch.zeros(*shape, device=dev, dtype=dtype)

        ws.B, ws.G, ws.R, ws.buf = B, G, 1, buf
        ws.ld_z, ws.ld_enc = round_up(n, self.vec), round_up(2 * n, self.vec)
        ws.ld_cls = round_up(N_CLASSES, self.vec)
        # one zeroed-per-step region: loss accumulators [3, 4] + the per-label gradient table of the text encoder [10, 2n]
        ws.zero_region = buf(12 + N_CLASSES * 2 * n, dtype=f32)
        ws.acc = ws.zero_region[:12].view(3, 4)
        ws.dtable = ws.zero_region[12:]
        ws.x = buf(B * N_PIXELS)
        ws.e1pre, ws.e1, ws.e2pre, ws.e2 = buf(B * h), buf(B * h), buf(B * h), buf(B * h)
        ws.de2pre, ws.de1pre = buf(B * h), buf(B * h)
        ws.encA, ws.encB = buf(B * 2 * n, dtype=f32), buf(B * 2 * n, dtype=f32)
        ws.dencA, ws.dencB = buf(B * ws.ld_enc), buf(B * ws.ld_enc)
        # text encoder, per label
        ws.t0, ws.t2pre, ws.t2 = buf(N_CLASSES * h), buf(N_CLASSES * h), buf(N_CLASSES * h)
        ws.table = buf(N_CLASSES * 2 * n, dtype=f32)
        ws.dtable_op = buf(N_CLASSES * ws.ld_enc)
        ws.dt2pre, ws.dt0 = buf(N_CLASSES * h), buf(N_CLASSES * h)
        # latent
        ws.z = buf(M3 * ws.ld_z)
        ws.dz, ws.dz_text = buf(M3 * n, dtype=f32), buf(M3 * n, dtype=f32)
        ws.mu, ws.logvar = buf(M3 * n, dtype=f32), buf(M3 * n, dtype=f32)
        # decoders
        ws.d1pre, ws.d1, ws.d2pre, ws.d2 = buf(M3 * h), buf(M3 * h), buf(M3 * h), buf(M3 * h)
        ws.dd2pre, ws.dd1pre = buf(M3 * h), buf(M3 * h)
        ws.dlog = buf(M3 * N_PIXELS)                       # gradient at the image logits (the logits themselves are never stored)
        ws.logits = buf(M3 * N_PIXELS, dtype=f32)          # only the loss-free (eval / module) forward stores logits
        ws.probs_image = buf(M3 * N_PIXELS, dtype=f32)
        ws.s1pre, ws.s1, ws.s2pre, ws.s2 = buf(M3 * h), buf(M3 * h), buf(M3 * h), buf(M3 * h)
        ws.ds2pre, ws.ds1pre = buf(M3 * h), buf(M3 * h)
        ws.tlogits, ws.logp = buf(M3 * N_CLASSES, dtype=f32), buf(M3 * N_CLASSES, dtype=f32)
        ws.dtlog = buf(M3 * ws.ld_cls)
        # per-sample missing-modality masks (MVAETrainer.step(has_image=, has_text=)): presence flags in, per-(term, row) weights
        ws.mask_image = torch.ones(B, device=dev, dtype=torch.uint8)
        ws.mask_text = torch.ones(B, device=dev, dtype=torch.uint8)
        ws.mask_weights = buf(M3, dtype=f32)
        ws.mask_counts = buf(4, dtype=f32)
        ws.row_weight = None          # = ws.mask_weights while a masked step runs
        return ws

    # ------------------------------------------------------------------ Linear (+ Swish) helpers
    def _lin_act(self, x, ldx, M, prefix, n_out, n_in, pre, act):
        """act = swish(pre), pre = x W^T + b: one launch."""
        w, ldw = self.operand(prefix + ".weight", n_out, n_in)
        _ops.gemm(x, w, pre, M, n_out, n_in, ldx, ldw, n_out, bias=self.P(prefix + ".bias"), act=SWISH, act_out=act)

    def _wgrad(self, dy, lddy, x, ldx, M, prefix, n_out, n_in):
        self._wgrad_aside(lambda: _ops.gemm(dy, x, self.G(prefix + ".weight"), n_out, n_in, M, lddy, ldx, n_in, a_major=1,
                                            b_major=1, accumulate=True))

    def _dgrad_act(self, dy, lddy, M, prefix, n_out, n_in, pre_in, d_pre_in, prev_prefix):
        """d(pre of the previous layer) = (dy W) * swish'(pre_in); its column sums are the previous Linear's bias gradient."""
        w, ldw = self._operand_cached(prefix + ".weight", n_in)
        _ops.gemm(dy, w, d_pre_in, M, n_in, n_out, lddy, ldw, n_in, b_major=1, act=SWISH, act_pre=pre_in, ld_act_pre=n_in,
                  col_sum=self.G(prev_prefix + ".bias"))

    # ------------------------------------------------------------------ forward
    def run_forward(self, ws, image, text, term_types: Sequence[int], eps, training: bool, lambdas, kl_weights,
                    want_probs: bool, with_loss: bool) -> None:
        B, n, h = ws.B, self.n_latents, self.hidden
        self.begin_forward()
        use_img = any(t != _lib.TERM_TEXT for t in term_types)
        use_txt = any(t != _lib.TERM_IMAGE for t in term_types)
        ws.training, ws.use_img, ws.use_txt = training, use_img, use_txt
        ws.text = text
        if ws.row_weight is not None:   # masks -> weights on the device: no host sync, fixed launch shapes (graph capturable)
            _ops.mask_weights(ws.mask_image, ws.mask_text, B, term_types, ws.mask_weights, ws.mask_counts)
        if image is not None:
            _ops.cast_pad_2d(image, B, N_PIXELS, N_PIXELS, ws.x, N_PIXELS)
        ws.have_image = image is not None
        if use_txt:
            self.on_mod_stream(lambda: self._text_encoder_fwd(ws, text))
        if use_img:
            self._lin_act(ws.x, N_PIXELS, B, "image_encoder.fc1", h, N_PIXELS, ws.e1pre, ws.e1)
            self._lin_act(ws.e1, h, B, "image_encoder.fc2", h, h, ws.e2pre, ws.e2)
            self.linear_fwd(ws.e2, h, B, "image_encoder.fc3", 2 * n, h, ws.encA, 2 * n)
        self.join_mod_stream()
        self.latent_forward(ws, term_types, kl_weights, eps, training, ws.encA if use_img else None,
                            ws.encB if use_txt else None, 1)
        self.decode(ws, lambdas, want_probs, with_loss)

    def _text_encoder_fwd(self, ws, text) -> None:
        """Per label: table[c] = fc3(swish(fc2(swish(embed[c])))), then encB[b] = table[text[b]]."""
        n, h = self.n_latents, self.hidden
        _ops.embed_forward(self._labels10, 0, 1, self.P("text_encoder.embed.weight"), N_CLASSES, h, SWISH, ws.t0, 0, h, N_CLASSES)
        self._lin_act(ws.t0, h, N_CLASSES, "text_encoder.fc2", h, h, ws.t2pre, ws.t2)
        self.linear_fwd(ws.t2, h, N_CLASSES, "text_encoder.fc3", 2 * n, h, ws.table, 2 * n)
        _ops.embed_forward(text, 0, 1, ws.table, N_CLASSES, 2 * n, NONE, ws.encB, 0, 2 * n, ws.B)

    def decode(self, ws, lambdas, want_probs: bool, with_loss: bool) -> None:
        B, G, n, h = ws.B, ws.G, self.n_latents, self.hidden
        M3 = G * B
        self.on_mod_stream(lambda: self._text_decoder_fwd(ws, lambdas, with_loss))
        self._lin_act(ws.z, ws.ld_z, M3, "image_decoder.fc1", h, n, ws.d1pre, ws.d1)
        self._lin_act(ws.d1, h, M3, "image_decoder.fc2", h, h, ws.d2pre, ws.d2)
        w, ldw = self.operand("image_decoder.fc3.weight", N_PIXELS, h)
        if with_loss and ws.have_image:
            # Linear + sigmoid + BCE(sum over pixels, mean over the batch) + dlogits + bias gradient in one launch
            sx = [float(lambdas[g][0]) / B for g in range(G)]
            _ops.gemm(ws.d2, w, ws.dlog, M3, N_PIXELS, h, h, ldw, N_PIXELS, bias=self.P("image_decoder.fc3.bias"),
                      col_sum=self.G("image_decoder.fc3.bias"), rows_per_group=B,
                      bce=dict(target=ws.x, ld_target=N_PIXELS, target_rows=B, scale=sx, loss=ws.acc[0],
                               probs=None, row_weight=ws.row_weight))
        else:
            _ops.gemm(ws.d2, w, ws.logits, M3, N_PIXELS, h, h, ldw, N_PIXELS, bias=self.P("image_decoder.fc3.bias"))
            if want_probs:
                _ops.sigmoid_bce(ws.logits, N_PIXELS, M3, N_PIXELS, probs=ws.probs_image, ld_probs=N_PIXELS)
        self.join_mod_stream()

    def _text_decoder_fwd(self, ws, lambdas, with_loss: bool) -> None:
        B, G, n, h = ws.B, ws.G, self.n_latents, self.hidden
        M3 = G * B
        sy = [float(lambdas[g][1]) / B for g in range(G)]
        self._lin_act(ws.z, ws.ld_z, M3, "text_decoder.fc1", h, n, ws.s1pre, ws.s1)
        self._lin_act(ws.s1, h, M3, "text_decoder.fc2", h, h, ws.s2pre, ws.s2)
        self.linear_fwd(ws.s2, h, M3, "text_decoder.fc3", N_CLASSES, h, ws.tlogits, N_CLASSES)
        _ops.logsoftmax_nll(ws.tlogits, N_CLASSES, M3, N_CLASSES, rows_per_group=B,
                            target=ws.text if with_loss else None, target_rows=B, grad_scale=sy,
                            loss=ws.acc[1] if with_loss else None, logp=ws.logp, ld_logp=N_CLASSES,
                            dlogits=ws.dtlog if with_loss else None, ld_dlogits=ws.ld_cls,
                            row_weight=ws.row_weight if with_loss else None)

    # ------------------------------------------------------------------ backward
    def backward_decoders(self, ws) -> None:
        B, G, n, h = ws.B, ws.G, self.n_latents, self.hidden
        M3 = G * B
        self.on_mod_stream(lambda: self._text_decoder_bwd(ws))
        self._wgrad(ws.dlog, N_PIXELS, ws.d2, h, M3, "image_decoder.fc3", N_PIXELS, h)
        self._dgrad_act(ws.dlog, N_PIXELS, M3, "image_decoder.fc3", N_PIXELS, h, ws.d2pre, ws.dd2pre, "image_decoder.fc2")
        self._wgrad(ws.dd2pre, h, ws.d1, h, M3, "image_decoder.fc2", h, h)
        self._dgrad_act(ws.dd2pre, h, M3, "image_decoder.fc2", h, h, ws.d1pre, ws.dd1pre, "image_decoder.fc1")
        self._wgrad(ws.dd1pre, h, ws.z, ws.ld_z, M3, "image_decoder.fc1", h, n)
        w, ldw = self._operand_cached("image_decoder.fc1.weight", n)
        _ops.gemm(ws.dd1pre, w, ws.dz, M3, n, h, h, ldw, n, b_major=1)
        self.join_mod_stream()
        _ops.copy_2d(ws.dz_text, 0, n, ws.dz, 0, n, M3, n, accumulate=True)
        self.latent_backward(ws, ws.dencA if ws.use_img else None, ws.dencB if ws.use_txt else None,
                             *getattr(ws, "upstream", (None, None)))
        self._join_side()

    def _text_decoder_bwd(self, ws) -> None:
        B, G, n, h = ws.B, ws.G, self.n_latents, self.hidden
        M3 = G * B
        # the text networks' weight gradients stay on this stream (the side stream belongs to the image networks)
        _ops.gemm(ws.dtlog, ws.s2, self.G("text_decoder.fc3.weight"), N_CLASSES, h, M3, ws.ld_cls, h, h, a_major=1, b_major=1,
                  accumulate=True)
        _ops.col_stats(ws.dtlog, M3, ws.ld_cls, self.G("text_decoder.fc3.bias"), valid_channels=N_CLASSES)
        self._dgrad_act(ws.dtlog, ws.ld_cls, M3, "text_decoder.fc3", N_CLASSES, h, ws.s2pre, ws.ds2pre, "text_decoder.fc2")
        _ops.gemm(ws.ds2pre, ws.s1, self.G("text_decoder.fc2.weight"), h, h, M3, h, h, h, a_major=1, b_major=1, accumulate=True)
        self._dgrad_act(ws.ds2pre, h, M3, "text_decoder.fc2", h, h, ws.s1pre, ws.ds1pre, "text_decoder.fc1")
        _ops.gemm(ws.ds1pre, ws.z, self.G("text_decoder.fc1.weight"), h, n, M3, h, ws.ld_z, n, a_major=1, b_major=1,
                  accumulate=True)
        w, ldw = self._operand_cached("text_decoder.fc1.weight", n)
        _ops.gemm(ws.ds1pre, w, ws.dz_text, M3, n, h, h, ldw, n, b_major=1)

    def backward_encoders(self, ws) -> None:
        B, n, h = ws.B, self.n_latents, self.hidden
        if ws.use_txt:
            self.on_mod_stream(lambda: self._text_encoder_bwd(ws))
        if ws.use_img:
            self._wgrad(ws.dencA, ws.ld_enc, ws.e2, h, B, "image_encoder.fc3", 2 * n, h)
            _ops.col_stats(ws.dencA, B, ws.ld_enc, self.G("image_encoder.fc3.bias"), valid_channels=2 * n)
            self._dgrad_act(ws.dencA, ws.ld_enc, B, "image_encoder.fc3", 2 * n, h, ws.e2pre, ws.de2pre, "image_encoder.fc2")
            self._wgrad(ws.de2pre, h, ws.e1, h, B, "image_encoder.fc2", h, h)
            self._dgrad_act(ws.de2pre, h, B, "image_encoder.fc2", h, h, ws.e1pre, ws.de1pre, "image_encoder.fc1")
            self._wgrad(ws.de1pre, h, ws.x, N_PIXELS, B, "image_encoder.fc1", h, N_PIXELS)
        self.join_mod_stream()
        self._join_side()

    def _text_encoder_bwd(self, ws) -> None:
        n, h = self.n_latents, self.hidden
        # d table[c] = sum of the rows with label c, then the ten-row MLP backward
        _ops.embed_backward(ws.text, 0, 1, ws.table, N_CLASSES, 2 * n, NONE, ws.dencB, 0, ws.ld_enc, ws.B, ws.dtable)
        _ops.cast_pad_2d(ws.dtable, N_CLASSES, 2 * n, 2 * n, ws.dtable_op, ws.ld_enc)
        _ops.gemm(ws.dtable_op, ws.t2, self.G("text_encoder.fc3.weight"), 2 * n, h, N_CLASSES, ws.ld_enc, h, h, a_major=1,
                  b_major=1, accumulate=True)
        _ops.col_stats(ws.dtable_op, N_CLASSES, ws.ld_enc, self.G("text_encoder.fc3.bias"), valid_channels=2 * n)
        self._dgrad_act(ws.dtable_op, ws.ld_enc, N_CLASSES, "text_encoder.fc3", 2 * n, h, ws.t2pre, ws.dt2pre, "text_encoder.fc2")
        _ops.gemm(ws.dt2pre, ws.t0, self.G("text_encoder.fc2.weight"), h, h, N_CLASSES, h, h, h, a_major=1, b_major=1,
                  accumulate=True)
        w, ldw = self._operand_cached("text_encoder.fc2.weight", h)
        _ops.gemm(ws.dt2pre, w, ws.dt0, N_CLASSES, h, h, h, ldw, h, b_major=1)
        _ops.embed_backward(self._labels10, 0, 1, self.P("text_encoder.embed.weight"), N_CLASSES, h, SWISH, ws.dt0, 0, h,
                            N_CLASSES, self.G("text_encoder.embed.weight"))

    # ------------------------------------------------------------------ module surface
    def module_outputs(self, ws):
        B, n = ws.B, self.n_latents
        return (ws.probs_image.view(-1, N_PIXELS)[:B].clone(), ws.logp.view(-1, N_CLASSES)[:B].clone(),
                ws.mu.view(-1, n)[:B].clone(), ws.logvar.view(-1, n)[:B].clone())

    def module_backward(self, ws, g_image, g_text, g_mu, g_logvar) -> None:
        """Backward from the gradients of (recon_image probabilities, recon_text log-probabilities, mu, logvar)."""
        B = ws.B
        if g_image is None:
            ws.dlog.zero_()
        else:
            _ops.sigmoid_bce(ws.logits, N_PIXELS, B, N_PIXELS, dprobs=g_image.reshape(B, N_PIXELS), ld_dprobs=N_PIXELS,
                             dlogits=ws.dlog, ld_dlogits=N_PIXELS)
        _ops.col_stats(ws.dlog, B, N_PIXELS, self.G("image_decoder.fc3.bias"))
        if g_text is None:
            ws.dtlog.zero_()
        else:
            _ops.logsoftmax_backward(ws.logp, 0, N_CLASSES, g_text.reshape(B, N_CLASSES), 0, N_CLASSES, B, N_CLASSES, ws.dtlog,
                                     ws.ld_cls)
        if g_mu is not None and g_logvar is None:
            g_logvar = torch.zeros_like(g_mu)
        if g_logvar is not None and g_mu is None:
            g_mu = torch.zeros_like(g_logvar)
        ws.upstream = (g_mu, g_logvar)
        ws.dtable.zero_()
        self.backward_decoders(ws)
        self.backward_encoders(ws)

    def forward(self, image: Optional[torch.Tensor] = None, text: Optional[torch.Tensor] = None,
                eps: Optional[torch.Tensor] = None):
        assert image is not None or text is not None
        t = _lib.TERM_JOINT if (image is not None and text is not None) else (_lib.TERM_IMAGE if image is not None else _lib.TERM_TEXT)
        B = (image if image is not None else text).shape[0]
        image = None if image is None else image.detach().to(self.device, torch.float32).reshape(B, N_PIXELS).contiguous()
        text = None if text is None else text.detach().to(self.device, torch.int64).contiguous()
        if eps is not None:
            eps = eps.to(self.device, torch.float32).contiguous()
        if self.training and torch.is_grad_enabled():
            return self._autograd_forward(image, text, t, eps)
        ws = self.workspace(B, 1)
        self.run_forward(ws, image, text, (t,), eps, self.training, ((0.0, 0.0),), (0.0,), True, False)
        return self.module_outputs(ws)

    __call__ = forward


class MVAETrainer(ConvMVAETrainer):
    """zero_grad + the three forwards + elbo_loss x 3 + backward + Adam (the loop of mnist/train.py:132-153 on this model) as
    one stream of kernels / one CUDA graph.  `annealing_factor` scales the KL term; `eps` optionally injects the
    reparametrize noise [n_terms, B, n]."""

    def __init__(self, model: MVAE, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, annealing_factor: float = 1.0,
                 use_cuda_graph: bool = False, group=None, overlap: bool = True):
        super().__init__(model, lr, betas, eps, annealing_factor, use_cuda_graph, group, overlap)

    @property
    def annealing_factor(self) -> float:
        return self.kl_lambda

    @annealing_factor.setter
    def annealing_factor(self, v: float) -> None:
        self.kl_lambda = float(v)

    def _prepare(self, image, text):
        m = self.model
        B = image.shape[0]
        return (image.to(m.device, torch.float32).reshape(B, N_PIXELS).contiguous(),
                text.to(m.device, torch.int64).contiguous())

    def step(self, image, text, terms: Sequence[str] = ("joint", "image", "text"),
             lambdas: Sequence[Tuple[float, float]] = ((1.0, 1.0),) * 3, eps: Optional[torch.Tensor] = None, adam: bool = True,
             has_image: Optional[torch.Tensor] = None, has_text: Optional[torch.Tensor] = None):
        """`has_image`, `has_text` ([B] bool, optional): per-SAMPLE missing-modality masks (the reference flips a coin
        per batch, mnist/paired_weak.py:82-104).  Term g counts only the rows that have the modalities it needs (joint: both, image:
        has_image, text: has_text) and its loss is the mean over those rows; the other (term, row) pairs get zero loss and zero
        gradient.  The model has no normalisation layer, so rows are independent: the masks become per-(term, row) weights on the
        device (mvae_mask_weights) inside the same fixed-shape launches - no host sync, CUDA-graph replay and data parallelism work
        unchanged (each rank normalises by its own row counts, like the per-rank batch means of the unmasked step)."""
        m = self.model
        B = image.shape[0]
        ws = m.workspace(B, len(terms))
        masked = has_image is not None or has_text is not None
        if masked:
            for dst, src in ((ws.mask_image, has_image), (ws.mask_text, has_text)):
                if src is None:
                    dst.fill_(1)
                else:
                    dst.copy_(torch.as_tensor(src).reshape(-1).to(torch.uint8), non_blocking=True)
        ws.row_weight = ws.mask_weights if masked else None
        self._key_extra = masked
        return super().step(image, text, terms, lambdas, eps, adam)

    def mask_counts(self) -> List[float]:
        """Rows that counted in each term of the last masked step (one small D2H copy)."""
        ws, tt, _ = self._last
        return ws.mask_counts[:len(tt)].cpu().tolist()

    def losses(self) -> List[Tuple[float, float, float, float]]:
        """Per-term (total, lambda_image * BCE, lambda_text * CE, annealing * KL), each a mean over the batch."""
        ws, tt, lambdas = self._last
        acc = ws.acc.cpu()
        B = ws.B
        out = []
        for g in range(len(tt)):
            x = float(acc[0, g])                            # the BCE epilogue accumulates lambda_image / B * sum already
            y = float(acc[1, g]) * lambdas[g][1] / B
            k = float(acc[2, g])
            out.append((x + y + k, x, y, k))
        return out
