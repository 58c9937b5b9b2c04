"""MultiMNIST MVAE (50x50 conv image encoder / decoder + <= 4 character text through GRUs) on the B200-native library.

Reference surface kept (multimnist/model.py:20-93, multimnist/train.py:69-87,148-175):
    MultimodalVAE(n_latents=20, use_cuda=False).forward(image=None, text=None)
        -> (image_recon [B,1,50,50] probs, text_recon [B,4,12] log-probs, mu, logvar)
    state_dict()/load_state_dict() with the reference's keys and shapes
    MultiMNISTTrainer.step(image, text): zero_grad + vae(image, text) + vae(image=image) + vae(text=text) + the three
        loss_function calls (lambdas (1,1), (1,.5), (0,1)) + backward + Adam.

The conv stacks, parameter store and trainer come from convnet.py.  The text side is composed from tcgen05 GEMMs (the
two projections of every GRU cell, z2h, h2p, h2o) and the kernels of csrc/text_ops.cu (embedding, GRU gate math,
log_softmax + NLL + greedy argmax).  The text encoder runs once per step on [B] rows (the reference runs it twice on
identical inputs); the autoregressive text decoder runs its 4 steps once on the stacked [3B] latents.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib, _ops
from .convnet import ConvMVAEBase, ConvMVAETrainer, Workspace, SWISH, round_up

N_CHARS, SOS, FILL, MAX_LEN, H = 12, 10, 11, 4, 100   # multimnist/utils.py:14-19, model.py:25-29
NONE = _lib.ACT_NONE


class MultimodalVAE(ConvMVAEBase):
    """Drop-in for multimnist/model.py:20-93.  `precision`: "bf16" (default) or "tf32" (fp32 storage; the parity path)."""

    TERMS = {"joint": _lib.TERM_JOINT, "image": _lib.TERM_IMAGE, "text": _lib.TERM_TEXT}
    IMG_C, IMG_H = 1, 50
    ENC_CONVS = (("image_encoder.features.0", 1, 32, 4, 2, 1, 50, None),
                 ("image_encoder.features.2", 32, 64, 4, 2, 1, 25, "image_encoder.features.3"),
                 ("image_encoder.features.5", 64, 128, 4, 2, 1, 12, "image_encoder.features.6"),
                 ("image_encoder.features.8", 128, 256, 4, 2, 0, 6, "image_encoder.features.9"))
    DEC_CONVS = (("image_decoder.hallucinate.0", 256, 128, 4, 2, 0, 6, "image_decoder.hallucinate.1"),
                 ("image_decoder.hallucinate.3", 128, 64, 4, 2, 1, 12, "image_decoder.hallucinate.4"),
                 ("image_decoder.hallucinate.6", 64, 32, 5, 2, 1, 25, "image_decoder.hallucinate.7"),
                 ("image_decoder.hallucinate.9", 32, 1, 4, 2, 1, 50, None))
    FLAT_C, FLAT_HW = 256, 4
    BN_LAYERS = {"image_encoder.features.3": 64, "image_encoder.features.6": 128, "image_encoder.features.9": 256,
                 "image_decoder.hallucinate.1": 128, "image_decoder.hallucinate.4": 64, "image_decoder.hallucinate.7": 32}

    def __init__(self, n_latents: int = 20, use_cuda: bool = True, precision: str = "bf16", dropout_p: float = 0.1,
                 device: Optional[torch.device] = None, seed: int = 0):
        super().__init__(n_latents, precision, dropout_p, device, seed)

    def reference_keys(self, n_latents: int) -> List[Tuple[str, Tuple[int, ...], str]]:
        """(key, reference shape, kind) in the reference's state_dict order (multimnist/model.py:22-31)."""
        n = n_latents
        out: List[Tuple[str, Tuple[int, ...], str]] = []

        def bn(p, c):
            out.extend([(p + ".weight", (c,), "plain"), (p + ".bias", (c,), "plain"), (p + ".running_mean", (c,), "rm"),
                        (p + ".running_var", (c,), "rv"), (p + ".num_batches_tracked", (), "nbt")])

        def lin(p, o, i, wkind="plain", bkind="plain"):
            out.extend([(p + ".weight", (o, i), wkind), (p + ".bias", (o,), bkind)])

        def gru(p, layer, inp, suffix=""):
            out.extend([("%s.weight_ih_l%d%s" % (p, layer, suffix), (3 * H, inp), "plain"),
                        ("%s.weight_hh_l%d%s" % (p, layer, suffix), (3 * H, H), "plain"),
                        ("%s.bias_ih_l%d%s" % (p, layer, suffix), (3 * H,), "plain"),
                        ("%s.bias_hh_l%d%s" % (p, layer, suffix), (3 * H,), "plain")])

        for pre, ci, co, k, _, _, _, b in self.ENC_CONVS:
            out.append((pre + ".weight", (co, ci, k, k), "conv"))
            if b:
                bn(b, co)
        lin("image_encoder.classifier.0", 400, 1024, "fc_in")
        lin("image_encoder.classifier.3", 200, 400)
        lin("image_encoder.classifier.6", 2 * n, 200)
        lin("image_decoder.upsample.0", 1024, n, "fc_out", "fc_out_bias")
        for pre, ci, co, k, _, _, _, b in self.DEC_CONVS:
            out.append((pre + ".weight", (ci, co, k, k), "convT"))
            if b:
                bn(b, co)
        out.append(("text_encoder.embed.weight", (N_CHARS, H), "plain"))
        gru("text_encoder.gru", 0, H)
        gru("text_encoder.gru", 0, H, "_reverse")
        lin("text_encoder.h2p", 2 * n, H)
        out.append(("text_decoder.embed.weight", (N_CHARS, H), "plain"))
        lin("text_decoder.z2h", H, n)
        gru("text_decoder.gru", 0, H + n)
        gru("text_decoder.gru", 1, H)
        lin("text_decoder.h2o", N_CHARS, H + n)
        return out

    def _init_tensor(self, key, shape, kind, g, sd):
        if key.endswith("embed.weight"):
            return torch.randn(shape, generator=g)
        if ".gru." in key:
            return (torch.rand(shape, generator=g) * 2 - 1) / H ** 0.5
        return super()._init_tensor(key, shape, kind, g, sd)

    def bn_increments(self, term_types) -> List[int]:
        ni = sum(1 for t in term_types if t != _lib.TERM_TEXT)
        g = len(term_types)
        return [ni, ni, ni, g, g, g]

    def linear_shapes(self, n_terms: int, n_img_terms: int):
        n = self.n_latents
        return [(400, 1024, 1), (200, 400, n_img_terms), (2 * n, 200, n_img_terms), (1024, n, n_terms),
                (3 * H, H, 5), (3 * H, H, 4), (2 * n, H, 1),                                  # text encoder cells + h2p
                (H, n, n_terms), (3 * H, H + n, 4 * n_terms), (3 * H, H, 12 * n_terms), (N_CHARS, H + n, 4 * n_terms)]

    # ------------------------------------------------------------------ workspace
    def _make_workspace(self, B: int, G: int) -> Workspace:
        ws = Workspace()
        ws.B, ws.G, ws.R = B, G, 1
        self.alloc_conv_buffers(ws, B, G)
        buf, n, f32, dev = ws.buf, self.n_latents, torch.float32, self.device
        M3, Rmax = G * B, 2
        ws.f1pre, ws.f1, ws.df1, ws.df1pre = buf(B * 400), buf(Rmax * B * 400), buf(Rmax * B * 400), buf(B * 400)
        ws.f2pre, ws.f2, ws.df2, ws.df2pre = buf(Rmax * B * 200), buf(Rmax * B * 200), buf(Rmax * B * 200), buf(Rmax * B * 200)
        ws.encA, ws.encB = buf(Rmax * B * 2 * n, dtype=f32), buf(B * 2 * n, dtype=f32)
        ws.dencA, ws.dencB = buf(Rmax * B * ws.ld_enc), buf(B * ws.ld_enc)
        ldH = ws.ldH = round_up(H, self.vec)
        ldc = ws.ldc = round_up(H + n, self.vec)
        ldg = ws.ldg = round_up(3 * H, self.vec)
        rows = max(M3, B)
        T = MAX_LEN

        def slices(flat, per, count):
            return [flat[i * per:(i + 1) * per] for i in range(count)]

        ws.gi, ws.gh = buf(T * rows * 3 * H, dtype=f32), buf(rows * 3 * H, dtype=f32)
        # gate gradients of every time step, stacked time-major: ONE weight-gradient GEMM / bias reduction per weight
        ws.dgi_all = [buf(T * rows * ldg), buf(T * rows * ldg)]
        ws.dgh_all = [buf(T * rows * ldg), buf(T * rows * ldg)]
        ws.dgi_r, ws.dgh_r = buf(B * ldg), buf(B * ldg)
        ws.gi_r = buf(B * 3 * H, dtype=f32)
        # text encoder (rows B): ex_all slot t = embedded character t; hf_all slot 0 = zeros, slot t+1 = h_t
        ws.ex_all = buf(T * B * ldH)
        ws.ex = slices(ws.ex_all, B * ldH, T)
        ws.zeros_h = buf(rows * ldH)
        ws.hf_all = buf((T + 1) * B * ldH)
        ws.hf = slices(ws.hf_all, B * ldH, T + 1)[1:]
        ws.hsum = buf(B * ldH)
        ws.te_saved = [buf(B * 4 * H, dtype=f32) for _ in range(T + 1)]
        ws.te_cells = [None] * (T + 1)
        ws.dhsum, ws.te_carry = buf(B * H, dtype=f32), buf(B * H, dtype=f32)
        ws.dex_all = buf(T * B * H, dtype=f32)
        ws.dex = slices(ws.dex_all, B * H, T)
        # text decoder (rows M3): h*_all slot 0 = z2h(z), slot t+1 = h_t of the layer
        ws.cat1_all, ws.cat2_all = buf(T * M3 * ldc), buf(T * M3 * ldc)
        ws.cat1, ws.cat2 = slices(ws.cat1_all, M3 * ldc, T), slices(ws.cat2_all, M3 * ldc, T)
        ws.h0_all, ws.h1_all = buf((T + 1) * M3 * ldH), buf((T + 1) * M3 * ldH)
        ws.h_init = ws.h0_all[:M3 * ldH]
        ws.h0, ws.h1 = slices(ws.h0_all, M3 * ldH, T + 1)[1:], slices(ws.h1_all, M3 * ldH, T + 1)[1:]
        ws.h0d_all = buf(T * M3 * ldH)
        ws.h0d = slices(ws.h0d_all, M3 * ldH, T)
        ws.td_saved = [[buf(M3 * 4 * H, dtype=f32) for _ in range(T)] for _ in range(2)]
        ws.td_cells = [[None] * T for _ in range(2)]
        ws.c_in = [torch.full((M3,), SOS, device=dev, dtype=torch.int64) for _ in range(T + 1)]
        ws.tlogits = buf(M3 * N_CHARS, dtype=f32)
        ws.words = buf(M3 * T * N_CHARS, dtype=f32)
        ws.ld_dlog = round_up(N_CHARS, self.vec)
        ws.dlog_all = buf(T * M3 * ws.ld_dlog)
        ws.dlog = slices(ws.dlog_all, M3 * ws.ld_dlog, T)
        ws.dcat = buf(M3 * (H + n), dtype=f32)
        ws.dz_text = buf(M3 * n, dtype=f32)
        ws.carry = [buf(M3 * H, dtype=f32), buf(M3 * H, dtype=f32)]
        ws.dx1, ws.dx1_t, ws.dx1d_t = buf(M3 * H, dtype=f32), buf(M3 * ldH), buf(M3 * ldH)
        ws.dhinit = buf(M3 * ldH)
        return ws

    # ------------------------------------------------------------------ GRU helpers
    def _cell_fwd(self, ws, rows, x, ldx, n_in, h_prev, prefix, suffix, saved, h_out, h_out2=None, ld_h_out2=0, addend=None,
                  gi=None):
        """One GRU cell: gi = x W_ih^T + b_ih (or precomputed for all steps), gh = h W_hh^T + b_hh (tcgen05 GEMMs), gates."""
        wh, ldwh = self.operand("%s.weight_hh_%s" % (prefix, suffix), 3 * H, H)
        if gi is None:
            wi, ldwi = self.operand("%s.weight_ih_%s" % (prefix, suffix), 3 * H, n_in)
            gi = ws.gi
            _ops.gemm(x, wi, gi, rows, 3 * H, n_in, ldx, ldwi, 3 * H, bias=self.P("%s.bias_ih_%s" % (prefix, suffix)))
        hp = h_prev if h_prev is not None else ws.zeros_h
        _ops.gemm(hp, wh, ws.gh, rows, 3 * H, H, ws.ldH, ldwh, 3 * H, bias=self.P("%s.bias_hh_%s" % (prefix, suffix)))
        return _ops.gru_cell_forward(rows, H, gi, ws.gh, h_prev, ws.ldH, h_out, ws.ldH, saved, h_out2, ld_h_out2, addend, ws.ldH)

    def _cell_bwd(self, ws, rows, cell, h_prev, prefix, suffix, dh_a, ld_dh_a, dh_b, carry, dgi, dgh, n_in=0, dx=None, lddx=0,
                  accumulate_dx=False):
        """Backward of one cell's gate math and its data gradients: dgi / dgh land in the caller's (time-stacked) buffers,
        carry = dh*z + dgh W_hh (gradient at h_prev), dx (=|+=) dgi W_ih.  Parameter gradients: _cells_param_grads."""
        _ops.gru_cell_backward(cell, dh_a, ld_dh_a, dh_b, H, dgi, dgh, ws.ldg, carry, H)
        ki, kh = "%s.weight_ih_%s" % (prefix, suffix), "%s.weight_hh_%s" % (prefix, suffix)
        if h_prev is not None and carry is not None:
            wh, ldwh = self._operand_cached(kh, H)
            _ops.gemm(dgh, wh, carry, rows, H, 3 * H, ws.ldg, ldwh, H, b_major=1, accumulate=True)
        if dx is not None:
            wi, ldwi = self._operand_cached(ki, n_in)
            _ops.gemm(dgi, wi, dx, rows, n_in, 3 * H, ws.ldg, ldwi, lddx, b_major=1, accumulate=accumulate_dx)

    def _cells_param_grads(self, ws, rows, dgi, dgh, x, ldx, n_in, h_prev, prefix, suffix) -> None:
        """dW_ih += dgi^T x, dW_hh += dgh^T h_prev, db_ih += colsum(dgi), db_hh += colsum(dgh) over `rows` stacked rows
        (all time steps of a layer at once)."""
        G = self.G
        _ops.gemm(dgi, x, G("%s.weight_ih_%s" % (prefix, suffix)), 3 * H, n_in, rows, ws.ldg, ldx, n_in, a_major=1, b_major=1,
                  accumulate=True)
        _ops.col_stats(dgi, rows, ws.ldg, G("%s.bias_ih_%s" % (prefix, suffix)), valid_channels=3 * H)
        _ops.col_stats(dgh, rows, ws.ldg, G("%s.bias_hh_%s" % (prefix, suffix)), valid_channels=3 * H)
        if h_prev is not None:
            _ops.gemm(dgh, h_prev, G("%s.weight_hh_%s" % (prefix, suffix)), 3 * H, H, rows, ws.ldg, ws.ldH, H, a_major=1, b_major=1,
                      accumulate=True)

    # ------------------------------------------------------------------ forward
    def run_forward(self, ws, image, text, term_types: Sequence[int], eps, training: bool, lambdas, kl_weights,
                    want_probs: bool, with_loss: bool) -> None:
        B, n = ws.B, self.n_latents
        self.begin_forward()
        use_img = any(t != _lib.TERM_TEXT for t in term_types)
        use_txt = any(t != _lib.TERM_IMAGE for t in term_types)
        n_img = sum(1 for t in term_types if t != _lib.TERM_TEXT)
        R = n_img if (training and self.dropout_p > 0 and n_img > 1) else 1
        p = self.dropout_p if training else 0.0
        ws.R, ws.training, ws.use_img, ws.use_txt = R, training, use_img, use_txt
        ws.image, ws.text = image, text
        if use_txt:
            self.on_mod_stream(lambda: self._text_encoder_fwd(ws, text, B))     # beside the image encoder
        if use_img:
            self.features_fwd(ws, image, B, training, n_img)
            # classifier: Linear(1024,400) Swish Dropout Linear(400,200) Swish Dropout Linear(200,2n)  multimnist/model.py:172-180
            self.linear_fwd(ws.enc_act[3], 1024, B, "image_encoder.classifier.0", 400, 1024, ws.f1pre, 400)
            _ops.act_forward(SWISH, ws.f1pre, ws.f1, B, 400, repeat=R, dropout_p=p, seed=self.noise_seed + 101,
                             step_counter=self._step_counter)
            self.linear_fwd(ws.f1, 400, R * B, "image_encoder.classifier.3", 200, 400, ws.f2pre, 200)
            _ops.act_forward(SWISH, ws.f2pre, ws.f2, R * B, 200, dropout_p=p, seed=self.noise_seed + 202,
                             step_counter=self._step_counter)
            self.linear_fwd(ws.f2, 200, R * B, "image_encoder.classifier.6", 2 * n, 200, ws.encA, 2 * n)
        self.join_mod_stream()
        self.latent_forward(ws, term_types, kl_weights, eps, training, ws.encA if use_img else None,
                            ws.encB if use_txt else None, R)
        self.decode(ws, training, lambdas, want_probs, with_loss)

    def _text_encoder_fwd(self, ws, text, B) -> None:
        """multimnist/model.py:220-249."""
        n = self.n_latents
        emb = self.P("text_encoder.embed.weight")
        g = "text_encoder.gru"
        for t in range(MAX_LEN):
            _ops.embed_forward(text, t, MAX_LEN, emb, N_CHARS, H, NONE, ws.ex[t], 0, ws.ldH, B)
        wi, ldwi = self.operand(g + ".weight_ih_l0", 3 * H, H)
        _ops.gemm(ws.ex_all, wi, ws.gi, MAX_LEN * B, 3 * H, H, ws.ldH, ldwi, 3 * H, bias=self.P(g + ".bias_ih_l0"))
        for t in range(MAX_LEN):
            ws.te_cells[t] = self._cell_fwd(ws, B, ws.ex[t], ws.ldH, H, ws.hf[t - 1] if t > 0 else None, g, "l0", ws.te_saved[t],
                                            ws.hf[t], gi=ws.gi[t * B * 3 * H:(t + 1) * B * 3 * H])
        # reverse direction at the last position: one cell from a zero state; its output is summed with the forward one
        wr, ldwr = self.operand(g + ".weight_ih_l0_reverse", 3 * H, H)
        _ops.gemm(ws.ex[MAX_LEN - 1], wr, ws.gi_r, B, 3 * H, H, ws.ldH, ldwr, 3 * H, bias=self.P(g + ".bias_ih_l0_reverse"))
        ws.te_cells[MAX_LEN] = self._cell_fwd(ws, B, ws.ex[MAX_LEN - 1], ws.ldH, H, None, g, "l0_reverse", ws.te_saved[MAX_LEN],
                                              ws.hsum, addend=ws.hf[MAX_LEN - 1], gi=ws.gi_r)
        self.linear_fwd(ws.hsum, ws.ldH, B, "text_encoder.h2p", 2 * n, H, ws.encB, 2 * n)

    def decode(self, ws, training: bool, lambdas, want_probs: bool, with_loss: bool) -> None:
        """Image decoder (multimnist/model.py:192-217) and greedy text decoder (:252-307) on the stacked latents."""
        self.on_mod_stream(lambda: self._decode_text(ws, training, lambdas, with_loss))   # beside the image decoder
        B, G, n = ws.B, ws.G, self.n_latents
        M3 = G * B
        npx = self.n_pixels
        self.linear_fwd(ws.z, ws.ld_z, M3, "image_decoder.upsample.0", 1024, n, ws.u1pre, 1024)
        _ops.act_forward(SWISH, ws.u1pre, ws.u1, M3, 1024)
        self.hallucinate_fwd(ws, M3, B, training)
        sx = [float(lambdas[g][0]) / (B * npx) for g in range(G)]
        _ops.sigmoid_bce(ws.logits, npx, M3, npx, rows_per_group=B, target=ws.image if with_loss else None, ld_target=npx,
                         target_rows=B, grad_scale=sx, loss=ws.acc[0] if with_loss else None,
                         probs=ws.probs_image if want_probs else None, ld_probs=npx,
                         dlogits=ws.logits if with_loss else None, ld_dlogits=npx)
        self.join_mod_stream()

    def _decode_text(self, ws, training: bool, lambdas, with_loss: bool) -> None:
        """Greedy 4-step text decoder (multimnist/model.py:252-307) on the stacked latents."""
        B, G, n = ws.B, ws.G, self.n_latents
        M3 = G * B
        sy = [float(lambdas[g][1]) / (B * MAX_LEN) for g in range(G)]
        p = self.dropout_p if training else 0.0
        ws.td_dropout = p
        g = "text_decoder.gru"
        emb = self.P("text_decoder.embed.weight")
        ldH, ldc = ws.ldH, ws.ldc
        wz, ldwz = self.operand("text_decoder.z2h.weight", H, n)
        _ops.gemm(ws.z, wz, ws.h_init, M3, H, n, ws.ld_z, ldwz, ldH, bias=self.P("text_decoder.z2h.bias"))
        _ops.copy_2d(ws.h_init, 0, ldH, ws.h1_all, 0, ldH, M3, H)          # slot 0 of layer 1 = the same initial state
        for t in range(MAX_LEN):
            _ops.embed_forward(ws.c_in[t], 0, 1, emb, N_CHARS, H, SWISH, ws.cat1[t], 0, ldc, M3)
            _ops.copy_2d(ws.z, 0, ws.ld_z, ws.cat1[t], H, ldc, M3, n)
            ws.td_cells[0][t] = self._cell_fwd(ws, M3, ws.cat1[t], ldc, H + n, ws.h0[t - 1] if t > 0 else ws.h_init, g, "l0",
                                               ws.td_saved[0][t], ws.h0[t])
            x1 = ws.h0[t]
            if p > 0:
                _ops.act_forward(NONE, ws.h0[t], ws.h0d[t], M3, ldH, dropout_p=p, seed=self.noise_seed + 303 + t,
                                 step_counter=self._step_counter)
                x1 = ws.h0d[t]
            ws.td_cells[1][t] = self._cell_fwd(ws, M3, x1, ldH, H, ws.h1[t - 1] if t > 0 else ws.h1_all[:M3 * ldH], g, "l1",
                                               ws.td_saved[1][t], ws.h1[t], h_out2=ws.cat2[t], ld_h_out2=ldc)
            _ops.copy_2d(ws.z, 0, ws.ld_z, ws.cat2[t], H, ldc, M3, n)
            self.linear_fwd(ws.cat2[t], ldc, M3, "text_decoder.h2o", N_CHARS, H + n, ws.tlogits, N_CHARS)
            _ops.logsoftmax_nll(ws.tlogits, N_CHARS, M3, N_CHARS, rows_per_group=B,
                                target=ws.text if with_loss else None, target_offset=t, target_stride=MAX_LEN, target_rows=B,
                                grad_scale=sy, loss=ws.acc[1] if with_loss else None, logp=ws.words, logp_offset=t * N_CHARS,
                                ld_logp=MAX_LEN * N_CHARS, argmax=ws.c_in[t + 1],
                                dlogits=ws.dlog[t] if with_loss else None, ld_dlogits=ws.ld_dlog)

    # ------------------------------------------------------------------ backward (multimnist/train.py:167-168)
    def backward_decoders(self, ws) -> None:
        B, G, n = ws.B, ws.G, self.n_latents
        M3 = G * B
        Gd = self.G
        # the text decoder's backward runs beside the image decoder's and collects its latent gradient in dz_text
        self.on_mod_stream(lambda: self._text_decoder_bwd(ws))
        self.hallucinate_bwd(ws, M3)
        _ops.act_backward(SWISH, ws.u1pre, ws.du1, ws.du1pre, M3, 1024, dbias=Gd("image_decoder.upsample.0.bias"))
        self.linear_bwd(ws.z, ws.ld_z, ws.du1pre, 1024, M3, "image_decoder.upsample.0", 1024, n, dx=ws.dz, lddx=n, bias=False)
        self.join_mod_stream()
        _ops.copy_2d(ws.dz_text, 0, n, ws.dz, 0, n, M3, n, accumulate=True)
        self.latent_backward(ws, ws.dencA if ws.use_img else None, ws.dencB if ws.use_txt else None, *getattr(ws, "upstream", (None, None)))

    def _text_decoder_bwd(self, ws) -> None:
        """Backward through time of the text decoder; the gradient at z is STORED into ws.dz_text by the first contribution."""
        B, G, n = ws.B, ws.G, self.n_latents
        M3 = G * B
        Gd = self.G
        g = "text_decoder.gru"
        emb = self.P("text_decoder.embed.weight")
        ldH, ldc, p = ws.ldH, ws.ldc, ws.td_dropout
        Kc = H + n
        T = MAX_LEN
        wo, ldwo = self._operand_cached("text_decoder.h2o.weight", Kc)
        sl = lambda flat, t: flat[t * M3 * ws.ldg:(t + 1) * M3 * ws.ldg]
        for t in range(T - 1, -1, -1):
            last = t == T - 1
            _ops.gemm(ws.dlog[t], wo, ws.dcat, M3, Kc, N_CHARS, ws.ld_dlog, ldwo, Kc, b_major=1)       # d[h1 | z] of step t
            _ops.copy_2d(ws.dcat, H, Kc, ws.dz_text, 0, n, M3, n, accumulate=not last)
            dx1 = ws.dx1_t if p > 0 else ws.dx1
            self._cell_bwd(ws, M3, ws.td_cells[1][t], ws.h1[t - 1] if t > 0 else ws.h1_all[:M3 * ldH], g, "l1", ws.dcat, Kc,
                           None if last else ws.carry[1], ws.carry[1], sl(ws.dgi_all[1], t), sl(ws.dgh_all[1], t), n_in=H,
                           dx=dx1, lddx=ldH if p > 0 else H)
            if p > 0:
                _ops.act_backward(NONE, ws.h0[t], ws.dx1_t, ws.dx1d_t, M3, ldH, dropout_p=p, seed=self.noise_seed + 303 + t,
                                  step_counter=self._step_counter)
                dh0, ld_dh0 = ws.dx1d_t, ldH
            else:
                dh0, ld_dh0 = ws.dx1, H
            self._cell_bwd(ws, M3, ws.td_cells[0][t], ws.h0[t - 1] if t > 0 else ws.h_init, g, "l0", dh0, ld_dh0,
                           None if last else ws.carry[0], ws.carry[0], sl(ws.dgi_all[0], t), sl(ws.dgh_all[0], t), n_in=Kc,
                           dx=ws.dcat, lddx=Kc)
            _ops.embed_backward(ws.c_in[t], 0, 1, emb, N_CHARS, H, SWISH, ws.dcat, 0, Kc, M3, Gd("text_decoder.embed.weight"))
            _ops.copy_2d(ws.dcat, H, Kc, ws.dz_text, 0, n, M3, n, accumulate=True)
        # parameter gradients of the 4 steps at once (time-stacked operands)
        TM = T * M3
        _ops.gemm(ws.dlog_all, ws.cat2_all, Gd("text_decoder.h2o.weight"), N_CHARS, Kc, TM, ws.ld_dlog, ldc, Kc, a_major=1, b_major=1,
                  accumulate=True)
        _ops.col_stats(ws.dlog_all, TM, ws.ld_dlog, Gd("text_decoder.h2o.bias"), valid_channels=N_CHARS)
        x1_all = ws.h0d_all if p > 0 else ws.h0_all[M3 * ldH:]
        self._cells_param_grads(ws, TM, ws.dgi_all[1][:TM * ws.ldg], ws.dgh_all[1][:TM * ws.ldg], x1_all, ldH, H, ws.h1_all, g, "l1")
        self._cells_param_grads(ws, TM, ws.dgi_all[0][:TM * ws.ldg], ws.dgh_all[0][:TM * ws.ldg], ws.cat1_all, ldc, Kc, ws.h0_all, g, "l0")
        # both GRU layers start from h = z2h(z)
        _ops.copy_2d(ws.carry[0], 0, H, ws.dhinit, 0, ldH, M3, H, src2=ws.carry[1], ld_src2=H)
        self.linear_bwd(ws.z, ws.ld_z, ws.dhinit, ldH, M3, "text_decoder.z2h", H, n, dx=ws.dz_text, lddx=n, accumulate_dx=True)

    def module_outputs(self, ws):
        B, n = ws.B, self.n_latents
        return (ws.probs_image.view(B, 1, 50, 50).clone(), ws.words.view(B, MAX_LEN, N_CHARS).clone(),
                ws.mu.view(1, B, n)[0].clone(), ws.logvar.view(1, B, n)[0].clone())

    def module_backward(self, ws, g_image, g_text, g_mu, g_logvar) -> None:
        """Backward from the gradients of (image_recon probs, text_recon log-probs, mu, logvar): the reference's
        loss.backward() (multimnist/train.py:168) when the loss was built by the caller from forward()'s outputs."""
        B = ws.B
        npx = self.n_pixels
        if g_image is None:
            ws.logits.zero_()
        else:
            _ops.sigmoid_bce(ws.logits, npx, B, npx, dprobs=g_image.reshape(B, npx), ld_dprobs=npx, dlogits=ws.logits, ld_dlogits=npx)
        if g_text is None:
            ws.dlog_all.zero_()
        else:
            g_text = g_text.reshape(B, MAX_LEN * N_CHARS)
            for t in range(MAX_LEN):
                _ops.logsoftmax_backward(ws.words, t * N_CHARS, MAX_LEN * N_CHARS, g_text, t * N_CHARS, MAX_LEN * N_CHARS, B, N_CHARS,
                                         ws.dlog[t], ws.ld_dlog)
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
        p = self.dropout_p if ws.training else 0.0
        if ws.use_txt:
            self.on_mod_stream(lambda: self._text_encoder_bwd(ws))            # beside the image encoder's backward
        if ws.use_img:
            self._image_encoder_bwd(ws, p)
        self.join_mod_stream()

    def _text_encoder_bwd(self, ws) -> None:
        B, n = ws.B, self.n_latents
        Gd = self.G
        g = "text_encoder.gru"
        ldH = ws.ldH
        self.linear_bwd(ws.hsum, ldH, ws.dencB, ws.ld_enc, B, "text_encoder.h2p", 2 * n, H, dx=ws.dhsum, lddx=H)
        T = MAX_LEN
        sl = lambda flat, t: flat[t * B * ws.ldg:(t + 1) * B * ws.ldg]
        for t in range(T - 1, -1, -1):
            last = t == T - 1
            self._cell_bwd(ws, B, ws.te_cells[t], ws.hf[t - 1] if t > 0 else None, g, "l0", ws.dhsum if last else None, H,
                           None if last else ws.te_carry, ws.te_carry, sl(ws.dgi_all[0], t), sl(ws.dgh_all[0], t))
        TB = T * B
        dgi_all, dgh_all = ws.dgi_all[0][:TB * ws.ldg], ws.dgh_all[0][:TB * ws.ldg]
        self._cells_param_grads(ws, TB, dgi_all, dgh_all, ws.ex_all, ldH, H, ws.hf_all, g, "l0")
        wi, ldwi = self._operand_cached(g + ".weight_ih_l0", H)
        _ops.gemm(dgi_all, wi, ws.dex_all, TB, H, 3 * H, ws.ldg, ldwi, H, b_major=1)              # gradient at every embedding
        # reverse-direction cell (zero initial state: no W_hh gradient, no carry); its input is the last character too
        self._cell_bwd(ws, B, ws.te_cells[T], None, g, "l0_reverse", ws.dhsum, H, None, None, ws.dgi_r, ws.dgh_r, n_in=H,
                       dx=ws.dex[T - 1], lddx=H, accumulate_dx=True)
        self._cells_param_grads(ws, B, ws.dgi_r, ws.dgh_r, ws.ex[T - 1], ldH, H, None, g, "l0_reverse")
        for t in range(T):
            _ops.embed_backward(ws.text, t, MAX_LEN, self.P("text_encoder.embed.weight"), N_CHARS, H, NONE, ws.dex[t], 0, H, B,
                                Gd("text_encoder.embed.weight"))

    def _image_encoder_bwd(self, ws, p) -> None:
        B, n, R = ws.B, self.n_latents, ws.R
        Gd = self.G
        RB = R * B
        self.linear_bwd(ws.f2, 200, ws.dencA, ws.ld_enc, RB, "image_encoder.classifier.6", 2 * n, 200, dx=ws.df2, lddx=200)
        _ops.act_backward(SWISH, ws.f2pre, ws.df2, ws.df2pre, RB, 200, dropout_p=p, seed=self.noise_seed + 202,
                          step_counter=self._step_counter, dbias=Gd("image_encoder.classifier.3.bias"))
        self.linear_bwd(ws.f1, 400, ws.df2pre, 200, RB, "image_encoder.classifier.3", 200, 400, dx=ws.df1, lddx=400, bias=False)
        _ops.act_backward(SWISH, ws.f1pre, ws.df1, ws.df1pre, B, 400, repeat=R, dropout_p=p, seed=self.noise_seed + 101,
                          step_counter=self._step_counter, dbias=Gd("image_encoder.classifier.0.bias"))
        self.linear_bwd(ws.enc_act[3], 1024, ws.df1pre, 400, B, "image_encoder.classifier.0", 400, 1024, dx=ws.enc_dact[3],
                        lddx=1024, bias=False)
        self.features_bwd(ws, B)

    # ------------------------------------------------------------------ module surface
    def forward(self, image: Optional[torch.Tensor] = None, text: Optional[torch.Tensor] = None, eps: Optional[torch.Tensor] = None):
        """multimnist/model.py:58-93: returns (image_recon, text_recon log-probs [B,4,12], mu, logvar); forward only."""
        assert image is not None or text is not None
        t = _lib.TERM_JOINT if (image is not None and text is not None) else (_lib.TERM_IMAGE if image is not None else _lib.TERM_TEXT)
        B = (image if image is not None else text).shape[0]
        image = None if image is None else image.detach().to(self.device, torch.float32).contiguous()
        text = None if text is None else text.to(self.device, torch.int64).contiguous()
        if eps is not None:
            eps = eps.to(self.device, torch.float32).contiguous()
        if self.training and torch.is_grad_enabled():
            return self._autograd_forward(image, text, t, eps)
        ws = self.workspace(B, 1)
        self.run_forward(ws, image, text, (t,), eps, self.training, ((0.0, 0.0),), (0.0,), True, False)
        n = self.n_latents
        return (ws.probs_image.view(B, 1, 50, 50).clone(), ws.words.view(B, MAX_LEN, N_CHARS).clone(),
                ws.mu.view(1, B, n)[0].clone(), ws.logvar.view(1, B, n)[0].clone())

    __call__ = forward

    def _decode_only(self, z: torch.Tensor):
        B = z.shape[0]
        ws = self.workspace(B, 1)
        ws.z.view(B, ws.ld_z)[:, :self.n_latents].copy_(z.to(self.device))
        self.begin_forward()
        self.decode(ws, self.training, ((0.0, 0.0),), True, False)
        return ws.probs_image.view(B, 1, 50, 50).clone(), ws.words.view(B, MAX_LEN, N_CHARS).clone()

    def decode_image(self, z):
        return self._decode_only(z)[0]

    image_decoder = decode_image

    def decode_text(self, z):
        return self._decode_only(z)[1]

    @property
    def text_decoder(self):
        """`vae.text_decoder(z)` / `vae.text_decoder.generate(z)` (multimnist/train.py:264, model.py:254-296)."""
        return _TextDecoderSurface(self)


class _TextDecoderSurface:
    """The two calls the reference makes on `vae.text_decoder` from outside the model: the greedy decode and `generate`."""

    def __init__(self, model: "MultimodalVAE"):
        self._m = model

    def __call__(self, z):
        """log-probabilities [B, 4, 12] of the greedy decode (multimnist/model.py:254-288)."""
        return self._m.decode_text(z)

    def generate(self, z, generator: Optional[torch.Generator] = None):
        """multimnist/model.py:290-296: one character index per position, drawn from the decoder's distribution; int64
        [B, 4].  The reference passes the LOG-probabilities to torch.multinomial (which rejects negative weights, so the call
        fails there); this samples from exp(log-probabilities), the evident intent."""
        words = self._m.decode_text(z)
        probs = words.reshape(-1, N_CHARS).float().exp()
        return torch.multinomial(probs, 1, generator=generator).view(words.shape[0], MAX_LEN)


class MultiMNISTTrainer(ConvMVAETrainer):
    """multimnist/train.py:148-175 with its lambdas (1,1), (1,.5), (0,1) and kl_lambda 1e-3 (:227)."""

    def _prepare(self, image, text):
        m = self.model
        return image.to(m.device, torch.float32).contiguous(), text.to(m.device, torch.int64).contiguous()

    def step(self, image, text, terms: Sequence[str] = ("joint", "image", "text"),
             lambdas: Sequence[Tuple[float, float]] = ((1.0, 1.0), (1.0, 0.5), (0.0, 1.0)), eps: Optional[torch.Tensor] = None,
             adam: bool = True):
        return super().step(image, text, terms, lambdas, eps, adam)

    def losses(self) -> List[Tuple[float, float, float, float]]:
        """Per-term (total, image BCE term, text NLL term, KL term) of the last step; multimnist/train.py:69-87."""
        ws, tt, lambdas = self._last
        acc = ws.acc.cpu()
        B = ws.B
        out = []
        for g in range(len(tt)):
            x = float(acc[0, g]) * lambdas[g][0] / (B * 2500)
            y = float(acc[1, g]) * lambdas[g][1] / (B * MAX_LEN)
            k = float(acc[2, g])
            out.append((x + y + k, x, y, k))
        return out


def loss_function(mu, logvar, recon_image=None, image=None, recon_text=None, text=None, kl_lambda=1e-3, lambda_xy=1.0,
                  lambda_yx=1.0):
    """multimnist/train.py:69-87 on the module outputs, differentiable, through the library's loss kernels."""
    from .functional import _ElboFn
    B = mu.shape[0]
    total = _ElboFn.apply(mu, logvar, recon_image, image, None, None, float(lambda_xy), 0.0, float(kl_lambda) / B)
    if recon_text is not None and text is not None:
        rt = recon_text.reshape(-1, recon_text.shape[-1])
        z = torch.zeros(rt.shape[0], 1, device=mu.device)
        total = total + _ElboFn.apply(z, z, None, None, rt, text.reshape(-1), 0.0, float(lambda_yx), 0.0)
    return total


# ---------------------------------------------------------------------------- label <-> character-index helpers
def charlist_tensor(charlists, device=None) -> torch.Tensor:
    """multimnist/utils.py:22-37 (`char_tensor` / `charlist_tensor`) for a whole batch at once: each entry is the list of
    digits shown in one image (at most MAX_LEN = 4); returns int64 [batch, 4] with the digits left-aligned and FILL (11)
    in the unused positions - the `text` input of MultimodalVAE.forward / MultiMNISTTrainer.step."""
    out = torch.full((len(charlists), MAX_LEN), FILL, dtype=torch.int64)
    for b, digits in enumerate(charlists):
        digits = [int(d) for d in (digits.tolist() if torch.is_tensor(digits) else digits)]
        if len(digits) > MAX_LEN or any(d < 0 or d > 9 for d in digits):
            raise ValueError("entry %d: at most %d digits in 0..9, got %r" % (b, MAX_LEN, digits))
        if digits:
            out[b, :len(digits)] = torch.tensor(digits, dtype=torch.int64)
    return out if device is None else out.to(device)


def tensor_to_string(indices) -> str:
    """multimnist/utils.py:40-56: digits as characters, SOS as '^', FILL as nothing."""
    return "".join("^" if int(i) == SOS else "" if int(i) == FILL else "0123456789"[int(i)] for i in indices)
