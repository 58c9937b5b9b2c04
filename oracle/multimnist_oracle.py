"""CPU ORACLE (test infrastructure, NOT product code) - MultiMNIST MVAE training step.

Restates multimnist/model.py:20-288 and multimnist/train.py:69-87,148-175 op by op in functional form over a flat
state dict keyed by the reference's state_dict names (GRUs written out as cell equations).  Same rules as
mnist_oracle.py: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it.

Pinned against the real reference: oracle/gen_golden_multimnist.py imports multimnist/model.py from /root/reference
(Dropout.p = 0, GRU.dropout = 0, injected reparametrize noise) and writes tests/golden/multimnist_*.npz;
tests/test_oracle_multimnist.py checks this restatement against them.  Arithmetic: PyTorch ATen fp32 on CPU; gradients
from torch.autograd like multimnist/train.py:168.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from celeba_oracle import batchnorm, swish, product_of_experts, binary_cross_entropy_mean, is_buffer, sample_flat  # noqa: F401

N_CHARS = 12      # multimnist/utils.py:15-19 ('0'..'9', SOS, FILL)
SOS = 10
FILL = 11
MAX_LEN = 4       # multimnist/utils.py:14
N_HID = 100       # multimnist/model.py:25-29
KL_LAMBDA = 1e-3  # multimnist/train.py:227
LAMBDAS = ((1.0, 1.0), (1.0, 0.5), (0.0, 1.0))   # multimnist/train.py:158-166

State = Dict[str, torch.Tensor]

# (key, Cin, Cout, k, stride, pad, BatchNorm key)
ENC_CONVS = (("image_encoder.features.0", 1, 32, 4, 2, 1, None), ("image_encoder.features.2", 32, 64, 4, 2, 1, "image_encoder.features.3"),
             ("image_encoder.features.5", 64, 128, 4, 2, 1, "image_encoder.features.6"),
             ("image_encoder.features.8", 128, 256, 4, 2, 0, "image_encoder.features.9"))
DEC_CONVS = (("image_decoder.hallucinate.0", 256, 128, 4, 2, 0, "image_decoder.hallucinate.1"),
             ("image_decoder.hallucinate.3", 128, 64, 4, 2, 1, "image_decoder.hallucinate.4"),
             ("image_decoder.hallucinate.6", 64, 32, 5, 2, 1, "image_decoder.hallucinate.7"),
             ("image_decoder.hallucinate.9", 32, 1, 4, 2, 1, None))
BN_KEYS = tuple(b for *_, b in ENC_CONVS + DEC_CONVS if b)


def param_shapes(n_latents: int) -> Dict[str, Tuple[int, ...]]:
    """state_dict layout of multimnist/model.py:20-31 (MultimodalVAE), registration order."""
    n, H = n_latents, N_HID
    s: Dict[str, Tuple[int, ...]] = {}

    def bn(p, c):
        s[p + ".weight"] = (c,); s[p + ".bias"] = (c,)
        s[p + ".running_mean"] = (c,); s[p + ".running_var"] = (c,); s[p + ".num_batches_tracked"] = ()

    def lin(p, o, i):
        s[p + ".weight"] = (o, i); s[p + ".bias"] = (o,)

    def gru(p, layer, inp, suffix=""):
        s["%s.weight_ih_l%d%s" % (p, layer, suffix)] = (3 * H, inp)
        s["%s.weight_hh_l%d%s" % (p, layer, suffix)] = (3 * H, H)
        s["%s.bias_ih_l%d%s" % (p, layer, suffix)] = (3 * H,)
        s["%s.bias_hh_l%d%s" % (p, layer, suffix)] = (3 * H,)

    for p, ci, co, k, _, _, b in ENC_CONVS:
        s[p + ".weight"] = (co, ci, k, k)
        if b:
            bn(b, co)
    lin("image_encoder.classifier.0", 400, 1024)
    lin("image_encoder.classifier.3", 200, 400)
    lin("image_encoder.classifier.6", 2 * n, 200)
    lin("image_decoder.upsample.0", 1024, n)
    for p, ci, co, k, _, _, b in DEC_CONVS:
        s[p + ".weight"] = (ci, co, k, k)
        if b:
            bn(b, co)
    s["text_encoder.embed.weight"] = (N_CHARS, H)
    gru("text_encoder.gru", 0, H)
    gru("text_encoder.gru", 0, H, "_reverse")
    lin("text_encoder.h2p", 2 * n, H)
    s["text_decoder.embed.weight"] = (N_CHARS, H)
    lin("text_decoder.z2h", H, n)
    gru("text_decoder.gru", 0, H + n)
    gru("text_decoder.gru", 1, H)
    lin("text_decoder.h2o", N_CHARS, H + n)
    return s


def init_state(n_latents: int, seed: int = 1234, dtype=torch.float32) -> State:
    """Random state with PyTorch-default-like initialisers and perturbed BatchNorm affines."""
    g = torch.Generator().manual_seed(seed)
    st: State = {}
    for k, shp in param_shapes(n_latents).items():
        base = k.rsplit(".", 1)[0]
        if k.endswith("num_batches_tracked"):
            st[k] = torch.zeros((), dtype=torch.int64)
        elif k.endswith("running_mean"):
            st[k] = torch.zeros(shp, dtype=dtype)
        elif k.endswith("running_var"):
            st[k] = torch.ones(shp, dtype=dtype)
        elif base in BN_KEYS:
            st[k] = (1.0 + 0.2 * torch.randn(shp, generator=g, dtype=dtype)) if k.endswith("weight") else 0.1 * torch.randn(shp, generator=g, dtype=dtype)
        elif k.endswith("embed.weight"):
            st[k] = torch.randn(shp, generator=g, dtype=dtype)
        elif ".gru." in k:
            st[k] = (torch.rand(shp, generator=g, dtype=dtype) * 2 - 1) / math.sqrt(N_HID)
        elif len(shp) == 4:
            fan_in = (shp[1] if "encoder" in k else shp[0]) * shp[2] * shp[3]
            st[k] = (torch.rand(shp, generator=g, dtype=dtype) * 2 - 1) / math.sqrt(fan_in)
        elif len(shp) == 2:
            st[k] = (torch.rand(shp, generator=g, dtype=dtype) * 2 - 1) / math.sqrt(shp[1])
        else:
            st[k] = (torch.rand(shp, generator=g, dtype=dtype) * 2 - 1) / math.sqrt(st[k[:-4] + "weight"].shape[1])
    return st


# ----------------------------------------------------------------------------- layers
def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """One nn.GRU step (gate order r, z, n): h' = (1 - z) * n + z * h."""
    H = h.shape[1]
    gi = F.linear(x, w_ih, b_ih)
    gh = F.linear(h, w_hh, b_hh)
    r = torch.sigmoid(gi[:, :H] + gh[:, :H])
    z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    return (1 - z) * n + z * h


def image_encoder(p: State, x, st=None, training=True, drop_masks=(None, None)):
    """multimnist/model.py:157-189."""
    h = x
    for pre, _, _, _, stride, pad, bn in ENC_CONVS:
        h = F.conv2d(h, p[pre + ".weight"], None, stride, pad)
        if bn:
            h = batchnorm(h, p, st, bn, training)
        h = swish(h)
    h = h.reshape(-1, 256 * 2 * 2)
    h = swish(F.linear(h, p["image_encoder.classifier.0.weight"], p["image_encoder.classifier.0.bias"]))
    if drop_masks[0] is not None and training:
        h = h * drop_masks[0]
    h = swish(F.linear(h, p["image_encoder.classifier.3.weight"], p["image_encoder.classifier.3.bias"]))
    if drop_masks[1] is not None and training:
        h = h * drop_masks[1]
    h = F.linear(h, p["image_encoder.classifier.6.weight"], p["image_encoder.classifier.6.bias"])
    n = h.shape[1] // 2
    return h[:, :n], h[:, n:]


def image_decoder_logits(p: State, z, st=None, training=True):
    """multimnist/model.py:192-217 up to the final sigmoid."""
    h = swish(F.linear(z, p["image_decoder.upsample.0.weight"], p["image_decoder.upsample.0.bias"]))
    h = h.view(-1, 256, 2, 2)
    for pre, _, _, _, stride, pad, bn in DEC_CONVS:
        h = F.conv_transpose2d(h, p[pre + ".weight"], None, stride, pad)
        if bn:
            h = swish(batchnorm(h, p, st, bn, training))
    return h


def text_encoder(p: State, text):
    """multimnist/model.py:220-249: Embedding -> bidirectional GRU; the LAST time step's outputs of both directions
    are summed (the reverse direction has then seen only the last character), then Linear."""
    e = p["text_encoder.embed.weight"][text]           # [B, T, H]
    B, T, H = e.shape
    g = "text_encoder.gru."
    hf = torch.zeros(B, H, dtype=e.dtype)
    for t in range(T):
        hf = gru_cell(e[:, t], hf, p[g + "weight_ih_l0"], p[g + "weight_hh_l0"], p[g + "bias_ih_l0"], p[g + "bias_hh_l0"])
    hb = gru_cell(e[:, T - 1], torch.zeros(B, H, dtype=e.dtype), p[g + "weight_ih_l0_reverse"], p[g + "weight_hh_l0_reverse"],
                  p[g + "bias_ih_l0_reverse"], p[g + "bias_hh_l0_reverse"])
    h = F.linear(hf + hb, p["text_encoder.h2p.weight"], p["text_encoder.h2p.bias"])
    n = h.shape[1] // 2
    return h[:, :n], h[:, n:]


def text_decoder(p: State, z, inter_layer_masks=None):
    """multimnist/model.py:252-307: greedy 4-step decode from SOS with a 2-layer GRU; returns log-probs [B, 4, 12].
    inter_layer_masks: optional per-step keep-masks of the GRU's inter-layer dropout (None = dropout off)."""
    B = z.shape[0]
    g = "text_decoder.gru."
    h0 = F.linear(z, p["text_decoder.z2h.weight"], p["text_decoder.z2h.bias"])
    h = [h0, h0]
    c_in = torch.full((B,), SOS, dtype=torch.long)
    words = []
    for i in range(MAX_LEN):
        c = swish(p["text_decoder.embed.weight"][c_in])
        x = torch.cat((c, z), dim=1)
        h[0] = gru_cell(x, h[0], p[g + "weight_ih_l0"], p[g + "weight_hh_l0"], p[g + "bias_ih_l0"], p[g + "bias_hh_l0"])
        x1 = h[0] if inter_layer_masks is None else h[0] * inter_layer_masks[i]
        h[1] = gru_cell(x1, h[1], p[g + "weight_ih_l1"], p[g + "weight_hh_l1"], p[g + "bias_ih_l1"], p[g + "bias_hh_l1"])
        o = F.linear(torch.cat((h[1], z), dim=1), p["text_decoder.h2o.weight"], p["text_decoder.h2o.bias"])
        lp = F.log_softmax(o, dim=1)
        words.append(lp)
        c_in = lp.argmax(dim=1)
    return torch.stack(words, dim=1)


def forward(p: State, image=None, text=None, noise=None, st=None, training=True):
    """multimnist/model.py:58-93.  Returns (image_recon, text_recon log-probs [B,4,12], mu, logvar, image_logits)."""
    assert image is not None or text is not None
    mus, lvs = [], []
    if image is not None:
        m, l = image_encoder(p, image, st, training)
        mus.append(m); lvs.append(l)
    if text is not None:
        m, l = text_encoder(p, text)
        mus.append(m); lvs.append(l)
    mu, logvar = product_of_experts(torch.stack(mus, 0), torch.stack(lvs, 0))
    z = mu + noise * torch.exp(0.5 * logvar) if training else mu
    il = image_decoder_logits(p, z, st, training)
    tr = text_decoder(p, z)
    return torch.sigmoid(il), tr, mu, logvar, il


def loss_function(mu, logvar, recon_image=None, image=None, recon_text=None, text=None, kl_lambda=KL_LAMBDA, lambda_xy=1.0,
                  lambda_yx=1.0):
    """multimnist/train.py:69-87."""
    B = mu.shape[0]
    ib, tb = 0.0, 0.0
    if recon_image is not None and image is not None:
        ib = lambda_xy * binary_cross_entropy_mean(recon_image.reshape(-1, 2500), image.reshape(-1, 2500))
    if recon_text is not None and text is not None:
        tb = lambda_yx * F.nll_loss(recon_text.reshape(-1, recon_text.shape[2]), text.reshape(-1))
    kld = -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp())
    return ib + tb + kld / B * kl_lambda


def train_step(p: State, image, text, noises: Sequence[torch.Tensor], lambdas=LAMBDAS, kl_lambda=KL_LAMBDA):
    """multimnist/train.py:148-168.  Returns (losses[3], grads, new buffers, outs)."""
    work: State = {}
    for k, v in p.items():
        work[k] = v.clone() if is_buffer(k) else v.detach().clone().requires_grad_(True)
    args = ((image, text), (image, None), (None, text))
    losses, outs = [], []
    for k in range(3):
        ri, rt, mu, lv, il = forward(work, args[k][0], args[k][1], noises[k], work, True)
        losses.append(loss_function(mu, lv, ri, image, rt, text, kl_lambda, lambdas[k][0], lambdas[k][1]))
        outs.append((ri, rt, mu, lv, il))
    total = losses[0] + losses[1] + losses[2]
    names = [k for k in work if not is_buffer(k)]
    gs = torch.autograd.grad(total, [work[k] for k in names], allow_unused=True)
    grads = {k: (torch.zeros_like(work[k]) if g is None else g) for k, g in zip(names, gs)}
    buffers = {k: v for k, v in work.items() if is_buffer(k)}
    return [float(l.detach()) for l in losses], grads, buffers, outs


def synthetic_batch(batch: int, n_latents: int, seed: int = 0, dtype=torch.float32):
    """image U[0,1) [B,1,50,50], text randint(0,12) [B,4], three N(0,1) draws."""
    g = torch.Generator().manual_seed(seed)
    image = torch.rand(batch, 1, 50, 50, generator=g, dtype=dtype)
    text = torch.randint(0, N_CHARS, (batch, MAX_LEN), generator=g)
    noises = [torch.randn(batch, n_latents, generator=g, dtype=dtype) for _ in range(3)]
    return image, text, noises
