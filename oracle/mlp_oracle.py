"""CPU ORACLE (test infrastructure, NOT product code) - the north-star MLP instantiation of the MNIST MVAE.

PARITY UNPINNED.  BASELINE.json::north_star describes "Linear+Swish MLP stacks (784 -> 512 -> 512 -> 2*n_latents)",
"ProductOfExperts fusion with the prior expert" and "elbo_loss(..., lambda_image, lambda_text, annealing_factor)".
None of that exists in the mounted reference (its MNIST model is 784 -> 400 -> 200 with BatchNorm +
ReLU, the prior expert is commented out, mnist/model.py:44-51,70-75).  This file therefore restates the model from the
reference's own building blocks - it cannot be checked against reference outputs, only block by block:
  * Swish               multimnist/model.py:379-381 (x * sigmoid(x))
  * encoder / decoder   the Sequential skeleton of mnist/model.py:99-170 with (BatchNorm1d, ReLU) replaced by Swish and
                        the widths (400, 200) by (hidden, hidden); the text encoder keeps Embedding -> ... -> Linear(2n)
  * PoE                 the precision-weighted product of paper/draft.tex:88 with the N(0,1) prior expert of
                        mnist/model.py:44-51,70-75 (mnist_oracle.product_of_experts_precision, checked in tests/test_oracle.py
                        against its closed form)
  * reparametrize       mnist/model.py:24-30
  * elbo                BCE summed over pixels (mnist/train.py:70 without the mean), cross entropy (mnist/train.py:75), KL
                        (mnist/train.py:79) scaled by annealing_factor, averaged over the batch
Gradients come from torch.autograd, fp32 on CPU.  Only tests/, smoke() and bench.py's baseline legs may import this module.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from mnist_oracle import product_of_experts_precision, reparametrize, adam_step  # noqa: F401  (same directory)

State = Dict[str, torch.Tensor]
POE_EPS = 1e-8
HIDDEN = 512


def param_shapes(n_latents: int, hidden: int = HIDDEN) -> Dict[str, Tuple[int, ...]]:
    n, h = n_latents, hidden
    s: Dict[str, Tuple[int, ...]] = {}

    def lin(p, o, i):
        s[p + ".weight"] = (o, i)
        s[p + ".bias"] = (o,)

    lin("image_encoder.fc1", h, 784); lin("image_encoder.fc2", h, h); lin("image_encoder.fc3", 2 * n, h)
    lin("image_decoder.fc1", h, n); lin("image_decoder.fc2", h, h); lin("image_decoder.fc3", 784, h)
    s["text_encoder.embed.weight"] = (10, h)
    lin("text_encoder.fc2", h, h); lin("text_encoder.fc3", 2 * n, h)
    lin("text_decoder.fc1", h, n); lin("text_decoder.fc2", h, h); lin("text_decoder.fc3", 10, h)
    return s


def init_state(n_latents: int, hidden: int = HIDDEN, seed: int = 1234) -> State:
    g = torch.Generator().manual_seed(seed)
    st: State = {}
    for k, shp in param_shapes(n_latents, hidden).items():
        fan_in = shp[1] if len(shp) == 2 else st[k[:-4] + "weight"].shape[1]
        st[k] = (torch.rand(shp, generator=g) * 2 - 1) / fan_in ** 0.5
    return st


def swish(x):
    """multimnist/model.py:379-381."""
    return x * torch.sigmoid(x)


def _mlp3(p: State, prefix: str, x):
    h = swish(F.linear(x, p[prefix + ".fc1.weight"], p[prefix + ".fc1.bias"]))
    h = swish(F.linear(h, p[prefix + ".fc2.weight"], p[prefix + ".fc2.bias"]))
    return F.linear(h, p[prefix + ".fc3.weight"], p[prefix + ".fc3.bias"])


def image_encoder(p: State, x):
    n = p["image_encoder.fc3.weight"].shape[0] // 2
    o = _mlp3(p, "image_encoder", x.reshape(x.shape[0], 784))
    return o[:, :n], o[:, n:]


def text_encoder(p: State, text):
    n = p["text_encoder.fc3.weight"].shape[0] // 2
    h = swish(F.embedding(text, p["text_encoder.embed.weight"]))
    h = swish(F.linear(h, p["text_encoder.fc2.weight"], p["text_encoder.fc2.bias"]))
    o = F.linear(h, p["text_encoder.fc3.weight"], p["text_encoder.fc3.bias"])
    return o[:, :n], o[:, n:]


def image_decoder_logits(p: State, z):
    return _mlp3(p, "image_decoder", z)


def text_decoder_logits(p: State, z):
    return _mlp3(p, "text_decoder", z)


def forward(p: State, image=None, text=None, noise=None, training=True, prior: bool = True):
    """MVAE.forward(image, text) -> (recon_image probabilities, recon_text log-probabilities, mu, logvar), the skeleton of
    mnist/model.py:53-84 with the prior expert switched on."""
    mus, lvs = [], []
    if image is not None:
        m, l = image_encoder(p, image)
        mus.append(m); lvs.append(l)
    if text is not None:
        m, l = text_encoder(p, text)
        mus.append(m); lvs.append(l)
    mu, logvar = product_of_experts_precision(torch.stack(mus), torch.stack(lvs), prior=prior, eps=POE_EPS)
    z = reparametrize(mu, logvar, noise, training)
    return torch.sigmoid(image_decoder_logits(p, z)), F.log_softmax(text_decoder_logits(p, z), dim=1), mu, logvar


def elbo_loss(recon_image, image, recon_text, text, mu, logvar, lambda_image=1.0, lambda_text=1.0, annealing_factor=1.0):
    """North-star signature: mean over the batch of lambda_image * BCE(sum over pixels) + lambda_text * CE + annealing * KL."""
    B = mu.shape[0]
    total = mu.new_zeros(())
    parts = [mu.new_zeros(()), mu.new_zeros(())]
    if recon_image is not None and image is not None:
        parts[0] = lambda_image * F.binary_cross_entropy(recon_image, image.reshape(B, 784), reduction="sum") / B
    if recon_text is not None and text is not None:
        parts[1] = lambda_text * F.nll_loss(recon_text, text, reduction="sum") / B
    kl = annealing_factor * (-0.5) * torch.sum(1 + logvar - mu.pow(2) - logvar.exp()) / B
    total = parts[0] + parts[1] + kl
    return total, parts[0], parts[1], kl


TERMS = ("joint", "image", "text")


def elbo_rows(recon_image, image, recon_text, text, mu, logvar, lambda_image=1.0, lambda_text=1.0, annealing_factor=1.0):
    """Per-row parts of elbo_loss (its value is their mean over the batch): (lambda_image * BCE, lambda_text * CE, annealing * KL)."""
    B = mu.shape[0]
    bce = lambda_image * F.binary_cross_entropy(recon_image, image.reshape(B, 784), reduction="none").sum(1)
    ce = lambda_text * F.nll_loss(recon_text, text, reduction="none")
    kl = annealing_factor * (-0.5) * (1 + logvar - mu.pow(2) - logvar.exp()).sum(1)
    return bce, ce, kl


def train_step(p: State, image, text, noises: Sequence[torch.Tensor], terms=TERMS, lambdas=((1.0, 1.0),) * 3,
               annealing_factor=1.0, prior: bool = True, has_image=None, has_text=None):
    """The three-term step of mnist/train.py:132-153 on this model: returns (per-term (total, bce, ce, kl), outputs, grads).
    With per-sample presence masks (has_image / has_text, [B] bool) term g only counts the rows that have the modalities it
    needs (joint: both, image: image, text: text) and its loss is the mean over those rows - mnist/paired_weak.py:82-104 applied
    per row instead of per batch (there is no normalisation layer, so rows are independent)."""
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    B = image.shape[0]
    hi = torch.ones(B, dtype=torch.bool) if has_image is None else has_image.bool()
    ht = torch.ones(B, dtype=torch.bool) if has_text is None else has_text.bool()
    losses, outs = [], []
    total = None
    for g, t in enumerate(terms):
        im = image if t != "text" else None
        tx = text if t != "image" else None
        ri, rt, mu, lv = forward(q, im, tx, noises[g], True, prior)
        on = (hi & ht) if t == "joint" else (hi if t == "image" else ht)
        w = on.float() / max(int(on.sum()), 1)
        parts = [(w * v).sum() for v in elbo_rows(ri, image, rt, text, mu, lv, lambdas[g][0], lambdas[g][1], annealing_factor)]
        l = (parts[0] + parts[1] + parts[2], parts[0], parts[1], parts[2])
        losses.append(tuple(float(v.detach()) for v in l))
        outs.append((ri.detach(), rt.detach(), mu.detach(), lv.detach()))
        total = l[0] if total is None else total + l[0]
    total.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in q.items()}
    return losses, outs, grads


def synthetic_batch(batch: int, n_latents: int, seed: int = 0):
    g = torch.Generator().manual_seed(seed)
    image = torch.rand(batch, 784, generator=g)
    text = torch.randint(0, 10, (batch,), generator=g)
    noises = [torch.randn(batch, n_latents, generator=g) for _ in range(3)]
    return image, text, noises
