"""CPU ORACLE (test infrastructure, NOT product code) - CelebA MVAE training step.

Restates celeba/model.py:13-241 and celeba/train.py:60-81,132-157 op by op in functional form over a
flat state dict keyed by the reference's own state_dict names.  Same rules as mnist_oracle.py: only
tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it; the product never does.

Pinned against the real reference: oracle/gen_golden_celeba.py imports celeba/model.py + the
loss_function of celeba/train.py from /root/reference (Dropout.p = 0, injected reparametrize noise) and
writes tests/golden/celeba_*.npz; tests/test_oracle_celeba.py checks this restatement against them.
The arithmetic substrate is PyTorch ATen fp32 on CPU (as the reference); gradients come from
torch.autograd like celeba/train.py:152.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
POE_EPS = 1e-8        # celeba/model.py:231
N_ATTRS = 18          # celeba/datasets.py:27
KL_LAMBDA = 1e-3      # celeba/train.py:61
DROPOUT_P = 0.1       # celeba/model.py:117

State = Dict[str, torch.Tensor]

# (key prefix, kind, shape builder) in state_dict order
ENC_CONVS = (("image_encoder.features.0", 3, 32, 2, 1), ("image_encoder.features.2", 32, 64, 2, 1),
             ("image_encoder.features.5", 64, 128, 2, 1), ("image_encoder.features.8", 128, 256, 1, 0))
ENC_BNS = (None, "image_encoder.features.3", "image_encoder.features.6", "image_encoder.features.9")
DEC_CONVS = (("image_decoder.hallucinate.0", 256, 128, 1, 0), ("image_decoder.hallucinate.3", 128, 64, 2, 1),
             ("image_decoder.hallucinate.6", 64, 32, 2, 1), ("image_decoder.hallucinate.9", 32, 3, 2, 1))
DEC_BNS = ("image_decoder.hallucinate.1", "image_decoder.hallucinate.4", "image_decoder.hallucinate.7", None)


def param_shapes(n_latents: int) -> Dict[str, Tuple[int, ...]]:
    """state_dict layout of celeba/model.py:13-21 (MultimodalVAE), in registration order."""
    n = n_latents
    s: Dict[str, Tuple[int, ...]] = {}

    def bn(p, c):
        s[p + ".weight"] = (c,)
        s[p + ".bias"] = (c,)
        s[p + ".running_mean"] = (c,)
        s[p + ".running_var"] = (c,)
        s[p + ".num_batches_tracked"] = ()

    def lin(p, o, i):
        s[p + ".weight"] = (o, i)
        s[p + ".bias"] = (o,)

    for (p, ci, co, _, _), b in zip(ENC_CONVS, ENC_BNS):   # celeba/model.py:101-113 (bias=False)
        s[p + ".weight"] = (co, ci, 4, 4)
        if b:
            bn(b, co)
    lin("image_encoder.classifier.0", 1024, 256 * 5 * 5)   # :115
    lin("image_encoder.classifier.3", 2 * n, 1024)         # :118
    lin("image_decoder.upsample.0", 256 * 5 * 5, n)        # :138
    for (p, ci, co, _, _), b in zip(DEC_CONVS, DEC_BNS):   # :142-152 ConvTranspose2d weight [in, out, k, k]
        s[p + ".weight"] = (ci, co, 4, 4)
        if b:
            bn(b, co)
    lin("attrs_encoder.net.0", 64, N_ATTRS)                # :171
    bn("attrs_encoder.net.1", 64)
    lin("attrs_encoder.net.3", 2 * n, 64)
    lin("attrs_decoder.net.0", 64, n)                      # :188
    bn("attrs_decoder.net.1", 64)
    lin("attrs_decoder.net.3", N_ATTRS, 64)
    return s


def is_buffer(name: str) -> bool:
    return name.endswith(("running_mean", "running_var", "num_batches_tracked"))


def _is_bn(name: str) -> bool:
    p = name.rsplit(".", 1)[0]
    return p in ENC_BNS or p in DEC_BNS or p in ("attrs_encoder.net.1", "attrs_decoder.net.1")


def init_state(n_latents: int, seed: int = 1234, perturb_bn: bool = True, dtype=torch.float32) -> State:
    """Random state with PyTorch-default-like initialisers (parity tests copy one state into both sides, so
    the values need not equal the reference's RNG stream; celeba/train.py:118 weight_init is a no-op)."""
    g = torch.Generator().manual_seed(seed)
    st: State = {}
    for k, shp in param_shapes(n_latents).items():
        if k.endswith("num_batches_tracked"):
            st[k] = torch.zeros((), dtype=torch.int64)
        elif k.endswith("running_mean"):
            st[k] = torch.zeros(shp, dtype=dtype)
        elif k.endswith("running_var"):
            st[k] = torch.ones(shp, dtype=dtype)
        elif _is_bn(k):
            if k.endswith("weight"):
                st[k] = 1.0 + (0.2 * torch.randn(shp, generator=g, dtype=dtype) if perturb_bn else 0.0) * torch.ones(shp)
            else:
                st[k] = (0.1 * torch.randn(shp, generator=g, dtype=dtype)) if perturb_bn else torch.zeros(shp, dtype=dtype)
        elif len(shp) == 4:
            fan_in = shp[1] * 16 if "encoder" in k else shp[0] * 16
            bound = 1.0 / math.sqrt(fan_in)
            st[k] = (torch.rand(shp, generator=g, dtype=dtype) * 2 - 1) * bound
        elif len(shp) == 2:
            bound = 1.0 / math.sqrt(shp[1])
            st[k] = (torch.rand(shp, generator=g, dtype=dtype) * 2 - 1) * bound
        else:
            w = st[k[:-4] + "weight"]
            bound = 1.0 / math.sqrt(w.shape[1])
            st[k] = (torch.rand(shp, generator=g, dtype=dtype) * 2 - 1) * bound
    return st


def sample_flat(t: torch.Tensor, max_n: int = 512) -> torch.Tensor:
    f = t.detach().reshape(-1)
    stride = max(1, f.numel() // max_n)
    return f[::stride].contiguous()


# ----------------------------------------------------------------------------- layers
def swish(x):
    """celeba/model.py:235-241."""
    return x * torch.sigmoid(x)


def batchnorm(x, p: State, st: Optional[State], prefix: str, training: bool):
    """nn.BatchNorm1d / nn.BatchNorm2d (channel axis 1): batch statistics over every other axis in train
    mode (biased variance normalises, unbiased one feeds running_var), running statistics in eval mode."""
    axes = [0] + list(range(2, x.dim()))
    shape = [1, -1] + [1] * (x.dim() - 2)
    g, b = p[prefix + ".weight"].view(shape), p[prefix + ".bias"].view(shape)
    if not training:
        src = st if st is not None else p
        return (x - src[prefix + ".running_mean"].view(shape)) / torch.sqrt(src[prefix + ".running_var"].view(shape) + BN_EPS) * g + b
    mean = x.mean(axes)
    var = x.var(axes, unbiased=False)
    y = (x - mean.view(shape)) / torch.sqrt(var.view(shape) + BN_EPS) * g + b
    if st is not None:
        n = x.numel() // x.shape[1]
        with torch.no_grad():
            unb = var * (n / (n - 1)) if n > 1 else var
            st[prefix + ".running_mean"] = (1 - BN_MOMENTUM) * st[prefix + ".running_mean"] + BN_MOMENTUM * mean.detach()
            st[prefix + ".running_var"] = (1 - BN_MOMENTUM) * st[prefix + ".running_var"] + BN_MOMENTUM * unb.detach()
            st[prefix + ".num_batches_tracked"] = st[prefix + ".num_batches_tracked"] + 1
    return y


def image_encoder(p: State, x, st=None, training=True, drop_mask=None):
    """celeba/model.py:99-131.  drop_mask: the [B,1024] keep-mask already scaled by 1/(1-p) (None = no dropout,
    which is what the parity fixtures use: Dropout.p = 0)."""
    h = x
    for (pre, _, _, stride, pad), bn in zip(ENC_CONVS, ENC_BNS):
        h = F.conv2d(h, p[pre + ".weight"], None, stride, pad)
        if bn:
            h = batchnorm(h, p, st, bn, training)
        h = swish(h)
    h = h.reshape(-1, 256 * 5 * 5)
    h = swish(F.linear(h, p["image_encoder.classifier.0.weight"], p["image_encoder.classifier.0.bias"]))
    if drop_mask is not None and training:
        h = h * drop_mask
    h = F.linear(h, p["image_encoder.classifier.3.weight"], p["image_encoder.classifier.3.bias"])
    n = h.shape[1] // 2
    return h[:, :n], h[:, n:]


def image_decoder_logits(p: State, z, st=None, training=True):
    """celeba/model.py:134-163 up to (excluding) the final sigmoid."""
    h = swish(F.linear(z, p["image_decoder.upsample.0.weight"], p["image_decoder.upsample.0.bias"]))
    h = h.view(-1, 256, 5, 5)
    for (pre, _, _, stride, pad), bn in zip(DEC_CONVS, DEC_BNS):
        h = F.conv_transpose2d(h, p[pre + ".weight"], None, stride, pad)
        if bn:
            h = swish(batchnorm(h, p, st, bn, training))
    return h


def attrs_encoder(p: State, a, st=None, training=True):
    """celeba/model.py:166-182."""
    h = F.linear(a, p["attrs_encoder.net.0.weight"], p["attrs_encoder.net.0.bias"])
    h = swish(batchnorm(h, p, st, "attrs_encoder.net.1", training))
    h = F.linear(h, p["attrs_encoder.net.3.weight"], p["attrs_encoder.net.3.bias"])
    n = h.shape[1] // 2
    return h[:, :n], h[:, n:]


def attrs_decoder_logits(p: State, z, st=None, training=True):
    """celeba/model.py:185-200 up to the final sigmoid."""
    h = F.linear(z, p["attrs_decoder.net.0.weight"], p["attrs_decoder.net.0.bias"])
    h = swish(batchnorm(h, p, st, "attrs_decoder.net.1", training))
    return F.linear(h, p["attrs_decoder.net.3.weight"], p["attrs_decoder.net.3.bias"])


def product_of_experts(mu, logvar, eps: float = POE_EPS):
    """celeba/model.py:230-235 (variance-weighted mean, as written in the reference)."""
    var = torch.exp(logvar) + eps
    pd_mu = torch.sum(mu * var, dim=0) / torch.sum(var, dim=0)
    pd_var = 1.0 / torch.sum(1.0 / var, dim=0)
    return pd_mu, torch.log(pd_var)


def forward(p: State, image=None, attrs=None, noise=None, st=None, training=True, drop_mask=None):
    """celeba/model.py:36-58.  Returns (image_recon, attrs_recon, mu, logvar, image_logits, attrs_logits)."""
    assert image is not None or attrs is not None
    mus, lvs = [], []
    if image is not None:
        m, l = image_encoder(p, image, st, training, drop_mask)
        mus.append(m); lvs.append(l)
    if attrs is not None:
        m, l = attrs_encoder(p, attrs, st, training)
        mus.append(m); lvs.append(l)
    mu, logvar = product_of_experts(torch.stack(mus, 0), torch.stack(lvs, 0))
    z = mu + noise * torch.exp(0.5 * logvar) if training else mu   # celeba/model.py:27-34
    il = image_decoder_logits(p, z, st, training)
    al = attrs_decoder_logits(p, z, st, training)
    return torch.sigmoid(il), torch.sigmoid(al), mu, logvar, il, al


def binary_cross_entropy_mean(prob, target):
    """F.binary_cross_entropy (mean): log terms clamped at -100 as ATen does."""
    lp = torch.clamp(torch.log(prob), min=-100.0)
    l1p = torch.clamp(torch.log(1.0 - prob), min=-100.0)
    return -(target * lp + (1.0 - target) * l1p).mean()


def loss_function(mu, logvar, recon_x=None, x=None, recon_y=None, y=None, kl_lambda=KL_LAMBDA, lambda_x=1.0, lambda_y=1.0):
    """celeba/train.py:60-81."""
    B = mu.shape[0]
    x_bce = 0.0
    y_bce = 0.0
    if recon_x is not None and x is not None:
        x_bce = binary_cross_entropy_mean(recon_x.reshape(-1, 3 * 64 * 64), x.reshape(-1, 3 * 64 * 64))
    if recon_y is not None and y is not None:
        y_bce = 0.0
        for i in range(y.shape[1]):
            y_bce = y_bce + binary_cross_entropy_mean(recon_y[:, i], y[:, i])
        y_bce = y_bce / y.shape[1]
    kld = -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp())
    return lambda_x * x_bce + lambda_y * y_bce + kld / B * kl_lambda


def train_step(p: State, image, attrs, noises: Sequence[torch.Tensor], drop_masks=(None, None)):
    """celeba/train.py:138-152: zero_grad, vae(image, attrs), vae(image), vae(attrs), three loss_function calls
    (each against BOTH targets, defaults), summed, backward.  drop_masks: keep-masks of the joint / image-only
    image-encoder passes.  Returns (losses[3], grads, new buffers, outs)."""
    work: State = {}
    for k, v in p.items():
        work[k] = v.clone() if is_buffer(k) else v.detach().clone().requires_grad_(True)
    args = ((image, attrs, drop_masks[0]), (image, None, drop_masks[1]), (None, attrs, None))
    losses, outs = [], []
    for k in range(3):
        ri, ra, mu, lv, il, al = forward(work, args[k][0], args[k][1], noises[k], work, True, args[k][2])
        losses.append(loss_function(mu, lv, ri, image, ra, attrs))
        outs.append((ri, ra, mu, lv, il, al))
    total = losses[0] + losses[1] + losses[2]
    names = [k for k in work if not is_buffer(k)]
    gs = torch.autograd.grad(total, [work[k] for k in names], allow_unused=True)
    grads = {k: (torch.zeros_like(work[k]) if g is None else g) for k, g in zip(names, gs)}
    buffers = {k: v for k, v in work.items() if is_buffer(k)}
    return [float(l.detach()) for l in losses], grads, buffers, outs


def synthetic_batch(batch: int, n_latents: int, seed: int = 0, dtype=torch.float32):
    """Inputs of SURVEY.md section 8d config 4: image U[0,1) [B,3,64,64], attrs in {0,1} [B,18], three N(0,1) draws."""
    g = torch.Generator().manual_seed(seed)
    image = torch.rand(batch, 3, 64, 64, generator=g, dtype=dtype)
    attrs = (torch.rand(batch, N_ATTRS, generator=g, dtype=dtype) > 0.5).to(dtype)
    noises = [torch.randn(batch, n_latents, generator=g, dtype=dtype) for _ in range(3)]
    return image, attrs, noises
