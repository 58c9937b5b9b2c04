"""CPU ORACLE (test infrastructure, NOT product code) - MNIST MVAE training step.

A plain restatement, op by op, of the reference's hot path so that the CUDA library can be
checked on a box where /root/reference does not exist.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; the product
(mvae_b200) never does.

Pinned against the real reference: oracle/gen_golden.py imports mnist/model.py + mnist/train.py
from /root/reference, runs them on seeded inputs and writes tests/golden/mnist_*.npz;
tests/test_oracle.py checks this restatement against those fixtures (and the KATs of
SURVEY.md section 4).  The arithmetic substrate is the same as the reference's (PyTorch ATen,
fp32 on CPU); gradients come from torch.autograd exactly as in the reference
(mnist/train.py:147-148).

Every function cites the reference lines it restates.  State is a flat dict keyed by the
reference's own state_dict names (image_encoder.net.0.weight, ...).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

BN_EPS = 1e-5       # nn.BatchNorm1d default (mnist/model.py:105)
BN_MOMENTUM = 0.1   # nn.BatchNorm1d default
POE_EPS = 1e-8      # mnist/model.py:180

State = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------- parameters
def param_shapes(n_latents: int) -> Dict[str, Tuple[int, ...]]:
    """state_dict layout of mnist/model.py:14-170 (MultimodalVAE)."""
    n = n_latents
    s: Dict[str, Tuple[int, ...]] = {}

    def lin(p, o, i):
        s[p + ".weight"] = (o, i)
        s[p + ".bias"] = (o,)

    def bn(p, f):
        s[p + ".weight"] = (f,)
        s[p + ".bias"] = (f,)
        s[p + ".running_mean"] = (f,)
        s[p + ".running_var"] = (f,)
        s[p + ".num_batches_tracked"] = ()

    lin("image_encoder.net.0", 400, 784); bn("image_encoder.net.1", 400)      # model.py:104-105
    lin("image_encoder.net.3", 200, 400); bn("image_encoder.net.4", 200)      # model.py:107-108
    lin("image_encoder.net.6", 2 * n, 200)                                    # model.py:110
    lin("image_decoder.net.0", 200, n); bn("image_decoder.net.1", 200)        # model.py:124-125
    lin("image_decoder.net.3", 400, 200); bn("image_decoder.net.4", 400)      # model.py:127-128
    lin("image_decoder.net.6", 784, 400)                                      # model.py:130
    s["text_encoder.net.0.weight"] = (10, 50)                                 # model.py:143
    bn("text_encoder.net.1", 50)                                              # model.py:144
    lin("text_encoder.net.3", 2 * n, 50)                                      # model.py:146
    lin("text_decoder.net.0", 10, n); bn("text_decoder.net.1", 10)            # model.py:162-163
    lin("text_decoder.net.3", 10, 10)                                         # model.py:165
    return s


# Linear biases that feed a train-mode BatchNorm: their true gradient is exactly zero (BN removes the
# batch mean), the reference holds ~1e-9 rounding noise there.
PRE_BN_BIASES = ("image_encoder.net.0.bias", "image_encoder.net.3.bias", "image_decoder.net.0.bias",
                 "image_decoder.net.3.bias", "text_decoder.net.0.bias")


def is_buffer(name: str) -> bool:
    return name.endswith(("running_mean", "running_var", "num_batches_tracked"))


def init_state(n_latents: int, seed: int = 1234, dtype=torch.float32) -> State:
    """Random state with PyTorch's default initialisers (values need not equal the reference's
    RNG stream - parity tests always copy one state into both sides)."""
    g = torch.Generator().manual_seed(seed)
    st: State = {}
    for k, shp in param_shapes(n_latents).items():
        if k.endswith("num_batches_tracked"):
            st[k] = torch.zeros((), dtype=torch.int64)
        elif k.endswith("running_mean"):
            st[k] = torch.zeros(shp, dtype=dtype)
        elif k.endswith("running_var"):
            st[k] = torch.ones(shp, dtype=dtype)
        elif ".net.1." in k or ".net.4." in k:  # BN affine
            st[k] = torch.ones(shp, dtype=dtype) if k.endswith("weight") else torch.zeros(shp, dtype=dtype)
        elif k == "text_encoder.net.0.weight":
            st[k] = torch.randn(shp, generator=g, dtype=dtype)
        else:
            fan_in = shp[1] if len(shp) == 2 else None
            if fan_in is None:  # bias of the Linear whose weight precedes it
                w = st[k[:-4] + "weight"]
                fan_in = w.shape[1]
            bound = 1.0 / math.sqrt(fan_in)
            st[k] = (torch.rand(shp, generator=g, dtype=dtype) * 2 - 1) * bound
    return st


def perturbed_state(n_latents: int, seed: int) -> State:
    """init_state with non-trivial BatchNorm affine parameters (used by the golden fixtures)."""
    st = init_state(n_latents, seed=1234 + seed)
    g = torch.Generator().manual_seed(seed + 77)
    for k in st:
        if (".net.1." in k or ".net.4." in k) and k.endswith("weight"):
            st[k] = 1.0 + 0.2 * torch.randn(st[k].shape, generator=g)
        if (".net.1." in k or ".net.4." in k) and k.endswith("bias"):
            st[k] = 0.1 * torch.randn(st[k].shape, generator=g)
    return st


def randomize_running_stats(state: State, seed: int) -> State:
    """Same weights, running_mean ~ 0.2 N(0,1), running_var ~ U(0.5, 1.5), deterministic in `seed` and the key order:
    the state of the eval-mode fixtures (oracle/gen_golden_eval.py).  Works for any of the three families' states."""
    g = torch.Generator().manual_seed(9000 + seed)
    out = {}
    for k, v in state.items():
        if k.endswith("running_mean"):
            out[k] = 0.2 * torch.randn(v.shape, generator=g)
        elif k.endswith("running_var"):
            out[k] = 0.5 + torch.rand(v.shape, generator=g)
        else:
            out[k] = v.clone()
    return out


def sample_flat(t: torch.Tensor, max_n: int = 512) -> torch.Tensor:
    """Deterministic strided sample of a tensor (fixtures store samples, not whole gradients)."""
    f = t.detach().reshape(-1)
    stride = max(1, f.numel() // max_n)
    return f[::stride].contiguous()


# ----------------------------------------------------------------------------- layers
# Optional emulation of the tensor-core operand rounding of the CUDA path (None = exact fp32 like the
# reference).  With "tf32" every GEMM operand - forward, dgrad and wgrad - is rounded to a 10-bit mantissa
# (round-to-nearest-even) before an exact product, which is what TMA + tcgen05 kind::tf32 do; the parity tests
# use it to separate LOGIC errors (must vanish against this oracle) from the intrinsic tf32 rounding noise
# (visible against the exact oracle, amplified by the mean-subtraction of BatchNorm's backward).
MATMUL_EMULATION: Optional[str] = None


def round_tf32(x: torch.Tensor) -> torch.Tensor:
    if x.dtype != torch.float32:
        return x
    i = x.contiguous().view(torch.int32)
    r = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
    return r.view(torch.float32)


def _round_operand(x):
    if MATMUL_EMULATION == "tf32":
        return round_tf32(x)
    if MATMUL_EMULATION == "bf16":
        return x.to(torch.bfloat16).to(x.dtype)
    return x


class _EmulatedMatmul(torch.autograd.Function):
    """y = x W^T with operand rounding in all three GEMMs (forward, dgrad, wgrad)."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return _round_operand(x) @ _round_operand(w).t()

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dyr = _round_operand(dy)
        return dyr @ _round_operand(w), dyr.t() @ _round_operand(x)


# Linear layers the CUDA path runs on tensor cores (operand rounding applies); the text encoder / decoder
# Linears run in plain fp32 inside the tail kernels.
TENSOR_CORE_LINEARS = ("image_encoder.net.0", "image_encoder.net.3", "image_encoder.net.6", "image_decoder.net.0",
                       "image_decoder.net.3", "image_decoder.net.6")

# Forward substitution for the backward-consistency test: {(tag, term): tensor}.  When set, the value of the
# tagged forward tensor is REPLACED by the given one (straight-through: gradients flow as if it had been
# computed here).  Gradients of a ReLU network are discontinuous in the forward activations - two valid tf32
# evaluations that differ by 1e-4 flip a handful of ReLU units and move inner-layer gradients by ~1 % - so the
# tight logic check feeds the device's own forward tensors in and compares the backward given THOSE.
FORWARD_OVERRIDE: Optional[Dict[Tuple[str, int], torch.Tensor]] = None
CURRENT_TERM = 0


def _override(h, tag):
    if FORWARD_OVERRIDE is None:
        return h
    v = FORWARD_OVERRIDE.get((tag, CURRENT_TERM))
    if v is None:
        return h
    return h + (v.to(h.dtype) - h).detach()


def linear(x, w, b=None, tag: Optional[str] = None):
    """nn.Linear: y = x W^T + b."""
    emulate = MATMUL_EMULATION is not None and (tag is None or tag in TENSOR_CORE_LINEARS)
    y = _EmulatedMatmul.apply(x, w) if emulate else x @ w.t()
    y = y if b is None else y + b
    return _override(y, tag) if tag is not None else y


def batchnorm_train(x, gamma, beta, st: Optional[State] = None, prefix: str = ""):
    """nn.BatchNorm1d in train mode: biased batch variance for normalisation; running stats
    updated with momentum 0.1 and the UNBIASED variance; num_batches_tracked += 1."""
    mean = x.mean(0)
    var = x.var(0, unbiased=False)
    y = (x - mean) / torch.sqrt(var + BN_EPS) * gamma + beta
    if st is not None:
        n = x.shape[0]
        with torch.no_grad():
            unb = var * (n / (n - 1)) if n > 1 else var
            st[prefix + ".running_mean"] = (1 - BN_MOMENTUM) * st[prefix + ".running_mean"] + BN_MOMENTUM * mean.detach()
            st[prefix + ".running_var"] = (1 - BN_MOMENTUM) * st[prefix + ".running_var"] + BN_MOMENTUM * unb.detach()
            st[prefix + ".num_batches_tracked"] = st[prefix + ".num_batches_tracked"] + 1
    return y


def batchnorm_eval(x, gamma, beta, st: State, prefix: str):
    return (x - st[prefix + ".running_mean"]) / torch.sqrt(st[prefix + ".running_var"] + BN_EPS) * gamma + beta


def _bn(x, p: State, st: Optional[State], prefix: str, training: bool):
    if training:
        return batchnorm_train(x, p[prefix + ".weight"], p[prefix + ".bias"], st, prefix)
    return batchnorm_eval(x, p[prefix + ".weight"], p[prefix + ".bias"], st if st is not None else p, prefix)


def image_encoder(p: State, x, st=None, training=True):
    """mnist/model.py:99-117."""
    h = linear(x, p["image_encoder.net.0.weight"], p["image_encoder.net.0.bias"], "image_encoder.net.0")
    h = torch.relu(_bn(h, p, st, "image_encoder.net.1", training))
    h = linear(h, p["image_encoder.net.3.weight"], p["image_encoder.net.3.bias"], "image_encoder.net.3")
    h = torch.relu(_bn(h, p, st, "image_encoder.net.4", training))
    h = linear(h, p["image_encoder.net.6.weight"], p["image_encoder.net.6.bias"], "image_encoder.net.6")
    n = h.shape[1] // 2
    return h[:, :n], h[:, n:]


def text_encoder(p: State, text, st=None, training=True):
    """mnist/model.py:138-153."""
    h = p["text_encoder.net.0.weight"][text]
    h = torch.relu(_bn(h, p, st, "text_encoder.net.1", training))
    h = linear(h, p["text_encoder.net.3.weight"], p["text_encoder.net.3.bias"], "text_encoder.net.3")
    n = h.shape[1] // 2
    return h[:, :n], h[:, n:]


def image_decoder_logits(p: State, z, st=None, training=True):
    """mnist/model.py:120-134 (everything before the sigmoid of :135)."""
    h = linear(z, p["image_decoder.net.0.weight"], p["image_decoder.net.0.bias"], "image_decoder.net.0")
    h = torch.relu(_bn(h, p, st, "image_decoder.net.1", training))
    h = linear(h, p["image_decoder.net.3.weight"], p["image_decoder.net.3.bias"], "image_decoder.net.3")
    h = torch.relu(_bn(h, p, st, "image_decoder.net.4", training))
    return linear(h, p["image_decoder.net.6.weight"], p["image_decoder.net.6.bias"], "image_decoder.net.6")


def text_decoder_logits(p: State, z, st=None, training=True):
    """mnist/model.py:156-169 (everything before the log_softmax of :170)."""
    h = linear(z, p["text_decoder.net.0.weight"], p["text_decoder.net.0.bias"], "text_decoder.net.0")
    h = torch.relu(_bn(h, p, st, "text_decoder.net.1", training))
    return linear(h, p["text_decoder.net.3.weight"], p["text_decoder.net.3.bias"], "text_decoder.net.3")


def product_of_experts(mu, logvar, eps: float = POE_EPS):
    """mnist/model.py:180-185 - NB: mu is VARIANCE-weighted in the reference."""
    var = torch.exp(logvar) + eps
    pd_mu = torch.sum(mu * var, dim=0) / torch.sum(var, dim=0)
    pd_var = 1.0 / torch.sum(1.0 / var, dim=0)
    return pd_mu, torch.log(pd_var)


def product_of_experts_precision(mu, logvar, mask=None, prior: bool = False, eps: float = POE_EPS):
    """The paper's precision-weighted product (paper/draft.tex:88), with an optional per-sample
    presence mask [M, B] and an optional N(0,1) prior expert (the commented-out lines
    mnist/model.py:70-75).  Not used by the reference code; this is the north-star variant."""
    var = torch.exp(logvar) + eps
    T = 1.0 / var
    if mask is not None:
        T = T * mask.unsqueeze(-1)
    num = (mu * T).sum(0)
    den = T.sum(0)
    if prior:
        den = den + 1.0  # prior: mu 0, precision 1
    pd_mu = num / den
    return pd_mu, torch.log(1.0 / den)


def reparametrize(mu, logvar, noise=None, training=True):
    """mnist/model.py:24-30, with the N(0,1) draw injectable."""
    if not training:
        return mu
    std = torch.exp(0.5 * logvar)
    if noise is None:
        noise = torch.randn_like(std)
    return noise * std + mu


# PoE arithmetic used by forward(): None = the reference's (mnist/model.py:180-185, variance-weighted mu, no prior);
# ("precision", prior) = the paper's precision-weighted product with an optional N(0,1) prior expert - the
# north-star variant.  The reference does not implement it ("parity unpinned": pinned only to the formula of
# paper/draft.tex:88 restated in product_of_experts_precision).
POE_VARIANT = None


def forward(p: State, image=None, text=None, noise=None, st=None, training=True):
    """MultimodalVAE.forward, mnist/model.py:53-84.  Returns (recon_image probs, recon_text
    log-probs, mu, logvar) like the reference, plus the decoder logits."""
    assert image is not None or text is not None
    mus, lvs = [], []
    if image is not None:
        m, l = image_encoder(p, image, st, training)
        mus.append(m); lvs.append(l)
    if text is not None:
        m, l = text_encoder(p, text, st, training)
        mus.append(m); lvs.append(l)
    if POE_VARIANT is None:
        mu, logvar = product_of_experts(torch.stack(mus, 0), torch.stack(lvs, 0))
    else:
        mu, logvar = product_of_experts_precision(torch.stack(mus, 0), torch.stack(lvs, 0), prior=bool(POE_VARIANT[1]))
    z = _override(reparametrize(mu, logvar, noise, training), "z")
    img_logits = image_decoder_logits(p, z, st, training)
    txt_logits = text_decoder_logits(p, z, st, training)
    return torch.sigmoid(img_logits), torch.log_softmax(txt_logits, dim=1), mu, logvar, img_logits, txt_logits


def binary_cross_entropy_mean(prob, target):
    """F.binary_cross_entropy (mean) incl. ATen's clamp of the logs at -100 (mnist/train.py:70)."""
    l1 = torch.clamp(torch.log(prob), min=-100.0)
    l0 = torch.clamp(torch.log(1.0 - prob), min=-100.0)
    return -(target * l1 + (1.0 - target) * l0).mean()


def loss_function(mu, logvar, recon_image=None, image=None, recon_text=None, text=None,
                  lambda_xy=1.0, lambda_yx=1.0):
    """mnist/train.py:64-81."""
    image_bce = 0.0
    text_bce = 0.0
    batch = mu.shape[0]
    if recon_image is not None and image is not None:
        image_bce = lambda_xy * binary_cross_entropy_mean(recon_image, image.reshape(-1, 784))
    if recon_text is not None and text is not None:
        text_bce = lambda_yx * (-recon_text.gather(1, text.view(-1, 1)).mean())
    kld = -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp())
    kld = kld / (batch * (784 / 3))
    return image_bce + text_bce + kld


# ----------------------------------------------------------------------------- the step
TERMS = ("joint", "image", "text")
DEFAULT_LAMBDAS = ((1.0, 1.0), (1.0, 1.0), (1.0, 1.0))  # mnist/train.py:140-145


def train_step_losses(p: State, image, text, noises: Sequence[torch.Tensor], st: Optional[State] = None,
                      lambdas=DEFAULT_LAMBDAS, terms=(True, True, True)):
    """mnist/train.py:136-146: three forwards (joint, image-only, text-only) and their losses.
    `terms` switches terms off (mnist/modal_weak.py:76-97)."""
    losses: List[torch.Tensor] = []
    outs = []
    args = ((image, text), (image, None), (None, text))
    global CURRENT_TERM
    for k in range(3):
        if not terms[k]:
            losses.append(torch.zeros((), dtype=image.dtype))
            outs.append(None)
            continue
        CURRENT_TERM = k
        ri, rt, mu, lv, il, tl = forward(p, args[k][0], args[k][1], noises[k], st, True)
        losses.append(loss_function(mu, lv, ri, image, rt, text, lambdas[k][0], lambdas[k][1]))
        outs.append((ri, rt, mu, lv, il, tl))
    return losses, outs


def train_step(p: State, image, text, noises, lambdas=DEFAULT_LAMBDAS, terms=(True, True, True)):
    """zero_grad + 3 forwards + 3 losses + backward (mnist/train.py:132-148).
    Returns (losses[3], grads dict, new buffer state, outs).  `p` is not modified."""
    work: State = {}
    for k, v in p.items():
        if is_buffer(k):
            work[k] = v.clone()
        else:
            work[k] = v.detach().clone().requires_grad_(True)
    losses, outs = train_step_losses(work, image, text, noises, work, lambdas, terms)
    total = losses[0] + losses[1] + losses[2]
    names = [k for k in work if not is_buffer(k)]
    gs = torch.autograd.grad(total, [work[k] for k in names], allow_unused=True)
    grads = {k: (torch.zeros_like(work[k]) if g is None else g) for k, g in zip(names, gs)}
    buffers = {k: v for k, v in work.items() if is_buffer(k)}
    return [float(l.detach()) for l in losses], grads, buffers, outs


def adam_step(p: State, grads: State, m: State, v: State, step: int, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam defaults (mnist/train.py:118,153); step is 1-based."""
    out = {}
    for k, w in p.items():
        if is_buffer(k):
            out[k] = w
            continue
        g = grads[k]
        m[k] = b1 * m[k] + (1 - b1) * g
        v[k] = b2 * v[k] + (1 - b2) * g * g
        bc1 = 1 - b1 ** step
        bc2 = 1 - b2 ** step
        denom = v[k].sqrt() / math.sqrt(bc2) + eps
        out[k] = w - (lr / bc1) * m[k] / denom
    return out


def synthetic_batch(batch: int, n_latents: int, seed: int = 0, dtype=torch.float32):
    """Inputs of SURVEY.md section 8d: image U[0,1), labels, three N(0,1) draws."""
    g = torch.Generator().manual_seed(seed)
    image = torch.rand(batch, 784, generator=g, dtype=dtype)
    text = torch.randint(0, 10, (batch,), generator=g)
    noises = [torch.randn(batch, n_latents, generator=g, dtype=dtype) for _ in range(3)]
    return image, text, noises


def test_mnist(p: State, batches) -> float:
    """mnist/test.py:18-36: fraction of labels recovered from the image alone, eval mode (argmax of recon_text)."""
    hits, total = 0, 0
    for image, text in batches:
        out = forward(p, image.reshape(image.shape[0], -1), None, None, None, training=False)
        hits += int((out[1].argmax(1) == text).sum())
        total += int(text.numel())
    return hits / float(total)


def compute_nll(p: State, batches, image_only=False, text_only=False, n_samples=1, generator=None):
    """mnist/loglikelihood.py:15-63: per-example (image NLL, text NLL) with z ~ q(z | inputs); ONE [n_samples, n_latents]
    normal tensor per batch, shared by its rows (:36-46), sums over pixels / classes (size_average=False, :51-52)."""
    assert not (image_only and text_only)
    tot_i = tot_t = 0.0
    total = 0
    for image, text in batches:
        image = image.reshape(image.shape[0], -1)
        out = forward(p, None if text_only else image, None if image_only else text, None, None, training=False)
        mu, logvar = out[2], out[3]
        sample = torch.randn(n_samples, mu.shape[1], generator=generator)
        std = torch.exp(0.5 * logvar)
        for i in range(n_samples):
            z = sample[i].unsqueeze(0) * std + mu
            ri = torch.sigmoid(image_decoder_logits(p, z, None, False))
            rt = torch.log_softmax(text_decoder_logits(p, z, None, False), dim=1)
            tot_i += float(torch.nn.functional.binary_cross_entropy(ri, image, reduction="sum")) / n_samples
            tot_t += float(torch.nn.functional.nll_loss(rt, text, reduction="sum")) / n_samples
        total += image.shape[0]
    return tot_i / total, tot_t / total
