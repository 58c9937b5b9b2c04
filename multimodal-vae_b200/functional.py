"""torch.autograd.Function wrappers over the C ABI: ProductOfExperts (with a missing-modality mask), the ELBO
loss on module outputs (reference signature `loss_function` and north-star signature `elbo_loss`).

Reference: ProductOfExperts mnist/model.py:173-185; loss_function mnist/train.py:64-81.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn as nn

from . import _lib

_MODES = {"ref": _lib.POE_REF, "precision": _lib.POE_PRECISION}


def _sp():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("mvae_b200 ops need CUDA tensors (there is no CPU path)")


class _PoEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, logvar, mask, mode, prior, eps):
        _need_cuda(mu, logvar, mask)
        mu_c = mu.contiguous().float()
        lv_c = logvar.contiguous().float()
        M = mu_c.shape[0]
        D = mu_c.shape[-1]
        B = mu_c[0].numel() // D
        mk = None
        if mask is not None:
            mk = mask.to(mu_c.device, torch.float32).reshape(M, B).contiguous()
        out_mu = torch.empty(mu_c.shape[1:], device=mu_c.device, dtype=torch.float32)
        out_lv = torch.empty_like(out_mu)
        _lib.check(_lib.load().mvae_poe_forward(mode, int(prior), float(eps), M, B, D, mu_c.data_ptr(), lv_c.data_ptr(),
                                                _lib.ptr(mk), out_mu.data_ptr(), out_lv.data_ptr(), _sp()),
                   "mvae_poe_forward")
        ctx.save_for_backward(mu_c, lv_c, mk)
        ctx.cfg = (mode, int(prior), float(eps), M, B, D)
        return out_mu, out_lv

    @staticmethod
    def backward(ctx, g_mu, g_lv):
        mu_c, lv_c, mk = ctx.saved_tensors
        mode, prior, eps, M, B, D = ctx.cfg
        g_mu = None if g_mu is None else g_mu.contiguous().float()
        g_lv = None if g_lv is None else g_lv.contiguous().float()
        d_mu = torch.empty_like(mu_c)
        d_lv = torch.empty_like(lv_c)
        _lib.check(_lib.load().mvae_poe_backward(mode, prior, eps, M, B, D, mu_c.data_ptr(), lv_c.data_ptr(),
                                                 _lib.ptr(mk), _lib.ptr(g_mu), _lib.ptr(g_lv), d_mu.data_ptr(),
                                                 d_lv.data_ptr(), _sp()), "mvae_poe_backward")
        return d_mu, d_lv, None, None, None, None


class ProductOfExperts(nn.Module):
    """Product of independent Gaussian experts.

    `ProductOfExperts()(mu, logvar)` is the reference's module (mnist/model.py:173-185): mu/logvar are
    [M, B, D] stacks, the result [B, D].  `mask` ([M, B], 1 = expert present for that sample) is the north-star
    extension for per-sample missing modalities.  mode="ref" reproduces the reference arithmetic exactly
    (variance-weighted mean, +eps, no prior); mode="precision" is the paper's precision-weighted product, with an
    optional N(0, 1) prior expert.
    """

    def __init__(self, mode: str = "ref", prior_expert: bool = False):
        super().__init__()
        self.mode = _MODES[mode]
        self.prior_expert = bool(prior_expert)

    def forward(self, mu, logvar, mask=None, eps: float = 1e-8):
        return _PoEFn.apply(mu, logvar, mask, self.mode, self.prior_expert, eps)


class _ElboFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, logvar, recon_image, image, recon_text, text, lam_img, lam_txt, kl_weight):
        _need_cuda(mu, logvar, recon_image, image, recon_text, text)
        mu_c, lv_c = mu.contiguous().float(), logvar.contiguous().float()
        B, n = mu_c.shape
        a = _lib.ElboLossArgs()
        a.batch, a.n_latents = B, n
        a.mu, a.logvar = mu_c.data_ptr(), lv_c.data_ptr()
        a.lambda_image, a.lambda_text, a.kl_weight = float(lam_img), float(lam_txt), float(kl_weight)
        keep = [mu_c, lv_c]
        ri = img = rt = tx = None
        if recon_image is not None and image is not None:
            ri = recon_image.reshape(B, -1).contiguous()
            if ri.dtype not in (torch.float32, torch.bfloat16):
                ri = ri.float()
            img = image.reshape(B, -1).to(ri.dtype).contiguous()
            a.image_dtype = _lib.DT_F32 if ri.dtype == torch.float32 else _lib.DT_BF16
            a.n_pixels = ri.shape[1]
            a.recon_image, a.image = ri.data_ptr(), img.data_ptr()
        if recon_text is not None and text is not None:
            rt = recon_text.contiguous().float()
            tx = text.to(mu_c.device).long().contiguous()
            a.n_classes = rt.shape[1]
            a.recon_text, a.text = rt.data_ptr(), tx.data_ptr()
        out = torch.empty(4, device=mu_c.device, dtype=torch.float32)
        _lib.check(_lib.load().mvae_elbo_loss_forward(C.byref(a), C.c_void_p(out.data_ptr()), _sp()),
                   "mvae_elbo_loss_forward")
        ctx.args = a
        ctx.keep = (mu_c, lv_c, ri, img, rt, tx)
        ctx.in_dtypes = (mu.dtype, logvar.dtype, None if recon_image is None else recon_image.dtype,
                         None if recon_text is None else recon_text.dtype)
        ctx.shapes = (None if recon_image is None else recon_image.shape,)
        ctx.mark_non_differentiable()
        ctx.parts = out
        return out[0]

    @staticmethod
    def backward(ctx, g):
        a = ctx.args
        mu_c, lv_c, ri, img, rt, tx = ctx.keep
        g = g.contiguous().float()
        d_ri = torch.empty_like(ri) if ri is not None else None
        d_rt = torch.empty_like(rt) if rt is not None else None
        d_mu, d_lv = torch.empty_like(mu_c), torch.empty_like(lv_c)
        _lib.check(_lib.load().mvae_elbo_loss_backward(
            C.byref(a), C.c_void_p(g.data_ptr()), C.c_void_p(_lib.ptr(d_ri)), C.c_void_p(_lib.ptr(d_rt)),
            C.c_void_p(d_mu.data_ptr()), C.c_void_p(d_lv.data_ptr()), _sp()), "mvae_elbo_loss_backward")
        if d_ri is not None:
            d_ri = d_ri.view(ctx.shapes[0]).to(ctx.in_dtypes[2])
        return d_mu, d_lv, d_ri, None, d_rt, None, None, None, None


def loss_function(mu, logvar, recon_image=None, image=None, recon_text=None, text=None, lambda_xy=1.0, lambda_yx=1.0):
    """The reference's signature and normalisation (mnist/train.py:64-81): BCE mean over B*784, NLL mean over B,
    KL / (B * 784 / 3)."""
    kl_weight = 3.0 / (784.0 * mu.shape[0])
    return _ElboFn.apply(mu, logvar, recon_image, image, recon_text, text, lambda_xy, lambda_yx, kl_weight)


def elbo_loss(recon_image, image, recon_text, text, mu, logvar, lambda_image=1.0, lambda_text=1.0,
              annealing_factor=1.0):
    """North-star signature; same normalisation as `loss_function`, KL scaled by `annealing_factor`."""
    kl_weight = annealing_factor * 3.0 / (784.0 * mu.shape[0])
    return _ElboFn.apply(mu, logvar, recon_image, image, recon_text, text, lambda_image, lambda_text, kl_weight)
