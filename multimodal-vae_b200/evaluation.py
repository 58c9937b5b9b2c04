"""The reference's evaluation entry points on the B200 library.

    test_mnist(model, loader)                      mnist/test.py:18-36         label accuracy of vae(image=image)
    compute_nll(model, loader, ..., n_samples)     mnist/loglikelihood.py:15-63 sampled reconstruction NLL

Everything numerical runs through the C ABI: the eval-mode forward (BatchNorm running statistics, z = mu), and per sample
one fused decode + BCE/NLL launch sequence (MVAE.decode_losses).  The host only draws the shared N(0,1) samples and scales
them by exp(logvar / 2) exactly as the reference script does.
"""
from __future__ import annotations

from typing import Iterable, Optional, Tuple

import torch


@torch.no_grad()
def test_mnist(model, loader: Iterable[Tuple[torch.Tensor, torch.Tensor]]) -> float:
    """mnist/test.py:18-36: fraction of labels recovered from the image alone (argmax of recon_text)."""
    model.eval()
    correct, total = 0, 0
    for image, text in loader:
        _, recon_text, _, _ = model(image=image.reshape(image.shape[0], -1))
        pred = recon_text.argmax(dim=1)
        correct += int((pred.cpu() == text.cpu().reshape(-1)).sum())
        total += int(text.numel())
    return correct / float(max(total, 1))


@torch.no_grad()
def compute_nll(model, loader: Iterable[Tuple[torch.Tensor, torch.Tensor]], image_only: bool = False, text_only: bool = False,
                n_samples: int = 1, generator: Optional[torch.Generator] = None) -> Tuple[float, float]:
    """mnist/loglikelihood.py:15-63: (image NLL, text NLL) per example, z ~ q(z | inputs) with `n_samples` draws that are
    shared by the whole batch (as in the reference: one [n_samples, n_latents] normal tensor per batch)."""
    assert not (image_only and text_only)
    model.eval()
    image_nll = torch.zeros((), device=model.device_, dtype=torch.float64)
    text_nll = torch.zeros((), device=model.device_, dtype=torch.float64)
    total = 0
    for image, text in loader:
        image = image.reshape(image.shape[0], -1)
        if image_only:
            _, _, mu, logvar = model(image=image)
        elif text_only:
            _, _, mu, logvar = model(text=text)
        else:
            _, _, mu, logvar = model(image, text)
        B, n = mu.shape
        sample = torch.randn(n_samples, n, generator=generator).to(mu.device)
        std = torch.exp(0.5 * logvar)
        for i in range(n_samples):
            z = sample[i].unsqueeze(0) * std + mu
            means = model.decode_losses(z, image, text).double()
            image_nll += means[0] * (B * 784) / n_samples
            text_nll += means[1] * B / n_samples
        total += B
    return float(image_nll) / max(total, 1), float(text_nll) / max(total, 1)
