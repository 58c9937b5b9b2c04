"""Checkpoint interop: the reference's `checkpoint.pth.tar` layout in and out, plus the optimiser state
the reference never restores (mnist/train.py:37-61,212-220; celeba/train.py:37-56,207-216; multimnist/train.py).

    save_checkpoint(state, is_best, folder, filename)        same signature and files as the reference
    load_checkpoint(path, family="mnist", ...)               -> model with the checkpoint's state_dict loaded
    trainer_state(trainer) / load_trainer_state(trainer, s)  Adam moments + step counter, keyed by the reference's
                                                             parameter names in the reference's shapes (portable
                                                             across precisions and across this library's internal layouts)
"""
from __future__ import annotations

import os
import shutil
from typing import Dict

import torch


def save_checkpoint(state: Dict, is_best: bool, folder: str = "./", filename: str = "checkpoint.pth.tar") -> None:
    """mnist/train.py:37-41."""
    os.makedirs(folder, exist_ok=True)
    torch.save(state, os.path.join(folder, filename))
    if is_best:
        shutil.copyfile(os.path.join(folder, filename), os.path.join(folder, "model_best.pth.tar"))


def _model_class(family: str):
    if family == "mnist":
        from .mnist import MVAE
        return MVAE
    if family == "celeba":
        from .celeba import MultimodalVAE
        return MultimodalVAE
    if family == "multimnist":
        from .multimnist import MultimodalVAE
        return MultimodalVAE
    raise ValueError("family must be mnist, celeba or multimnist")


def load_checkpoint(file_path: str, family: str = "mnist", precision: str = "tf32", **model_kwargs):
    """mnist/train.py:44-61: a model of the checkpoint's n_latents (default 20) with its state_dict loaded.  Reference
    checkpoints (CPU or CUDA tensors, reference shapes) load unchanged.  The default precision is "tf32" (any even
    n_latents that is a multiple of 4, e.g. the reference's 20); MNIST in "bf16" needs n_latents % 8 == 0 and falls
    back to "tf32" otherwise."""
    checkpoint = torch.load(file_path, map_location="cpu", weights_only=False)
    n_latents = int(checkpoint.get("n_latents", 20))
    if family == "mnist" and precision == "bf16" and n_latents % 8:
        precision = "tf32"
    vae = _model_class(family)(n_latents, precision=precision, **model_kwargs)
    vae.load_state_dict(checkpoint["state_dict"])
    return vae


def _param_views(model, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Reference-shaped copies of a flat per-parameter buffer (Adam moments share the parameters' layout)."""
    out = {}
    if hasattr(model, "layouts"):                     # conv models: internal layouts -> reference layouts
        for k, l in model.layouts.items():
            out[k] = l.to_reference(flat[l.offset:l.offset + l.numel]).cpu()
    else:                                             # MNIST: the flat buffer is already in state_dict order / shapes
        for name, kind, shape, off in model._table:
            if kind == 0:
                n = 1
                for s in shape:
                    n *= s
                out[name] = flat[off:off + n].view(shape).clone().cpu()
    return out


def _fill_flat(model, flat: torch.Tensor, tensors: Dict[str, torch.Tensor]) -> None:
    if hasattr(model, "layouts"):
        for k, l in model.layouts.items():
            flat[l.offset:l.offset + l.numel].copy_(l.to_internal(tensors[k].to(torch.float32)).reshape(-1))
    else:
        for name, kind, shape, off in model._table:
            if kind == 0:
                n = 1
                for s in shape:
                    n *= s
                flat[off:off + n].copy_(tensors[name].to(torch.float32).reshape(-1))


def trainer_state(trainer) -> Dict:
    """Adam first / second moments and the step counter of an MVAETrainer / CelebATrainer / MultiMNISTTrainer."""
    m = trainer.model
    if hasattr(trainer, "adam"):
        mom, vel, lr, betas, eps = trainer.adam["m"], trainer.adam["v"], trainer.adam["lr"], trainer.adam["betas"], trainer.adam["eps"]
    else:
        mom, vel, lr, betas, eps = trainer.adam_m, trainer.adam_v, trainer.lr, trainer.betas, trainer.eps
    return {"step": int(m._adam_counter.item()), "exp_avg": _param_views(m, mom), "exp_avg_sq": _param_views(m, vel),
            "lr": lr, "betas": tuple(betas), "eps": eps}


def load_trainer_state(trainer, state: Dict) -> None:
    m = trainer.model
    mom, vel = (trainer.adam["m"], trainer.adam["v"]) if hasattr(trainer, "adam") else (trainer.adam_m, trainer.adam_v)
    _fill_flat(m, mom, state["exp_avg"])
    _fill_flat(m, vel, state["exp_avg_sq"])
    m._adam_counter.fill_(int(state["step"]))
    m._step_counter.fill_(int(state["step"]))   # the noise counter resumes from the same point (any value is valid)
    if hasattr(trainer, "_graphs"):
        trainer._graphs.clear()       # captured graphs stay valid (state lives in the same buffers) but drop them to be safe
