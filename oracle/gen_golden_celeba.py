"""Generate tests/golden/celeba_*.npz from the REAL reference (run in the build container only).

Imports celeba/model.py unmodified from /root/reference with a stub `datasets` module (the real
celeba/datasets.py:10 does not import on Python 3; only N_ATTRS = 18 is needed, celeba/datasets.py:27) and
takes loss_function from celeba/train.py by executing just its `def` (the module body needs torchvision
datasets and `xrange`).  The step celeba/train.py:138-152 is restated around the reference classes with
Dropout.p = 0 and the reparametrize noise injected.

    python oracle/gen_golden_celeba.py
"""
from __future__ import annotations

import ast
import builtins
import os
import sys
import types
import warnings

import numpy as np
import torch

REF = os.environ.get("MVAE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, HERE)
import celeba_oracle as O  # noqa: E402


def import_reference_celeba():
    for m in ("model", "train", "datasets", "utils"):
        sys.modules.pop(m, None)
    stub = types.ModuleType("datasets")
    stub.N_ATTRS = O.N_ATTRS
    sys.modules["datasets"] = stub
    sys.path.insert(0, os.path.join(REF, "celeba"))
    import model  # type: ignore
    sys.path.pop(0)
    # loss_function: compile only that function definition out of celeba/train.py
    src = open(os.path.join(REF, "celeba", "train.py")).read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "loss_function"]
    ns = {"torch": torch, "F": torch.nn.functional, "xrange": range}
    exec(compile(ast.Module(body=fn, type_ignores=[]), "celeba/train.py", "exec"), ns)
    return model, ns["loss_function"]


def ref_step(model_mod, loss_fn, n_latents, state, image, attrs, noises):
    vae = model_mod.MultimodalVAE(n_latents=n_latents)
    vae.load_state_dict({k: v.clone() for k, v in state.items()})
    vae.train()
    for m in vae.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    it = iter(noises)

    def reparam(mu, logvar):
        return next(it).mul(logvar.mul(0.5).exp()).add(mu)

    vae.reparametrize = reparam
    vae.zero_grad()
    calls = (dict(image=image, attrs=attrs), dict(image=image), dict(attrs=attrs))
    losses, outs, total = [], [], 0
    for k in range(3):
        ri, ra, mu, lv = vae(**calls[k])
        l = loss_fn(mu, lv, recon_x=ri, x=image, recon_y=ra, y=attrs)
        losses.append(l); outs.append((ri, ra, mu, lv)); total = total + l
    total.backward()
    grads = {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in vae.named_parameters()}
    return [float(l) for l in losses], grads, {k: v.clone() for k, v in vae.state_dict().items()}, outs


def dump(name, batch, n_latents, seed):
    model_mod, loss_fn = import_reference_celeba()
    state = O.init_state(n_latents, seed=1234 + seed)
    ref_keys = list(model_mod.MultimodalVAE(n_latents=n_latents).state_dict().keys())
    assert ref_keys == list(state.keys()), "state_dict layout differs from the reference"
    image, attrs, noises = O.synthetic_batch(batch, n_latents, seed)
    losses, grads, new_state, outs = ref_step(model_mod, loss_fn, n_latents, state, image, attrs, noises)
    out = {"batch": batch, "n_latents": n_latents, "seed": seed, "losses": np.array(losses, dtype=np.float64)}
    for k, v in new_state.items():
        if O.is_buffer(k):
            out["newbuf/" + k] = v.numpy()
    for k, v in grads.items():
        out["gradnorm/" + k] = np.array(float(v.double().norm()))
        out["gradsample/" + k] = O.sample_flat(v).numpy()
    for t, (ri, ra, mu, lv) in enumerate(outs):
        out["out%d/recon_image_s" % t] = O.sample_flat(ri, 2048).numpy()
        out["out%d/recon_attrs" % t] = ra.detach().numpy()
        out["out%d/mu" % t] = mu.detach().numpy()
        out["out%d/logvar" % t] = lv.detach().numpy()
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%.1f KB) losses=%s" % (path, os.path.getsize(path) / 1024, losses))


if __name__ == "__main__":
    warnings.filterwarnings("ignore")
    dump("celeba_b8_n16", 8, 16, seed=2)
    dump("celeba_b16_n100", 16, 100, seed=0)
