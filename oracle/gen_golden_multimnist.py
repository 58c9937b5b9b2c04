"""Generate tests/golden/multimnist_*.npz from the REAL reference (run in the build container only).

Imports multimnist/model.py unmodified from /root/reference (`builtins.xrange = range` first, multimnist/model.py:282;
its `utils` import resolves to multimnist/utils.py) and takes loss_function from multimnist/train.py by executing just
its `def`.  The step multimnist/train.py:148-168 is restated around the reference classes with Dropout.p = 0,
GRU.dropout = 0 and the reparametrize noise injected.

    python oracle/gen_golden_multimnist.py
"""
from __future__ import annotations

import ast
import builtins
import os
import sys
import warnings

import numpy as np
import torch

REF = os.environ.get("MVAE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, HERE)
import multimnist_oracle as O  # noqa: E402


def import_reference_multimnist():
    for m in ("model", "train", "datasets", "utils"):
        sys.modules.pop(m, None)
    builtins.xrange = range
    sys.path.insert(0, os.path.join(REF, "multimnist"))
    import model  # type: ignore
    sys.path.pop(0)
    src = open(os.path.join(REF, "multimnist", "train.py")).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "loss_function"]
    ns = {"torch": torch, "F": torch.nn.functional}
    exec(compile(ast.Module(body=fn, type_ignores=[]), "multimnist/train.py", "exec"), ns)
    return model, ns["loss_function"]


def ref_step(model_mod, loss_fn, n_latents, state, image, text, noises):
    vae = model_mod.MultimodalVAE(n_latents=n_latents)
    vae.load_state_dict({k: v.clone() for k, v in state.items()})
    vae.train()
    for m in vae.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.GRU):
            m.dropout = 0.0
    it = iter(noises)
    vae.reparametrize = lambda mu, logvar: next(it).mul(logvar.mul(0.5).exp()).add(mu)
    vae.zero_grad()
    calls = (dict(image=image, text=text), dict(image=image), dict(text=text))
    losses, outs, total = [], [], 0
    for k in range(3):
        ri, rt, mu, lv = vae(**calls[k])
        l = loss_fn(mu, lv, recon_image=ri, image=image, recon_text=rt, text=text, kl_lambda=O.KL_LAMBDA,
                    lambda_xy=O.LAMBDAS[k][0], lambda_yx=O.LAMBDAS[k][1])
        losses.append(l); outs.append((ri, rt, mu, lv)); total = total + l
    total.backward()
    grads = {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in vae.named_parameters()}
    return [float(l) for l in losses], grads, {k: v.clone() for k, v in vae.state_dict().items()}, outs


def dump(name, batch, n_latents, seed):
    model_mod, loss_fn = import_reference_multimnist()
    state = O.init_state(n_latents, seed=1234 + seed)
    ref_keys = list(model_mod.MultimodalVAE(n_latents=n_latents).state_dict().keys())
    assert ref_keys == list(state.keys()), [a for a, b in zip(ref_keys, state.keys()) if a != b][:5]
    image, text, noises = O.synthetic_batch(batch, n_latents, seed)
    losses, grads, new_state, outs = ref_step(model_mod, loss_fn, n_latents, state, image, text, noises)
    out = {"batch": batch, "n_latents": n_latents, "seed": seed, "losses": np.array(losses, dtype=np.float64)}
    for k, v in new_state.items():
        if O.is_buffer(k):
            out["newbuf/" + k] = v.numpy()
    for k, v in grads.items():
        out["gradnorm/" + k] = np.array(float(v.double().norm()))
        out["gradsample/" + k] = O.sample_flat(v).numpy()
    for t, (ri, rt, mu, lv) in enumerate(outs):
        out["out%d/recon_image_s" % t] = O.sample_flat(ri, 2048).numpy()
        out["out%d/recon_text" % t] = rt.detach().numpy()
        out["out%d/mu" % t] = mu.detach().numpy()
        out["out%d/logvar" % t] = lv.detach().numpy()
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%.1f KB) losses=%s" % (path, os.path.getsize(path) / 1024, losses))


if __name__ == "__main__":
    warnings.filterwarnings("ignore")
    dump("multimnist_b8_n16", 8, 16, seed=2)
    dump("multimnist_b16_n100", 16, 100, seed=0)
