"""Generate tests/golden/*_eval.npz from the REAL reference in eval mode (run in the build container only).

The reference's evaluation scripts (mnist/test.py:18-36, mnist/loglikelihood.py:24-38, */sample.py) call the models under
`vae.eval()`: BatchNorm uses its running statistics, Dropout is off and `reparametrize` returns `mu`
(mnist/model.py:24-30).  For each family this loads a deterministic state with NON-trivial running statistics into the
reference's own `MultimodalVAE`, runs the three forward signatures and stores every output.  The fixtures pin the
oracles' eval path (tests/test_oracle*.py), against which the device's eval path is tested on the GPU.

    python oracle/gen_golden_eval.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, HERE)


def run(vae, state, calls):
    vae.load_state_dict({k: v.clone() for k, v in state.items()})
    vae.eval()
    outs = {}
    with torch.no_grad():
        for name, kw in calls.items():
            ri, ro, mu, lv = vae(**kw)
            outs[name + "/recon_image"] = ri.numpy().astype(np.float32)
            outs[name + "/recon_other"] = ro.numpy().astype(np.float32)
            outs[name + "/mu"] = mu.numpy().astype(np.float32)
            outs[name + "/logvar"] = lv.numpy().astype(np.float32)
    return outs


def main():
    os.makedirs(GOLD, exist_ok=True)
    B, n, seed = 6, 16, 4

    import gen_golden as GM
    import mnist_oracle as MO
    model_mod, _ = GM.import_reference_mnist()
    state = MO.randomize_running_stats(MO.perturbed_state(n, seed), seed)
    image, text, _ = MO.synthetic_batch(B, n, seed)
    outs = run(model_mod.MultimodalVAE(n_latents=n), state, {"joint": dict(image=image, text=text), "image": dict(image=image),
                                                             "text": dict(text=text)})
    np.savez_compressed(os.path.join(GOLD, "mnist_eval.npz"), batch=B, n_latents=n, seed=seed, **outs)

    # the reference's evaluation entry points themselves (mnist/test.py:18-36, mnist/loglikelihood.py:15-63), compiled out of
    # their files (the modules import torchvision / a py2-only train.py at top level) and run on two synthetic batches
    import ast
    from torch.autograd import Variable
    REF = GM.REF
    ns = {"torch": torch, "F": torch.nn.functional, "Variable": Variable, "xrange": range, "print": lambda *a, **k: None}
    for rel_path, fn_name in (("mnist/test.py", "test_mnist"), ("mnist/loglikelihood.py", "compute_nll")):
        tree = ast.parse(open(os.path.join(REF, rel_path)).read())
        fn = [node for node in tree.body if isinstance(node, ast.FunctionDef) and node.name == fn_name]
        exec(compile(ast.Module(body=fn, type_ignores=[]), rel_path, "exec"), ns)

    class Loader(list):
        @property
        def dataset(self):
            return range(sum(len(b[1]) for b in self))

    Be, S = 40, 3
    batches = Loader([MO.synthetic_batch(Be, n, s)[:2] for s in (1, 2)])
    vae = model_mod.MultimodalVAE(n_latents=n)
    vae.load_state_dict({k: v.clone() for k, v in state.items()})
    acc = float(ns["test_mnist"](vae, batches, verbose=False))
    res = {}
    for key, kw in (("joint", {}), ("image_only", {"image_only": True}), ("text_only", {"text_only": True})):
        torch.manual_seed(11)
        i_nll, t_nll = ns["compute_nll"](vae, batches, n_samples=S, **kw)
        res[key] = (float(i_nll), float(t_nll))
    np.savez_compressed(os.path.join(GOLD, "mnist_eval_scripts.npz"), batch=Be, n_latents=n, seed=seed, n_samples=S, noise_seed=11,
                        batch_seeds=np.array([1, 2]), accuracy=acc, **{"nll_" + k: np.array(v) for k, v in res.items()})
    print("mnist eval scripts: accuracy %.4f nll %s" % (acc, res))

    import gen_golden_celeba as GC
    import celeba_oracle as CO
    model_mod, _ = GC.import_reference_celeba()
    state = MO.randomize_running_stats(CO.init_state(n, seed=1234 + seed), seed)
    Bc = 3   # 3x64x64 images: keep the fixture small
    image, attrs, _ = CO.synthetic_batch(Bc, n, seed)
    outs = run(model_mod.MultimodalVAE(n_latents=n), state, {"joint": dict(image=image, attrs=attrs), "image": dict(image=image),
                                                             "attrs": dict(attrs=attrs)})
    np.savez_compressed(os.path.join(GOLD, "celeba_eval.npz"), batch=Bc, n_latents=n, seed=seed, **outs)

    import gen_golden_multimnist as GX
    import multimnist_oracle as XO
    model_mod, _ = GX.import_reference_multimnist()
    state = MO.randomize_running_stats(XO.init_state(n, seed=1234 + seed), seed)
    image, text, _ = XO.synthetic_batch(B, n, seed)
    outs = run(model_mod.MultimodalVAE(n_latents=n), state, {"joint": dict(image=image, text=text), "image": dict(image=image),
                                                             "text": dict(text=text)})
    np.savez_compressed(os.path.join(GOLD, "multimnist_eval.npz"), batch=B, n_latents=n, seed=seed, **outs)
    # multimnist/utils.py:22-56: charlist_tensor / tensor_to_string of the reference on a few label lists
    import builtins
    builtins.xrange = range
    for m in ("utils", "model", "train", "datasets"):
        sys.modules.pop(m, None)
    sys.path.insert(0, os.path.join(REF, "multimnist"))
    import utils as U  # type: ignore
    sys.path.pop(0)
    cases = [[], [7], [1, 2], [0, 0, 9], [9, 8, 7, 6], [3, 3], [5, 0, 5, 0]]
    exp = torch.stack([U.charlist_tensor(c) for c in cases]).numpy()
    strs = [U.tensor_to_string(torch.tensor(r)) for r in exp] + [U.tensor_to_string(torch.tensor([10, 4, 11, 2]))]
    flat = np.full((len(cases), 4), -1, dtype=np.int64)
    for i, c in enumerate(cases):
        flat[i, :len(c)] = c
    np.savez_compressed(os.path.join(GOLD, "multimnist_charlist.npz"), digits=flat, expected=exp, strings=np.array(strs))

    for f in ("mnist_eval", "celeba_eval", "multimnist_eval"):
        print(f, os.path.getsize(os.path.join(GOLD, f + ".npz")), "bytes")


if __name__ == "__main__":
    main()
