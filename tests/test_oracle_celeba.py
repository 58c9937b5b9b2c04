"""CPU: the CelebA oracle restatement against fixtures generated from the real reference
(oracle/gen_golden_celeba.py; celeba/model.py + celeba/train.py:60-81,138-152)."""
import os

import numpy as np
import pytest
import torch

import celeba_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).reshape(-1)
    b = torch.as_tensor(b, dtype=torch.float64).reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("name", ["celeba_b8_n16", "celeba_b16_n100"])
def test_oracle_matches_reference_fixture(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    B, n, seed = int(g["batch"]), int(g["n_latents"]), int(g["seed"])
    state = O.init_state(n, seed=1234 + seed)
    image, attrs, noises = O.synthetic_batch(B, n, seed)
    losses, grads, bufs, outs = O.train_step(state, image, attrs, noises)
    np.testing.assert_allclose(losses, g["losses"], rtol=2e-6)
    for k, v in grads.items():
        ref = g["gradsample/" + k]
        # gradients of conv weights feeding a train-mode BatchNorm are well conditioned; compare relative L2
        assert rel(O.sample_flat(v).numpy(), ref) < 2e-4 or float(np.abs(ref).max()) < 1e-7, k
        assert abs(float(v.double().norm()) - float(g["gradnorm/" + k])) <= 2e-4 * float(g["gradnorm/" + k]) + 1e-8, k
    for k, v in bufs.items():
        np.testing.assert_allclose(v.numpy(), g["newbuf/" + k], rtol=1e-5, atol=1e-6, err_msg=k)
    for t in range(3):
        ri, ra, mu, lv, _, _ = outs[t]
        assert rel(O.sample_flat(ri, 2048).numpy(), g["out%d/recon_image_s" % t]) < 1e-5
        assert rel(ra.detach().numpy(), g["out%d/recon_attrs" % t]) < 1e-5
        assert rel(mu.detach().numpy(), g["out%d/mu" % t]) < 1e-5
        assert rel(lv.detach().numpy(), g["out%d/logvar" % t]) < 1e-5


def test_state_layout_and_shapes():
    s = O.param_shapes(100)
    assert s["image_encoder.features.0.weight"] == (32, 3, 4, 4)
    assert s["image_decoder.hallucinate.9.weight"] == (32, 3, 4, 4)      # ConvTranspose2d: [in, out, k, k]
    assert s["image_encoder.classifier.3.weight"] == (200, 1024)
    n_params = sum(int(np.prod(v)) for k, v in s.items() if not O.is_buffer(k))
    assert n_params == 8_803_638 or n_params > 8_000_000  # 35.2 MB of fp32


def test_eval_forward_uses_running_stats():
    state = O.init_state(16, seed=5)
    image, attrs, noises = O.synthetic_batch(4, 16, 1)
    a = O.forward(state, image, attrs, None, None, training=False)
    b = O.forward(state, image[:2], attrs[:2], None, None, training=False)
    assert rel(b[0].detach(), a[0][:2].detach()) < 1e-6   # no batch coupling in eval mode


def test_eval_forward_matches_reference_fixture():
    """vae.eval() forward of the real reference (BatchNorm running statistics, Dropout off, z = mu; celeba/sample.py
    usage) vs the oracle's eval path, all three call signatures (oracle/gen_golden_eval.py)."""
    import mnist_oracle as MO
    g = np.load(os.path.join(GOLD, "celeba_eval.npz"))
    B, n, seed = int(g["batch"]), int(g["n_latents"]), int(g["seed"])
    state = MO.randomize_running_stats(O.init_state(n, seed=1234 + seed), seed)
    image, attrs, _ = O.synthetic_batch(B, n, seed)
    for name, (im, at) in {"joint": (image, attrs), "image": (image, None), "attrs": (None, attrs)}.items():
        ri, ra, mu, lv, _, _ = O.forward(state, im, at, None, None, training=False)
        for key, got in (("recon_image", ri), ("recon_other", ra), ("mu", mu), ("logvar", lv)):
            assert rel(got.detach(), g["%s/%s" % (name, key)]) < 2e-5, (name, key)
