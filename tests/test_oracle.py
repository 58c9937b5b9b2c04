"""The CPU oracle against the golden vectors produced by the real reference
(oracle/gen_golden.py) and the known-answer tests of SURVEY.md section 4."""
import math
import os

import numpy as np
import pytest
import torch

import mnist_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def test_kat1_poe_is_variance_weighted():
    g = load("mnist_kats")
    mu = torch.tensor([[0.0], [2.0]]).view(2, 1, 1)
    lv = torch.tensor([[0.0], [math.log(3.0)]]).view(2, 1, 1)
    pm, pl = O.product_of_experts(mu, lv)
    assert float(pm) == pytest.approx(1.5, abs=1e-6) == pytest.approx(float(g["kat1_mu"]), abs=1e-7)
    assert float(pl) == pytest.approx(math.log(0.75), abs=1e-6)
    assert float(pl) == pytest.approx(float(g["kat1_logvar"]), abs=1e-7)
    # precision-weighted variant gives 0.5 (the north-star math), so the two modes are distinguishable
    pm2, _ = O.product_of_experts_precision(mu, lv)
    assert float(pm2) == pytest.approx(0.5, abs=1e-6)


def test_kat_poe_random_matches_reference():
    g = load("mnist_kats")
    pm, pl = O.product_of_experts(torch.from_numpy(g["poe_mu_in"]), torch.from_numpy(g["poe_logvar_in"]))
    np.testing.assert_allclose(pm.numpy(), g["poe_mu"], rtol=0, atol=0)
    np.testing.assert_allclose(pl.numpy(), g["poe_logvar"], rtol=0, atol=0)


def test_kat2_kat3_loss():
    g = load("mnist_kats")
    l2 = O.loss_function(torch.zeros(4, 64), torch.zeros(4, 64), torch.full((4, 784), 0.5), torch.rand(4, 784),
                         torch.full((4, 10), math.log(0.1)), torch.randint(0, 10, (4,)))
    assert float(l2) == pytest.approx(math.log(2) + math.log(10), abs=1e-6)
    assert float(l2) == pytest.approx(float(g["kat2"]), abs=1e-6)
    l3 = O.loss_function(torch.ones(5, 64), torch.zeros(5, 64))
    assert float(l3) == pytest.approx(0.5 * 64 * 3 / 784, abs=1e-7)
    assert float(l3) == pytest.approx(float(g["kat3"]), abs=1e-7)


def test_kat5_single_expert_identity():
    mu, lv = torch.randn(1, 9, 16), torch.randn(1, 9, 16)
    pm, pl = O.product_of_experts(mu, lv)
    assert (pm - mu[0]).abs().max() < 1e-6  # mu*var/var: exact up to one rounding
    assert (pl - lv[0]).abs().max() < 1e-6


def test_kat6_eval_reparametrize_returns_mu():
    mu, lv = torch.randn(5, 8), torch.randn(5, 8)
    assert O.reparametrize(mu, lv, training=False) is mu


@pytest.mark.parametrize("name", ["mnist_b24_n8", "mnist_b100_n64", "mnist_b32_n20_weak"])
def test_step_matches_reference_golden(name):
    g = load(name)
    B, n, seed = int(g["batch"]), int(g["n_latents"]), int(g["seed"])
    state = O.perturbed_state(n, seed)
    image, text, noises = O.synthetic_batch(B, n, seed)
    assert np.array_equal(text.numpy(), g["text"])
    lambdas = tuple(tuple(float(x) for x in r) for r in g["lambdas"])
    terms = tuple(bool(x) for x in g["terms"])
    losses, grads, bufs, outs = O.train_step(state, image, text, noises, lambdas, terms)
    np.testing.assert_allclose(losses, g["losses"], rtol=2e-6, atol=1e-7)
    for k, v in grads.items():
        ref_s = g["gradsample/" + k]
        if k in O.PRE_BN_BIASES:
            # a bias feeding train-mode BatchNorm has an exactly-zero gradient; both sides hold rounding noise
            assert float(v.abs().max()) < 1e-6 and float(np.abs(ref_s).max()) < 1e-6, k
            continue
        scale = max(float(np.abs(ref_s).max()), 1e-8)
        np.testing.assert_allclose(O.sample_flat(v).numpy(), ref_s, rtol=2e-4, atol=2e-5 * scale + 1e-9, err_msg=k)
        assert float(v.double().norm()) == pytest.approx(float(g["gradnorm/" + k]), rel=1e-4, abs=1e-8), k
    for k, v in bufs.items():
        np.testing.assert_allclose(v.numpy(), g["newbuf/" + k], rtol=1e-5, atol=1e-6, err_msg=k)
    for t in range(3):
        if not terms[t]:
            continue
        ri, rt, mu, lv, _, _ = outs[t]
        np.testing.assert_allclose(ri.detach()[:, ::7].numpy(), g["out%d/recon_image_s" % t], rtol=1e-5, atol=1e-6)
        if "out%d/mu" % t in g:
            np.testing.assert_allclose(mu.detach().numpy(), g["out%d/mu" % t], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(lv.detach().numpy(), g["out%d/logvar" % t], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(rt.detach().numpy(), g["out%d/recon_text" % t], rtol=1e-5, atol=1e-6)


def test_kat4_bn_bookkeeping():
    """After ONE 3-term step num_batches_tracked is 2 for encoder BNs and 3 for decoder BNs."""
    state = O.init_state(8)
    image, text, noises = O.synthetic_batch(16, 8, 1)
    _, _, bufs, _ = O.train_step(state, image, text, noises)
    assert int(bufs["image_encoder.net.1.num_batches_tracked"]) == 2
    assert int(bufs["text_encoder.net.1.num_batches_tracked"]) == 2
    assert int(bufs["image_decoder.net.4.num_batches_tracked"]) == 3
    assert int(bufs["text_decoder.net.1.num_batches_tracked"]) == 3


def test_adam_matches_torch_optim():
    torch.manual_seed(0)
    w = torch.randn(7, 5)
    p = {"a.weight": w.clone()}
    ref = torch.nn.Parameter(w.clone())
    opt = torch.optim.Adam([ref], lr=1e-3)
    m = {"a.weight": torch.zeros_like(w)}
    v = {"a.weight": torch.zeros_like(w)}
    for step in range(1, 4):
        g = torch.randn(7, 5)
        ref.grad = g.clone()
        opt.step()
        p = O.adam_step(p, {"a.weight": g}, m, v, step)
        np.testing.assert_allclose(p["a.weight"].numpy(), ref.detach().numpy(), rtol=1e-6, atol=1e-7)


def test_eval_forward_matches_reference_fixture():
    """vae.eval() forward of the real reference (running statistics, z = mu; mnist/test.py:18-36) vs the oracle."""
    g = np.load(os.path.join(GOLD, "mnist_eval.npz"))
    B, n, seed = int(g["batch"]), int(g["n_latents"]), int(g["seed"])
    state = O.randomize_running_stats(O.perturbed_state(n, seed), seed)
    image, text, _ = O.synthetic_batch(B, n, seed)
    for name, (im, tx) in {"joint": (image, text), "image": (image, None), "text": (None, text)}.items():
        ri, rt, mu, lv, _, _ = O.forward(state, im, tx, None, None, training=False)
        for key, got in (("recon_image", ri), ("recon_other", rt), ("mu", mu), ("logvar", lv)):
            np.testing.assert_allclose(got.detach().numpy(), g["%s/%s" % (name, key)], rtol=2e-5, atol=2e-6, err_msg=name + key)


def test_evaluation_entry_points_match_reference_functions():
    """mnist/test.py:18-36 (`test_mnist`) and mnist/loglikelihood.py:15-63 (`compute_nll`), the reference's OWN functions run
    on its own model (oracle/gen_golden_eval.py), vs the oracle's restatements that the GPU evaluation test checks against."""
    g = np.load(os.path.join(GOLD, "mnist_eval_scripts.npz"))
    B, n, seed, S = int(g["batch"]), int(g["n_latents"]), int(g["seed"]), int(g["n_samples"])
    state = O.randomize_running_stats(O.perturbed_state(n, seed), seed)
    batches = [O.synthetic_batch(B, n, int(s))[:2] for s in g["batch_seeds"]]
    assert O.test_mnist(state, batches) == pytest.approx(float(g["accuracy"]), abs=1e-7)   # the reference divides in fp32
    for key, kw in (("joint", {}), ("image_only", {"image_only": True}), ("text_only", {"text_only": True})):
        gen = torch.Generator().manual_seed(int(g["noise_seed"]))
        got = O.compute_nll(state, batches, n_samples=S, generator=gen, **kw)
        np.testing.assert_allclose(got, g["nll_" + key], rtol=2e-6, err_msg=key)
