"""CPU: the MultiMNIST oracle restatement against fixtures generated from the real reference
(oracle/gen_golden_multimnist.py; multimnist/model.py + multimnist/train.py:69-87,148-168)."""
import os

import numpy as np
import pytest
import torch

import multimnist_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).reshape(-1)
    b = torch.as_tensor(b, dtype=torch.float64).reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("name", ["multimnist_b8_n16", "multimnist_b16_n100"])
def test_oracle_matches_reference_fixture(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    B, n, seed = int(g["batch"]), int(g["n_latents"]), int(g["seed"])
    state = O.init_state(n, seed=1234 + seed)
    image, text, noises = O.synthetic_batch(B, n, seed)
    losses, grads, bufs, outs = O.train_step(state, image, text, noises)
    np.testing.assert_allclose(losses, g["losses"], rtol=3e-6)
    for k, v in grads.items():
        ref = g["gradsample/" + k]
        assert rel(O.sample_flat(v).numpy(), ref) < 3e-4 or float(np.abs(ref).max()) < 1e-7, k
    for k, v in bufs.items():
        np.testing.assert_allclose(v.numpy(), g["newbuf/" + k], rtol=1e-5, atol=1e-6, err_msg=k)
    for t in range(3):
        ri, rt, mu, lv, _ = outs[t]
        assert rel(O.sample_flat(ri, 2048).numpy(), g["out%d/recon_image_s" % t]) < 1e-5
        assert rel(rt.detach().numpy(), g["out%d/recon_text" % t]) < 1e-5
        assert rel(mu.detach().numpy(), g["out%d/mu" % t]) < 1e-5
        assert rel(lv.detach().numpy(), g["out%d/logvar" % t]) < 1e-5


def test_text_encoder_reverse_direction_sees_only_last_character():
    """multimnist/model.py:243-246 takes x[-1]: the reverse GRU's output there depends on the last character alone."""
    st = O.init_state(16, seed=3)
    a = torch.tensor([[1, 2, 3, 4]]); b = torch.tensor([[9, 8, 7, 4]])
    g = "text_encoder.gru."
    e = st["text_encoder.embed.weight"]
    z = torch.zeros(1, O.N_HID)
    hb = lambda t: O.gru_cell(e[t[:, -1]], z, st[g + "weight_ih_l0_reverse"], st[g + "weight_hh_l0_reverse"],
                              st[g + "bias_ih_l0_reverse"], st[g + "bias_hh_l0_reverse"])
    assert torch.equal(hb(a), hb(b))
    ma, _ = O.text_encoder(st, a); mb, _ = O.text_encoder(st, b)
    assert not torch.allclose(ma, mb)      # the forward direction does see the whole string


def test_text_decoder_shape_and_normalisation():
    st = O.init_state(16, seed=4)
    lp = O.text_decoder(st, torch.randn(5, 16, generator=torch.Generator().manual_seed(0)))
    assert lp.shape == (5, O.MAX_LEN, O.N_CHARS)
    assert torch.allclose(lp.exp().sum(-1), torch.ones(5, O.MAX_LEN), atol=1e-5)


def test_eval_forward_matches_reference_fixture():
    """vae.eval() forward of the real reference (running statistics, Dropout off, z = mu, greedy text decode) vs the
    oracle's eval path, all three call signatures (oracle/gen_golden_eval.py)."""
    import mnist_oracle as MO
    g = np.load(os.path.join(GOLD, "multimnist_eval.npz"))
    B, n, seed = int(g["batch"]), int(g["n_latents"]), int(g["seed"])
    state = MO.randomize_running_stats(O.init_state(n, seed=1234 + seed), seed)
    image, text, _ = O.synthetic_batch(B, n, seed)
    for name, (im, tx) in {"joint": (image, text), "image": (image, None), "text": (None, text)}.items():
        ri, rt, mu, lv, _ = O.forward(state, im, tx, None, None, training=False)
        for key, got in (("recon_image", ri), ("recon_other", rt), ("mu", mu), ("logvar", lv)):
            assert rel(got.detach(), g["%s/%s" % (name, key)]) < 2e-5, (name, key)


def test_charlist_tensor_matches_reference_utils():
    """multimnist/utils.py:22-56 (`charlist_tensor`, `tensor_to_string`) of the real reference vs the batched host helpers
    (fixture written from the reference's functions by oracle/gen_golden_eval.py)."""
    import mvae_b200  # noqa: F401
    from mvae_b200 import multimnist as MM
    g = np.load(os.path.join(GOLD, "multimnist_charlist.npz"))
    cases = [[int(d) for d in row if d >= 0] for row in g["digits"]]
    got = MM.charlist_tensor(cases)
    assert got.dtype == torch.int64 and torch.equal(got, torch.from_numpy(g["expected"]))
    strings = [str(s) for s in g["strings"]]
    assert [MM.tensor_to_string(r) for r in got] == strings[:-1]
    assert MM.tensor_to_string(torch.tensor([10, 4, 11, 2])) == strings[-1]
    with pytest.raises(ValueError):
        MM.charlist_tensor([[1, 2, 3, 4, 5]])
