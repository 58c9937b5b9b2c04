"""CPU tests of oracle/mlp_oracle.py (the north-star MLP instantiation; "parity unpinned" - the reference has no such model,
so the oracle is pinned block by block): Swish and the reparametrize / loss terms against the reference-pinned MNIST oracle,
the prior-expert product against its closed form, and the analytic gradients of the whole step against central differences in
float64."""
import torch
import torch.nn.functional as F

import mlp_oracle as O
import mnist_oracle as M


def test_blocks_agree_with_the_pinned_mnist_oracle():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(7, 5, generator=g)
    assert torch.allclose(O.swish(x), x * torch.sigmoid(x))
    # product of experts with the prior expert == an explicit third expert N(0, 1)
    mu, lv = torch.randn(2, 6, 4, generator=g), torch.randn(2, 6, 4, generator=g)
    a = M.product_of_experts_precision(mu, lv, prior=True, eps=0.0)
    mu3 = torch.cat([mu, torch.zeros(1, 6, 4)])
    lv3 = torch.cat([lv, torch.zeros(1, 6, 4)])
    b = M.product_of_experts_precision(mu3, lv3, prior=False, eps=0.0)
    assert torch.allclose(a[0], b[0], atol=1e-6) and torch.allclose(a[1], b[1], atol=1e-6)
    # loss terms: BCE summed over pixels / B = 784 x the reference's mean; CE and KL as mnist/train.py:75,79
    B, n = 6, 4
    p, t = torch.rand(B, 784, generator=g), torch.rand(B, 784, generator=g)
    logp = F.log_softmax(torch.randn(B, 10, generator=g), dim=1)
    lab = torch.randint(0, 10, (B,), generator=g)
    m_, l_ = torch.randn(B, n, generator=g), torch.randn(B, n, generator=g)
    tot, bce, ce, kl = O.elbo_loss(p, t, logp, lab, m_, l_, 2.0, 3.0, 0.5)
    assert torch.allclose(bce, 2.0 * 784 * M.binary_cross_entropy_mean(p, t), rtol=1e-5)
    assert torch.allclose(ce, 3.0 * F.nll_loss(logp, lab), rtol=1e-6)
    assert torch.allclose(kl, 0.5 * (-0.5) * torch.sum(1 + l_ - m_ ** 2 - l_.exp()) / B, rtol=1e-6)
    assert torch.allclose(tot, bce + ce + kl)


def test_step_gradients_match_central_differences_in_float64():
    n, h, B = 4, 8, 5
    st = {k: v.double() for k, v in O.init_state(n, h, seed=3).items()}
    image, text, noises = O.synthetic_batch(B, n, 1)
    image, noises = image.double(), [e.double() for e in noises]
    lambdas = ((1.0, 10.0), (1.0, 0.0), (0.0, 50.0))
    _, _, grads = O.train_step(st, image, text, noises, lambdas=lambdas, annealing_factor=0.3)

    def total(state):
        return sum(l[0] for l in O.train_step(state, image, text, noises, lambdas=lambdas, annealing_factor=0.3)[0])

    gen = torch.Generator().manual_seed(9)
    for k, v in st.items():
        for _ in range(2):
            idx = int(torch.randint(0, v.numel(), (1,), generator=gen))
            if k == "text_encoder.embed.weight":      # only rows of labels present in the batch have a gradient: pick one
                idx = int(text[0]) * h + idx % h
            d = 1e-5
            up = {a: b.clone() for a, b in st.items()}
            dn = {a: b.clone() for a, b in st.items()}
            up[k].view(-1)[idx] += d
            dn[k].view(-1)[idx] -= d
            num = (total(up) - total(dn)) / (2 * d)
            ana = float(grads[k].view(-1)[idx])
            assert abs(num - ana) <= 1e-6 + 1e-4 * abs(ana), (k, idx, num, ana)


def test_text_encoder_is_a_function_of_the_label_only():
    """The product evaluates the text encoder on the ten labels and gathers rows; the oracle evaluates it per sample."""
    st = O.init_state(8, 16, seed=5)
    text = torch.tensor([3, 3, 7, 0, 9, 7])
    mu, lv = O.text_encoder(st, text)
    tmu, tlv = O.text_encoder(st, torch.arange(10))
    assert torch.equal(mu, tmu[text]) and torch.equal(lv, tlv[text])


def test_masked_step_equals_the_step_on_each_terms_own_rows():
    """Per-sample masks: term g of the masked step == the same term of an unmasked step fed only the rows that have the
    modalities the term needs (rows are independent: no normalisation layer), losses and gradients."""
    n, h, B = 4, 8, 12
    st = O.init_state(n, h, seed=7)
    image, text, noises = O.synthetic_batch(B, n, 2)
    g = torch.Generator().manual_seed(1)
    hi, ht = torch.rand(B, generator=g) < 0.6, torch.rand(B, generator=g) < 0.6
    hi[:2], ht[:2] = True, True
    lambdas = ((1.0, 2.0), (1.0, 0.0), (0.0, 3.0))
    losses, _, grads = O.train_step(st, image, text, noises, lambdas=lambdas, has_image=hi, has_text=ht)
    total = {k: torch.zeros_like(v) for k, v in st.items()}
    for t, name, rows in ((0, "joint", hi & ht), (1, "image", hi), (2, "text", ht)):
        idx = torch.nonzero(rows).reshape(-1)
        l, _, gr = O.train_step(st, image[idx], text[idx], [noises[t][idx]], terms=(name,), lambdas=(lambdas[t],))
        assert abs(l[0][0] - losses[t][0]) <= 1e-5 * abs(l[0][0])
        total = {k: total[k] + gr[k] for k in total}
    for k in grads:
        assert torch.allclose(grads[k], total[k], rtol=1e-4, atol=1e-6), k
