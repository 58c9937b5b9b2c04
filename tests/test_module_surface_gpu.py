"""GPU tests of the reference-facing module surface: ProductOfExperts (+mask), loss_function / elbo_loss,
MVAE.forward with autograd (the reference's own three-forward training loop), eval mode, sub-calls."""
import math
import os

import numpy as np
import pytest
import torch

import mnist_oracle as O
from helpers import rel_l2

pytestmark = pytest.mark.gpu


def test_poe_kat1_ref_vs_precision():
    import mvae_b200
    mu = torch.tensor([[0.0], [2.0]]).view(2, 1, 1).cuda()
    lv = torch.tensor([[0.0], [math.log(3.0)]]).view(2, 1, 1).cuda()
    pm, pl = mvae_b200.ProductOfExperts()(mu, lv)
    assert float(pm) == pytest.approx(1.5, abs=1e-6)           # the reference's variance-weighted mean (KAT-1)
    assert float(pl) == pytest.approx(math.log(0.75), abs=1e-6)
    pm2, pl2 = mvae_b200.ProductOfExperts("precision")(mu, lv)
    assert float(pm2) == pytest.approx(0.5, abs=1e-6)          # the paper's precision-weighted mean
    assert float(pl2) == pytest.approx(math.log(0.75), abs=1e-6)


@pytest.mark.parametrize("mode,prior,masked", [("ref", False, False), ("precision", False, False),
                                               ("precision", True, True), ("ref", False, True)])
def test_poe_matches_oracle_forward_backward(mode, prior, masked):
    import mvae_b200
    g = torch.Generator().manual_seed(3)
    M, B, D = 3, 37, 20
    mu = torch.randn(M, B, D, generator=g)
    lv = torch.randn(M, B, D, generator=g)
    mask = None
    if masked:
        mask = (torch.rand(M, B, generator=g) > 0.4).float()
        mask[0] = 1.0  # at least one expert per sample
    mu_r, lv_r = mu.clone().requires_grad_(True), lv.clone().requires_grad_(True)
    if mode == "ref" and not masked:
        pm, pl = O.product_of_experts(mu_r, lv_r)
    elif mode == "ref":
        var = torch.exp(lv_r) + 1e-8
        w = mask.unsqueeze(-1)
        pm = (mu_r * var * w).sum(0) / (var * w).sum(0)
        pl = torch.log(1.0 / (w / var).sum(0))
    else:
        pm, pl = O.product_of_experts_precision(mu_r, lv_r, mask, prior)
    gm, gl = torch.randn(B, D, generator=g), torch.randn(B, D, generator=g)
    (pm * gm + pl * gl).sum().backward()
    mu_d, lv_d = mu.cuda().requires_grad_(True), lv.cuda().requires_grad_(True)
    dm, dl = mvae_b200.ProductOfExperts(mode, prior)(mu_d, lv_d, None if mask is None else mask.cuda())
    (dm * gm.cuda() + dl * gl.cuda()).sum().backward()
    np.testing.assert_allclose(dm.detach().cpu().numpy(), pm.detach().numpy(), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(dl.detach().cpu().numpy(), pl.detach().numpy(), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(mu_d.grad.cpu().numpy(), mu_r.grad.numpy(), rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(lv_d.grad.cpu().numpy(), lv_r.grad.numpy(), rtol=2e-4, atol=2e-6)


def test_loss_function_matches_reference_semantics():
    import mvae_b200
    g = torch.Generator().manual_seed(5)
    B, n = 33, 16
    mu, lv = torch.randn(B, n, generator=g), torch.randn(B, n, generator=g) * 0.3
    ri = torch.rand(B, 784, generator=g).clamp(1e-4, 1 - 1e-4)
    img = torch.rand(B, 784, generator=g)
    rt = torch.log_softmax(torch.randn(B, 10, generator=g), 1)
    txt = torch.randint(0, 10, (B,), generator=g)
    # KAT-2 / KAT-3 of SURVEY section 4
    k2 = mvae_b200.loss_function(torch.zeros(4, 64).cuda(), torch.zeros(4, 64).cuda(), torch.full((4, 784), 0.5).cuda(),
                                 torch.rand(4, 784).cuda(), torch.full((4, 10), math.log(0.1)).cuda(),
                                 torch.randint(0, 10, (4,)).cuda())
    assert float(k2) == pytest.approx(math.log(2) + math.log(10), abs=2e-6)
    k3 = mvae_b200.loss_function(torch.ones(5, 64).cuda(), torch.zeros(5, 64).cuda())
    assert float(k3) == pytest.approx(0.5 * 64 * 3 / 784, abs=1e-7)
    for lam in ((1.0, 1.0), (1.0, 0.5), (0.0, 1.0)):
        leaves = [t.clone().requires_grad_(True) for t in (mu, lv, ri, rt)]
        ref = O.loss_function(leaves[0], leaves[1], leaves[2], img, leaves[3], txt, lam[0], lam[1])
        ref.backward()
        dleaves = [t.cuda().requires_grad_(True) for t in (mu, lv, ri, rt)]
        got = mvae_b200.loss_function(dleaves[0], dleaves[1], dleaves[2], img.cuda(), dleaves[3], txt.cuda(), lam[0], lam[1])
        got.backward()
        assert float(got.detach()) == pytest.approx(float(ref.detach()), rel=2e-6)
        for a, b in zip(dleaves, leaves):
            np.testing.assert_allclose(a.grad.cpu().numpy(), b.grad.numpy(), rtol=1e-4, atol=1e-8)
    e = mvae_b200.elbo_loss(ri.cuda(), img.cuda(), rt.cuda(), txt.cuda(), mu.cuda(), lv.cuda(), 1.0, 1.0, 0.25)
    r = O.loss_function(mu, lv, ri, img, rt, txt) - 0.75 * O.loss_function(mu, lv)
    assert float(e) == pytest.approx(float(r), rel=2e-6)


def test_reference_training_loop_with_autograd():
    """mnist/train.py:132-148 verbatim on the drop-in module: three forwards, three loss_function calls, backward."""
    import mvae_b200
    B, n, seed = 96, 16, 2
    state = O.perturbed_state(n, seed)
    image, text, noises = O.synthetic_batch(B, n, seed)
    losses, grads, bufs, _ = O.train_step(state, image, text, noises)
    vae = mvae_b200.MultimodalVAE(n_latents=n, precision="tf32")
    vae.load_state_dict(state)
    vae.train()
    vae._injected_noise = [x.clone() for x in noises]
    optimizer = torch.optim.Adam(vae.parameters(), lr=1e-3)
    optimizer.zero_grad()
    img_d, txt_d = image.cuda(), text.cuda()
    r1 = vae(img_d, txt_d)
    r2 = vae(image=img_d)
    r3 = vae(text=txt_d)
    l1 = mvae_b200.loss_function(r1[2], r1[3], recon_image=r1[0], image=img_d, recon_text=r1[1], text=txt_d)
    l2 = mvae_b200.loss_function(r2[2], r2[3], recon_image=r2[0], image=img_d, recon_text=r2[1], text=txt_d)
    l3 = mvae_b200.loss_function(r3[2], r3[3], recon_image=r3[0], image=img_d, recon_text=r3[1], text=txt_d)
    (l1 + l2 + l3).backward()
    np.testing.assert_allclose([float(l1), float(l2), float(l3)], losses, rtol=3e-5)
    sd = vae.state_dict()
    for k, v in bufs.items():                      # KAT-4 bookkeeping: 2 / 2 / 3 / 3
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == int(v), k
    for name, p in vae.named_parameters():
        if name in O.PRE_BN_BIASES:
            continue
        assert rel_l2(p.grad, grads[name]) < 6e-2, (name, rel_l2(p.grad, grads[name]))
    optimizer.step()


def test_reference_training_loop_bf16_sees_the_optimizer_update():
    """in bf16 the GEMMs read a bf16 mirror of the parameters; the module surface must refresh it after a
    torch optimizer stepped the fp32 master - the loss of the reference loop has to move from step to step."""
    import mvae_b200
    B, n = 128, 16
    state = O.perturbed_state(n, 4)
    image, text, _ = O.synthetic_batch(B, n, 4)
    vae = mvae_b200.MultimodalVAE(n_latents=n, precision="bf16")
    vae.load_state_dict(state)
    vae.train()
    optimizer = torch.optim.Adam(vae.parameters(), lr=1e-2)
    img_d, txt_d = image.cuda(), text.cuda()
    g = torch.Generator().manual_seed(0)
    fixed = [torch.randn(B, n, generator=g) for _ in range(3)]
    curve = []
    for it in range(6):
        vae._injected_noise = [x.clone() for x in fixed]
        optimizer.zero_grad()
        r1, r2, r3 = vae(img_d, txt_d), vae(image=img_d), vae(text=txt_d)
        loss = sum(mvae_b200.loss_function(r[2], r[3], recon_image=r[0], image=img_d, recon_text=r[1], text=txt_d)
                   for r in (r1, r2, r3))
        loss.backward()
        optimizer.step()
        curve.append(float(loss))
    assert curve[-1] < curve[0] - 1e-3, curve          # the weights the GEMMs see really move
    assert len(set(round(c, 6) for c in curve)) == len(curve), curve
    vae.eval()
    before = vae(img_d, txt_d)[0].float().clone()
    with torch.no_grad():
        for p in vae.parameters():
            p.mul_(1.05)
    after = vae(img_d, txt_d)[0].float()
    assert float((before - after).abs().max()) > 1e-3   # eval forward also sees edited parameters


def test_load_checkpoint_default_handles_reference_n_latents(tmp_path):
    """the reference's stock checkpoint has n_latents = 20 (mnist/train.py:54,87); load_checkpoint(path) must
    load it with its defaults (and an explicit bf16 request falls back to tf32 for it)."""
    from mvae_b200 import checkpoint
    state = O.init_state(20, seed=5)
    path = str(tmp_path / "ck")
    checkpoint.save_checkpoint({"state_dict": state, "n_latents": 20}, False, folder=path)
    for kw in ({}, {"precision": "bf16"}):
        vae = checkpoint.load_checkpoint(path + "/checkpoint.pth.tar", **kw)
        assert vae.n_latents == 20 and vae.precision == "tf32"
        sd = vae.state_dict()
        for k, v in state.items():
            assert torch.equal(sd[k].cpu().to(v.dtype), v), k


def test_eval_mode_forward_and_subcalls():
    import mvae_b200
    B, n, seed = 40, 16, 7
    state = O.perturbed_state(n, seed)
    g = torch.Generator().manual_seed(1)
    for k in state:   # non-trivial running statistics
        if k.endswith("running_mean"):
            state[k] = 0.1 * torch.randn(state[k].shape, generator=g)
        if k.endswith("running_var"):
            state[k] = 0.5 + torch.rand(state[k].shape, generator=g)
    image, text, _ = O.synthetic_batch(B, n, seed)
    vae = mvae_b200.MVAE(n, precision="tf32")
    vae.load_state_dict(state)
    vae.eval()
    for args in ((image, text), (image, None), (None, text)):
        ri, rt, mu, lv = vae(None if args[0] is None else args[0].cuda(), None if args[1] is None else args[1].cuda())
        o = O.forward(state, args[0], args[1], None, None, training=False)
        assert rel_l2(ri.float(), o[0]) < 1e-3 and rel_l2(rt, o[1]) < 1e-3
        assert rel_l2(mu, o[2]) < 2e-3 and rel_l2(lv, o[3]) < 2e-3
    sd = vae.state_dict()
    for k, v in state.items():                      # eval mode updates nothing
        if O.is_buffer(k):
            assert torch.equal(sd[k].cpu(), v), k
    # sub-calls of mnist/sample.py / manifold.py
    mi, li = vae.encode_image(image.cuda())
    om, ol = O.image_encoder(state, image, None, training=False)
    assert rel_l2(mi, om) < 2e-3 and rel_l2(li, ol) < 2e-3
    mt, lt = vae.encode_text(text.cuda())
    tm, tl = O.text_encoder(state, text, None, training=False)
    assert rel_l2(mt, tm) < 1e-4 and rel_l2(lt, tl) < 1e-4
    z = torch.randn(B, n, generator=g)
    di, dt_ = vae.decode_image(z.cuda()), vae.decode_text(z.cuda())
    assert rel_l2(di.float(), torch.sigmoid(O.image_decoder_logits(state, z, None, training=False))) < 1e-3
    assert rel_l2(dt_, torch.log_softmax(O.text_decoder_logits(state, z, None, training=False), 1)) < 1e-4
    pm, pl = vae.experts(torch.stack((mi, mt)), torch.stack((li, lt)))
    rm, rl = O.product_of_experts(torch.stack((mi.cpu(), mt.cpu())), torch.stack((li.cpu(), lt.cpu())))
    assert rel_l2(pm, rm) < 1e-5 and rel_l2(pl, rl) < 1e-5
    assert vae.gen_latents(image.cuda(), text.cuda()).shape == (B, n)
    assert vae.n_latents == n


def test_evaluation_entry_points_match_oracle():
    """mnist/test.py:18-36 and mnist/loglikelihood.py:15-63 on the library (eval forward, fused decode + losses)."""
    import mnist_oracle as O
    import mvae_b200
    from mvae_b200 import evaluation
    n, B, S = 16, 40, 3
    state = O.perturbed_state(n, 4)
    g = torch.Generator().manual_seed(3)
    for k in state:                                     # non-trivial running statistics
        if k.endswith("running_mean"):
            state[k] = 0.1 * torch.randn(state[k].shape, generator=g)
        if k.endswith("running_var"):
            state[k] = 0.5 + torch.rand(state[k].shape, generator=g)
    m = mvae_b200.MVAE(n, precision="tf32")
    m.load_state_dict(state)
    batches = [O.synthetic_batch(B, n, s)[:2] for s in (1, 2)]
    # accuracy (O.test_mnist / O.compute_nll are pinned to the reference's own functions: tests/test_oracle.py)
    acc = evaluation.test_mnist(m, batches)
    assert abs(acc - O.test_mnist(state, batches)) <= 1.0 / (2 * B) + 1e-9          # tf32: at most one near-tie flip
    # sampled NLL with the same shared draws
    gen = torch.Generator().manual_seed(11)
    got = evaluation.compute_nll(m, batches, n_samples=S, generator=gen)
    gen = torch.Generator().manual_seed(11)
    ref_i, ref_t = O.compute_nll(state, batches, n_samples=S, generator=gen)
    assert abs(got[0] - ref_i) <= 2e-3 * ref_i and abs(got[1] - ref_t) <= 5e-3 * ref_t, (got, ref_i, ref_t)


@pytest.mark.parametrize("family", ["mnist", "celeba", "multimnist"])
def test_checkpoint_roundtrip_and_resume(family, tmp_path):
    """mnist/train.py:37-61,212-220: save_checkpoint / load_checkpoint with the reference's dict layout; resuming with the
    saved optimiser state continues exactly like the uninterrupted run (up to the atomics-order noise of a step)."""
    import importlib
    import mvae_b200
    from mvae_b200 import checkpoint as ck
    n, B = 16, 16
    if family == "mnist":
        import mnist_oracle as O
        Model, Trainer = mvae_b200.MVAE, mvae_b200.MVAETrainer
        mk = lambda: Model(n, precision="tf32")
    else:
        O = importlib.import_module(family + "_oracle")
        mod = importlib.import_module("mvae_b200." + family)
        Model = mod.MultimodalVAE
        Trainer = mod.CelebATrainer if family == "celeba" else mod.MultiMNISTTrainer
        mk = lambda: Model(n_latents=n, precision="tf32", dropout_p=0.0)
    image, other, noises = O.synthetic_batch(B, n, 3)
    eps = torch.stack(noises).cuda()
    image, other = image.cuda(), other.cuda()

    def total(tr):
        if family == "mnist":
            return float(tr.step(image, other, eps=eps)[0][:, 0].sum())
        tr.step(image, other, eps=eps)
        return sum(l[0] for l in tr.losses())

    m1 = mk(); t1 = Trainer(m1)
    for _ in range(3):
        total(t1)
    ck.save_checkpoint({"state_dict": m1.state_dict(), "n_latents": n, "optimizer": ck.trainer_state(t1)}, True, str(tmp_path))
    assert os.path.exists(os.path.join(str(tmp_path), "model_best.pth.tar"))
    cont = [total(t1) for _ in range(2)]
    m2 = ck.load_checkpoint(os.path.join(str(tmp_path), "checkpoint.pth.tar"), family=family, precision="tf32",
                            **({} if family == "mnist" else {"dropout_p": 0.0}))
    sd1 = torch.load(os.path.join(str(tmp_path), "checkpoint.pth.tar"), weights_only=False)["state_dict"]
    sd2 = m2.state_dict()
    for k in sd1:
        assert torch.equal(sd1[k].cpu(), sd2[k].cpu()), k
    t2 = Trainer(m2)
    ck.load_trainer_state(t2, torch.load(os.path.join(str(tmp_path), "checkpoint.pth.tar"), weights_only=False)["optimizer"])
    resumed = [total(t2) for _ in range(2)]
    for a, b in zip(cont, resumed):
        assert abs(a - b) <= 2e-3 * abs(a), (cont, resumed)


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
def test_uint8_images_are_converted_on_the_device(precision):
    """uint8 pixels -> activation dtype by mvae_u8_to_act (x / 255, transforms.ToTensor semantics of the
    reference's loaders, mnist/train.py:100-108), host or device input, [B,1,28,28] or [B,784]."""
    import mvae_b200
    m = mvae_b200.MVAE(8, precision=precision)
    g = torch.Generator().manual_seed(2)
    u8 = torch.randint(0, 256, (37, 1, 28, 28), generator=g, dtype=torch.uint8)
    ref = u8.reshape(37, 784).float() / 255.0
    for src in (u8, u8.cuda(), u8.reshape(37, 784)):
        x = m.to_act(src)
        assert x.shape == (37, 784) and x.dtype == m.act_dtype() and x.is_cuda
        err = float((x.float().cpu() - ref).abs().max())
        assert err <= (1e-6 if precision == "tf32" else 8e-3), err
    assert float(m.to_act(torch.zeros(4, 784, dtype=torch.uint8)).float().abs().max()) == 0.0
    assert abs(float(m.to_act(torch.full((4, 784), 255, dtype=torch.uint8)).float().min()) - 1.0) <= 1e-6


def test_host_pipeline_matches_direct_steps():
    """The end-to-end path bench.py times (HostPipeline over a CUDA-graph trainer: pinned uint8 batches uploaded on a copy
    stream, converted by the staging copy into the graph's static activation buffer, losses read back on their own stream)
    against the same steps enqueued one by one from device tensors - same parameters, same Philox stream."""
    import mvae_b200
    from mvae_b200 import HostPipeline
    n, B, steps = 64, 256, 7
    state = O.perturbed_state(n, 3)
    g = torch.Generator().manual_seed(5)
    xs = [torch.randint(0, 256, (B, 784), generator=g, dtype=torch.uint8).pin_memory() for _ in range(steps)]
    ys = [torch.randint(0, 10, (B,), generator=g).pin_memory() for _ in range(steps)]

    def make(graph):
        m = mvae_b200.MVAE(n, precision="bf16", seed=7)
        m.load_state_dict(state)
        return m, mvae_b200.MVAETrainer(m, use_cuda_graph=graph)

    m1, t1 = make(True)
    got = [l.clone() for l in HostPipeline(t1).run(zip(xs, ys))]
    assert len(got) == steps
    m2, t2 = make(False)
    want = []
    for x, y in zip(xs, ys):
        l, _ = t2.step(x.cuda(), y.cuda())
        want.append(l.cpu())
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(zip(got, want)):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=3e-3, err_msg="step %d" % i)
    assert float(got[-1][:, 0].sum()) < float(got[0][:, 0].sum())          # it trains
    # (bf16 steps are not bit-reproducible - atomics - and early Adam steps turn any gradient difference into +-lr)
    assert rel_l2(m1.flat_params, m2.flat_params) < 1e-2
