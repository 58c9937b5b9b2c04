"""Per-sample missing-modality masks for the conv trainers (multimnist/paired_weak.py:85-110 applied per row instead
of per batch): ConvMVAETrainer.step_masked must equal the oracle fed the paired / image-only / other-only subsets as three
batches, with the gradients summed.  tf32 path; tolerances as in the step tests (losses 2e-3, gradients 1e-2 relative L2)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("family", ["celeba", "multimnist"])
def test_step_masked_matches_oracle_on_the_three_subsets(family):
    if family == "celeba":
        import celeba_oracle as O
        from mvae_b200.celeba import MultimodalVAE, CelebATrainer as Trainer
        kw = lambda lx, ly: dict(lambda_x=lx, lambda_y=ly)
        names = ("recon_x", "x", "recon_y", "y")
    else:
        import multimnist_oracle as O
        from mvae_b200.multimnist import MultimodalVAE, MultiMNISTTrainer as Trainer
        kw = lambda lx, ly: dict(lambda_xy=lx, lambda_yx=ly)
        names = ("recon_image", "image", "recon_text", "text")
    B, n, seed = 24, 16, 4
    state = O.init_state(n, seed=1234 + seed)
    image, other, noises = O.synthetic_batch(B, n, seed)
    g = torch.Generator().manual_seed(7)
    has_image = torch.rand(B, generator=g) < 0.7
    has_other = torch.rand(B, generator=g) < 0.6
    has_image[:2], has_other[:2] = False, False          # rows with neither modality are skipped
    has_image[2:6], has_other[2:6] = True, True
    has_image[6:9], has_other[6:9] = True, False
    has_image[9:12], has_other[9:12] = False, True
    classes = [("paired", has_image & has_other, (0, 1, 2), ((1., 1.), (1., 1.), (0., 1.))),
               ("image_only", has_image & ~has_other, (1,), ((1., 0.),)),
               ("other_only", ~has_image & has_other, (2,), ((0., 1.),))]
    work = {k: (v.clone() if O.is_buffer(k) else v.detach().clone().requires_grad_(True)) for k, v in state.items()}
    total, ref = 0.0, {}
    for cname, rows, terms, lambdas in classes:
        idx = torch.nonzero(rows).reshape(-1)
        assert idx.numel() >= 2
        im, ot = image[idx], other[idx]
        ref[cname] = []
        for t, (lx, ly) in zip(terms, lambdas):
            out = O.forward(work, im if t != 2 else None, ot if t != 1 else None, noises[t][idx], work, True)
            args = {names[0]: out[0], names[1]: im, names[2]: out[1], names[3]: ot}
            l = O.loss_function(out[2], out[3], **args, **kw(lx, ly))
            ref[cname].append(float(l.detach()))
            total = total + l
    keys = [k for k in work if not O.is_buffer(k)]
    gs = torch.autograd.grad(total, [work[k] for k in keys], allow_unused=True)
    grads = {k: (torch.zeros_like(work[k]) if gg is None else gg) for k, gg in zip(keys, gs)}

    m = MultimodalVAE(n_latents=n, precision="tf32", dropout_p=0.0)
    m.load_state_dict(state)
    tr = Trainer(m)
    out = tr.step_masked(image.cuda(), other.cuda(), has_image, has_other, eps=torch.stack(noises).cuda(), update=False)
    torch.cuda.synchronize()
    assert set(out) == set(ref)
    for cname in ref:
        got = [l[0] for l in out[cname]]
        for a, b in zip(got, ref[cname]):
            assert abs(a - b) <= 2e-3 * abs(b), (cname, got, ref[cname])
    dg = m.grads_reference()
    bad = {}
    for k, r in grads.items():
        if float(r.abs().max()) < 1e-7:
            continue
        e = rel(dg[k], r)
        if e > 1e-2:
            bad[k] = e
    assert not bad, bad
    # one Adam update moves the parameters and ticks the optimizer clock once
    before = m.flat_params.clone()
    tr.step_masked(image.cuda(), other.cuda(), has_image, has_other, eps=torch.stack(noises).cuda())
    torch.cuda.synchronize()
    assert int(m._adam_counter) == 1 and float((m.flat_params - before).abs().max()) > 0
