"""GPU parity of the CelebA MVAE path (celeba/model.py, celeba/train.py:60-81,132-157) through the C ABI.

Operator tests compare each kernel with a plain fp32 PyTorch restatement of the same op on the same inputs; the step
tests compare the whole training step with oracle/celeba_oracle.py (pinned to the reference by tests/golden/celeba_*.npz).
Tolerances: tf32 path 2e-3 relative L2 per tensor (operands rounded to 10-bit mantissas, Swish is smooth so there is no
ReLU-mask discontinuity), bf16 path 4e-2; element-wise fp32 kernels 1e-5.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def _ops():
    import mvae_b200  # noqa: F401
    from mvae_b200 import _ops as ops, _lib
    return ops, _lib


# ----------------------------------------------------------------------------- operators
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("cfg", [(2, 8, 16, 4, 2, 1), (3, 8, 128, 4, 1, 0), (2, 10, 8, 3, 1, 1), (1, 9, 24, 5, 2, 2)])
def test_im2col_col2im_nhwc(dtype, cfg):
    ops, _ = _ops()
    B, H, Cc, k, s, p = cfg
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, H, H, Cc, generator=g).to(dtype).cuda()
    Ho = ops.out_size(H, k, s, p)
    K = k * k * Cc
    col = torch.empty(B * Ho * Ho, K, device="cuda", dtype=dtype)
    geo = ops.geometry(B, H, H, Cc, k, s, p)
    ops.im2col(geo, x, col, K)
    # torch unfold orders K as (c, kh, kw); ours is (kh, kw, c)
    ref = F.unfold(x.float().permute(0, 3, 1, 2), k, padding=p, stride=s)          # [B, C*k*k, L]
    ref = ref.view(B, Cc, k * k, Ho * Ho).permute(0, 3, 2, 1).reshape(B * Ho * Ho, K)
    assert torch.equal(col.float(), ref)
    # adjoint: col2im(col) == fold(col)
    y = torch.empty(B, H, H, Cc, device="cuda", dtype=dtype)
    ops.col2im(geo, col, K, y)
    back = F.fold(col.float().view(B, Ho * Ho, k * k, Cc).permute(0, 3, 2, 1).reshape(B, Cc * k * k, Ho * Ho), (H, H), k,
                  padding=p, stride=s).permute(0, 2, 3, 1)
    assert rel(y.float(), back) < (1e-6 if dtype == torch.float32 else 4e-3)


def test_im2col_col2im_nchw_strided():
    ops, _ = _ops()
    B, H, Cc, k, s, p = 2, 64, 3, 4, 2, 1
    x = torch.rand(B, Cc, H, H, generator=torch.Generator().manual_seed(2)).cuda()
    Ho = ops.out_size(H, k, s, p)
    K = k * k * Cc
    col = torch.empty(B * Ho * Ho, K, device="cuda", dtype=torch.bfloat16)
    geo = ops.geometry(B, H, H, Cc, k, s, p, ops.nchw_strides(Cc, H, H))
    ops.im2col(geo, x, col, K)
    ref = F.unfold(x, k, padding=p, stride=s).view(B, Cc, k * k, Ho * Ho).permute(0, 3, 2, 1).reshape(B * Ho * Ho, K)
    assert torch.equal(col.float(), ref.to(torch.bfloat16).float())
    y = torch.empty(B, Cc, H, H, device="cuda", dtype=torch.float32)
    ops.col2im(geo, col, K, y)
    back = F.fold(col.float().view(B, Ho * Ho, k * k, Cc).permute(0, 3, 2, 1).reshape(B, Cc * k * k, Ho * Ho), (H, H), k,
                  padding=p, stride=s)
    assert rel(y, back) < 1e-6


@pytest.mark.parametrize("cfg", [(8, 32, 32, 64, 4, 2, 1), (5, 8, 128, 256, 4, 1, 0), (3, 25, 32, 64, 4, 2, 1),
                                 (6, 12, 64, 40, 4, 2, 1), (2, 23, 32, 8, 5, 2, 1), (4, 16, 8, 300, 4, 2, 1)])
def test_implicit_gemm_equals_im2col_then_gemm(cfg):
    """mvae_conv_gemm (the patch matrix gathered inside the GEMM) is bit-identical to mvae_im2col + mvae_gemm: forward /
    transposed-conv dgrad form (patch matrix = A) and weight-gradient form (patch matrix = B, pixels contracted)."""
    ops, _ = _ops()
    B, H, Cc, Co, k, s, p = cfg
    g = torch.Generator().manual_seed(11)
    bf = torch.bfloat16
    x = torch.randn(B, H, H, Cc, generator=g).to(bf).cuda()
    Ho = ops.out_size(H, k, s, p)
    rows, K = B * Ho * Ho, k * k * Cc
    geo = ops.geometry(B, H, H, Cc, k, s, p)
    col = torch.empty(rows, K, device="cuda", dtype=bf)
    ops.im2col(geo, x, col, K)
    w = (torch.randn(Co, K, generator=g) / K ** 0.5).to(bf).cuda()
    ldo = (Co + 7) // 8 * 8
    for block_n, stages in ((0, 0), (64, 2), (128, 4), (32, 1)):
        y_ref = torch.zeros(rows, ldo, device="cuda", dtype=bf)
        y_imp = torch.zeros(rows, ldo, device="cuda", dtype=bf)
        ops.gemm(col, w, y_ref, rows, Co, K, K, K, ldo, block_n=block_n, stages=stages)
        ops.gemm(x, w, y_imp, rows, Co, K, 0, K, ldo, patch=(geo, 1), block_n=block_n, stages=stages)
        assert torch.equal(y_ref, y_imp), (cfg, block_n, stages, float((y_ref.float() - y_imp.float()).abs().max()))
    # weight gradient: dW[Co, K] = dy^T col, both operands with the pixel index as the contraction
    dy = torch.randn(rows, ldo, generator=g).to(bf).cuda()
    for block_n, stages, split in ((0, 0, 1), (256, 3, 1), (64, 2, 1), (0, 0, 0)):
        dw_ref = torch.zeros(Co, K, device="cuda")
        dw_imp = torch.zeros(Co, K, device="cuda")
        ops.gemm(dy, col, dw_ref, Co, K, rows, ldo, K, K, a_major=1, b_major=1, accumulate=True, split_k=split,
                 block_n=block_n, stages=stages)
        ops.gemm(dy, x, dw_imp, Co, K, rows, ldo, 0, K, a_major=1, b_major=1, accumulate=True, split_k=split,
                 block_n=block_n, stages=stages, patch=(geo, 2))
        if split == 1:
            assert torch.equal(dw_ref, dw_imp), (cfg, block_n, stages, float((dw_ref - dw_imp).abs().max()))
        else:   # split-K accumulates with atomics: order differs run to run
            assert rel(dw_imp, dw_ref) < 1e-5
    ref = dy[:, :Co].float().t() @ col.float()
    assert rel(dw_imp, ref) < 2e-3


@pytest.mark.parametrize("cfg", [(4, 8, 128, 64, 4, 2, 1), (3, 5, 256, 128, 4, 1, 0), (2, 2, 256, 128, 4, 2, 0),
                                 (2, 12, 64, 32, 5, 2, 1), (5, 16, 64, 32, 4, 2, 1)])
def test_transposed_conv_implicit_matches_torch(cfg):
    """ConvTranspose2d forward as one gather GEMM per output-parity class (no patch matrix, no col2im) against torch."""
    ops, _ = _ops()
    B, hin, ci, co, k, s, p = cfg
    g = torch.Generator().manual_seed(13)
    bf = torch.bfloat16
    x = torch.randn(B, hin, hin, ci, generator=g).to(bf).cuda()
    w = (torch.randn(ci, k, k, co, generator=g) / (ci * k) ** 0.5).to(bf).cuda()
    hout = (hin - 1) * s - 2 * p + k
    out = torch.full((B, hout, hout, co), float("nan"), device="cuda", dtype=bf)
    assert ops.transposed_conv_implicit(x, w, out, B, hin, ci, co, k, s, p) == hout
    ref = F.conv_transpose2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), stride=s, padding=p).permute(0, 2, 3, 1)
    assert not bool(torch.isnan(out.float()).any())
    assert rel(out.float(), ref) < 4e-3


@pytest.mark.parametrize("cfg", [(4, 8, 128, 64, 4, 2, 1), (3, 5, 256, 128, 4, 1, 0), (2, 2, 256, 128, 4, 2, 0),
                                 (2, 12, 64, 32, 5, 2, 1), (5, 16, 64, 32, 4, 2, 1)])
def test_transposed_conv_merged_classes_matches_torch(cfg):
    """As test_transposed_conv_implicit_matches_torch with merged=True (k5 s2 has unequal classes: exercises the fallback)."""
    ops, _ = _ops()
    B, hin, ci, co, k, s, p = cfg
    g = torch.Generator().manual_seed(14)
    bf = torch.bfloat16
    x = torch.randn(B, hin, hin, ci, generator=g).to(bf).cuda()
    w = (torch.randn(ci, k, k, co, generator=g) / (ci * k) ** 0.5).to(bf).cuda()
    hout = (hin - 1) * s - 2 * p + k
    out = torch.full((B, hout, hout, co), float("nan"), device="cuda", dtype=bf)
    ops.transposed_conv_implicit(x, w, out, B, hin, ci, co, k, s, p, merged=True)
    per_class = torch.empty_like(out)
    ops.transposed_conv_implicit(x, w, per_class, B, hin, ci, co, k, s, p)
    ref = F.conv_transpose2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), stride=s, padding=p).permute(0, 2, 3, 1)
    assert not bool(torch.isnan(out.float()).any())
    assert rel(out.float(), ref) < 4e-3
    assert torch.equal(out, per_class)      # same tiles, same accumulation order


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(3 * 40, 64, 40), (1000, 32, 1000), (257, 256, 257), (96, 1024, 32)])
def test_bn_swish_forward_backward(dtype, shape):
    ops, lib = _ops()
    rows, Cc, rpg = shape
    groups = (rows + rpg - 1) // rpg
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(rows, Cc, generator=g) * 1.5 + 0.3).to(dtype).cuda()
    gamma = (1 + 0.2 * torch.randn(Cc, generator=g)).cuda()
    beta = (0.1 * torch.randn(Cc, generator=g)).cuda()
    dy = torch.randn(rows, Cc, generator=g).to(dtype).cuda()
    rm, rv = torch.zeros(Cc, device="cuda"), torch.ones(Cc, device="cuda")
    f = lambda *s: torch.zeros(*s, device="cuda")
    keep = [f(groups, Cc) for _ in range(6)]   # the argument struct holds raw pointers: keep the tensors alive
    a = ops.bn_args(x, rows, Cc, rpg, lib.ACT_SWISH, True, gamma, beta, keep[0], keep[1], keep[2], keep[3], rm, rv, updates=2)
    y = torch.empty_like(x)
    ops.bn_act_forward(a, y)
    dx = torch.empty_like(x)
    dgamma, dbeta = f(Cc), f(Cc)
    ops.bn_act_backward(a, dy, dx, keep[4], keep[5], dgamma, dbeta)
    # fp32 torch restatement on the same (stored) inputs
    xr = x.float().clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    outs = []
    rm_r, rv_r = torch.zeros(Cc, device="cuda"), torch.ones(Cc, device="cuda")
    for gi in range(groups):
        xs = xr[gi * rpg:(gi + 1) * rpg]
        m, v = xs.mean(0), xs.var(0, unbiased=False)
        u = (xs - m) / torch.sqrt(v + 1e-5) * gr + br
        outs.append(u * torch.sigmoid(u))
        for _ in range(2):
            rm_r = 0.9 * rm_r + 0.1 * m.detach()
            rv_r = 0.9 * rv_r + 0.1 * (v.detach() * xs.shape[0] / max(xs.shape[0] - 1, 1))
    yr = torch.cat(outs)
    yr.backward(dy.float())
    tol = 2e-5 if dtype == torch.float32 else 6e-3
    assert rel(y.float(), yr) < tol
    assert rel(dx.float(), xr.grad) < (2e-4 if dtype == torch.float32 else 1.5e-2)
    assert rel(dgamma, gr.grad) < 2e-4 and rel(dbeta, br.grad) < 2e-4
    assert rel(rm, rm_r) < 1e-5 and rel(rv, rv_r) < 1e-5
    # eval mode
    a.training = 0
    ops.bn_act_forward(a, y)
    u = (x.float() - rm) / torch.sqrt(rv + 1e-5) * gamma + beta
    assert rel(y.float(), u * torch.sigmoid(u)) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_swish_dropout_forward_backward(dtype):
    ops, lib = _ops()
    rows, Cc, R, p = 64, 1024, 2, 0.1
    g = torch.Generator().manual_seed(4)
    x = torch.randn(rows, Cc, generator=g).to(dtype).cuda()
    step = torch.tensor([7], device="cuda", dtype=torch.int32)
    y = torch.empty(R * rows, Cc, device="cuda", dtype=dtype)
    ops.act_forward(lib.ACT_SWISH, x, y, rows, Cc, repeat=R, dropout_p=p, seed=11, step_counter=step)
    sw = x.float() * torch.sigmoid(x.float())
    yf = y.float().view(R, rows, Cc)
    keep = yf != 0
    frac = float(keep.float().mean())
    assert abs(frac - (1 - p)) < 0.01                        # keep probability
    assert not torch.equal(keep[0], keep[1])                 # independent masks per replica
    scale = 65536.0 / round((1 - p) * 65536)
    assert rel(yf[keep], (sw.unsqueeze(0).expand(R, -1, -1) * scale)[keep]) < (1e-5 if dtype == torch.float32 else 5e-3)
    dy = torch.randn(R * rows, Cc, generator=g).to(dtype).cuda()
    dx = torch.empty(rows, Cc, device="cuda", dtype=dtype)
    dbias = torch.zeros(Cc, device="cuda")
    ops.act_backward(lib.ACT_SWISH, x, dy, dx, rows, Cc, repeat=R, dropout_p=p, seed=11, step_counter=step, dbias=dbias)
    s = torch.sigmoid(x.float())
    dact = s * (1 + x.float() * (1 - s))
    ref = (dy.float().view(R, rows, Cc) * keep.float() * scale).sum(0) * dact
    assert rel(dx.float(), ref) < (1e-5 if dtype == torch.float32 else 6e-3)
    assert rel(dbias, ref.sum(0)) < (1e-4 if dtype == torch.float32 else 6e-3)
    # no dropout, single replica == plain swish
    ops.act_forward(lib.ACT_SWISH, x, y, rows, Cc)
    assert rel(y[:rows].float(), sw) < (1e-6 if dtype == torch.float32 else 5e-3)


def test_sigmoid_bce_value_and_gradient():
    ops, _ = _ops()
    B, G, cols = 5, 3, 777
    g = torch.Generator().manual_seed(5)
    logits = (torch.randn(G * B, cols, generator=g) * 3).cuda()
    target = torch.rand(B, cols, generator=g).cuda()
    loss = torch.zeros(4, device="cuda")
    probs = torch.empty_like(logits)
    dl = torch.empty_like(logits)
    scale = (0.5, 1.0, 2.0)
    ops.sigmoid_bce(logits, cols, G * B, cols, rows_per_group=B, target=target, ld_target=cols, target_rows=B, grad_scale=scale,
                    loss=loss, probs=probs, ld_probs=cols, dlogits=dl, ld_dlogits=cols)
    lr = logits.clone().requires_grad_(True)
    tot = 0
    for gi in range(G):
        pr = torch.sigmoid(lr[gi * B:(gi + 1) * B])
        l = F.binary_cross_entropy(pr, target, reduction="sum")
        assert abs(float(loss[gi]) - float(l.detach())) < 2e-5 * float(l.detach())
        tot = tot + scale[gi] * l
    tot.backward()
    assert rel(probs, torch.sigmoid(logits)) < 1e-6
    assert rel(dl, lr.grad) < 1e-5
    # padded bf16 gradient rows (attribute decoder): columns >= cols are zero
    small = (torch.randn(7, 18, generator=g)).cuda()
    tgt = (torch.rand(7, 18, generator=g) > 0.5).float().cuda()
    d16 = torch.full((7, 24), 5.0, device="cuda", dtype=torch.bfloat16)
    ops.sigmoid_bce(small, 18, 7, 18, target=tgt, ld_target=18, target_rows=7, grad_scale=(1.0, 0, 0), loss=loss, dlogits=d16,
                    ld_dlogits=24)
    assert float(d16[:, 18:].abs().max()) == 0.0
    assert rel(d16[:, :18].float(), torch.sigmoid(small) - tgt) < 5e-3


def test_gemm_conv_shapes_against_torch():
    """The operand-major / leading-dimension combinations the conv path uses (bf16)."""
    ops, _ = _ops()
    g = torch.Generator().manual_seed(6)
    bf = torch.bfloat16
    # conv forward: K-major both, K = 48 (< one 64-wide k block), N = 32
    A = torch.randn(2048, 48, generator=g).to(bf).cuda()
    W = torch.randn(32, 48, generator=g).to(bf).cuda()
    out = torch.empty(2048, 32, device="cuda", dtype=bf)
    ops.gemm(A, W, out, 2048, 32, 48, 48, 48, 32)
    assert rel(out.float(), A.float() @ W.float().t()) < 4e-3
    # transposed-conv forward: B stored [K, N] (b_major = 1), N = 2048
    X = torch.randn(75, 256, generator=g).to(bf).cuda()
    Wt = torch.randn(256, 2048, generator=g).to(bf).cuda()
    col = torch.empty(75, 2048, device="cuda", dtype=bf)
    ops.gemm(X, Wt, col, 75, 2048, 256, 256, 2048, 2048, b_major=1)
    assert rel(col.float(), X.float() @ Wt.float()) < 4e-3
    # weight gradient: both operands row-index-contiguous, tiny M (3 output channels would be 32 here), long K
    dY = torch.randn(6144, 32, generator=g).to(bf).cuda()
    C2 = torch.randn(6144, 48, generator=g).to(bf).cuda()
    dW = torch.zeros(32, 48, device="cuda")
    ops.gemm(dY, C2, dW, 32, 48, 6144, 32, 48, 48, a_major=1, b_major=1, accumulate=True)
    assert rel(dW, dY.float().t() @ C2.float()) < 2e-3
    # padded leading dimensions: K = 18 inside ld = 24, N = 100 inside ld = 104
    At = torch.zeros(40, 24, device="cuda", dtype=bf)
    At[:, :18] = torch.randn(40, 18, generator=g).to(bf).cuda()
    Wp = torch.zeros(64, 24, device="cuda", dtype=bf)
    Wp[:, :18] = torch.randn(64, 18, generator=g).to(bf).cuda()
    o = torch.empty(40, 64, device="cuda")
    ops.gemm(At, Wp, o, 40, 64, 18, 24, 24, 64)
    assert rel(o, At.float() @ Wp.float().t()) < 2e-3
    Wz = torch.zeros(64, 104, device="cuda", dtype=bf)
    Wz[:, :100] = torch.randn(64, 100, generator=g).to(bf).cuda()
    dH = torch.randn(40, 64, generator=g).to(bf).cuda()
    dz = torch.empty(40, 100, device="cuda")
    ops.gemm(dH, Wz, dz, 40, 100, 64, 64, 104, 100, b_major=1)
    assert rel(dz, dH.float() @ Wz.float()[:, :100]) < 2e-3


# ----------------------------------------------------------------------------- whole step
def _device_step(precision, B, n, seed, dropout_p=0.0, graph=False, adam=False):
    import celeba_oracle as O
    from mvae_b200.celeba import MultimodalVAE, CelebATrainer
    state = O.init_state(n, seed=1234 + seed)
    image, attrs, noises = O.synthetic_batch(B, n, seed)
    m = MultimodalVAE(n_latents=n, precision=precision, dropout_p=dropout_p)
    m.load_state_dict(state)
    tr = CelebATrainer(m, use_cuda_graph=graph)
    eps = torch.stack(noises).cuda()
    tr.step(image.cuda(), attrs.cuda(), eps=eps, adam=adam)
    torch.cuda.synchronize()
    return O, m, tr, state, image, attrs, noises


@pytest.mark.parametrize("precision,tol_out,tol_grad", [("tf32", 1e-3, 3e-3), ("bf16", 2e-2, 6e-2)])
@pytest.mark.parametrize("B,n,seed", [(8, 16, 2), (16, 100, 0)])
def test_step_matches_oracle(precision, tol_out, tol_grad, B, n, seed):
    O, m, tr, state, image, attrs, noises = _device_step(precision, B, n, seed)
    losses, grads, bufs, outs = O.train_step(state, image, attrs, noises)
    dev_losses = tr.losses()
    for g in range(3):
        assert abs(dev_losses[g][0] - losses[g]) <= tol_out * abs(losses[g]), (g, dev_losses[g], losses[g])
    ws = m.workspace(B, 3)
    mu = ws.mu.view(3, B, n)
    lv = ws.logvar.view(3, B, n)
    for g in range(3):
        assert rel(mu[g], outs[g][2]) < tol_out * 3
        assert rel(lv[g], outs[g][3]) < tol_out * 3
    dg = m.grads_reference()
    worst = {}
    for k, v in grads.items():
        if float(v.abs().max()) < 1e-7:       # biases feeding a train-mode BatchNorm: true gradient 0
            assert float(dg[k].abs().max()) < 1e-5, k
            continue
        worst[k] = rel(dg[k], v)
    bad = {k: e for k, e in worst.items() if e > tol_grad}
    assert not bad, bad
    sd = m.state_dict()
    for k, v in bufs.items():
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == int(v), k
        else:
            assert rel(sd[k], v) < 5 * tol_out, k


def test_step_matches_reference_fixture_tf32():
    """Device (tf32) against the fixture generated from the real reference classes (oracle/gen_golden_celeba.py)."""
    import celeba_oracle as O
    g = np.load(os.path.join(GOLD, "celeba_b16_n100.npz"))
    B, n, seed = int(g["batch"]), int(g["n_latents"]), int(g["seed"])
    _, m, tr, *_ = _device_step("tf32", B, n, seed)
    dev = tr.losses()
    np.testing.assert_allclose([d[0] for d in dev], g["losses"], rtol=1e-3)
    dg = m.grads_reference()
    for k in dg:
        ref = torch.from_numpy(g["gradsample/" + k])
        if float(ref.abs().max()) < 1e-7:
            continue
        assert rel(O.sample_flat(dg[k].cpu()), ref) < 4e-3, k


def test_forward_surface_eval_and_state_dict_roundtrip():
    import celeba_oracle as O
    from mvae_b200.celeba import MultimodalVAE
    n, B = 16, 6
    state = O.init_state(n, seed=77)
    for k in state:                      # non-trivial running statistics
        if k.endswith("running_mean"):
            state[k] = 0.1 * torch.randn(state[k].shape, generator=torch.Generator().manual_seed(1))
        if k.endswith("running_var"):
            state[k] = 0.5 + torch.rand(state[k].shape, generator=torch.Generator().manual_seed(2))
    m = MultimodalVAE(n_latents=n, precision="tf32")
    m.load_state_dict(state)
    sd = m.state_dict()
    assert list(sd.keys()) == list(state.keys())
    for k in state:
        assert torch.equal(sd[k].cpu(), state[k]), k
    image, attrs, _ = O.synthetic_batch(B, n, 3)
    m.eval()
    for kw in (dict(image=image, attrs=attrs), dict(image=image), dict(attrs=attrs)):
        ri, ra, mu, lv = m(**{k: v.cuda() for k, v in kw.items()})
        ref = O.forward(state, kw.get("image"), kw.get("attrs"), None, None, training=False)
        assert ri.shape == (B, 3, 64, 64) and ra.shape == (B, 18)
        assert rel(ri, ref[0]) < 2e-3 and rel(ra, ref[1]) < 2e-3
        assert rel(mu, ref[2]) < 2e-3 and rel(lv, ref[3]) < 2e-3
    z = torch.randn(B, n, generator=torch.Generator().manual_seed(4))
    assert rel(m.image_decoder(z.cuda()), torch.sigmoid(O.image_decoder_logits(state, z, None, False))) < 2e-3


def test_graph_replay_and_adam_decrease_loss_bf16():
    """CUDA-graph replay of the whole step (dropout on, in-kernel noise): the ELBO goes down and stays finite."""
    import celeba_oracle as O
    from mvae_b200.celeba import MultimodalVAE, CelebATrainer
    n, B = 32, 32
    m = MultimodalVAE(n_latents=n, precision="bf16", dropout_p=0.1)
    tr = CelebATrainer(m, use_cuda_graph=True)
    image, attrs, _ = O.synthetic_batch(B, n, 1)
    image, attrs = image.cuda(), attrs.cuda()
    hist = []
    for it in range(60):
        tr.step(image, attrs)
        if it % 10 == 9 or it == 0:
            hist.append(sum(l[0] for l in tr.losses()))
    assert all(np.isfinite(hist)), hist
    assert hist[-1] < hist[0] - 0.05, hist
    assert int(m.state_dict()["image_decoder.hallucinate.1.num_batches_tracked"]) == 180
    assert int(m.state_dict()["image_encoder.features.3.num_batches_tracked"]) == 120


def test_bf16_training_curve_tracks_fp32_oracle():
    """north_star: the bf16 path's ELBO curve stays within 1% of the fp32 reference run on the same data, weights and
    noise.  150 Adam steps (celeba/train.py:132-157 semantics) on one fixed batch; device bf16 vs CPU oracle fp32."""
    import celeba_oracle as O
    import mnist_oracle as MN
    from mvae_b200.celeba import MultimodalVAE, CelebATrainer
    n, B, steps = 16, 16, 150
    state = O.init_state(n, seed=99)
    image, attrs, _ = O.synthetic_batch(B, n, 5)
    g = torch.Generator().manual_seed(123)
    noise = [torch.randn(3, B, n, generator=g) for _ in range(steps)]
    m = MultimodalVAE(n_latents=n, precision="bf16", dropout_p=0.0)
    m.load_state_dict(state)
    tr = CelebATrainer(m, lr=1e-3)
    dev_curve = []
    img_d, att_d = image.cuda(), attrs.cuda()
    for it in range(steps):
        tr.step(img_d, att_d, eps=noise[it].cuda())
        if it % 10 == 9:
            dev_curve.append(sum(l[0] for l in tr.losses()))
    p = {k: v.clone() for k, v in state.items()}
    mom = {k: torch.zeros_like(v) for k, v in p.items() if not O.is_buffer(k)}
    vel = {k: torch.zeros_like(v) for k, v in p.items() if not O.is_buffer(k)}
    ref_curve = []
    for it in range(steps):
        losses, grads, bufs, _ = O.train_step(p, image, attrs, list(noise[it]))
        p = MN.adam_step(p, grads, mom, vel, it + 1)
        p.update(bufs)
        if it % 10 == 9:
            ref_curve.append(sum(losses))
    dev_curve, ref_curve = np.array(dev_curve), np.array(ref_curve)
    assert ref_curve[-1] < ref_curve[0] - 0.05, ref_curve          # it does train
    rel_err = np.abs(dev_curve - ref_curve) / np.abs(ref_curve)
    assert rel_err.max() < 0.01, (rel_err, dev_curve, ref_curve)


def test_reference_training_loop_with_autograd():
    """The reference's own loop (celeba/train.py:138-152): three vae(...) calls, three loss_function calls, loss.backward(),
    through the module surface with autograd; gradients against the oracle, then one torch.optim.Adam step."""
    import celeba_oracle as O
    from mvae_b200.celeba import MultimodalVAE, loss_function
    n, B, seed = 16, 8, 2
    state = O.init_state(n, seed=1234 + seed)
    image, attrs, noises = O.synthetic_batch(B, n, seed)
    ref_losses, ref_grads, ref_bufs, _ = O.train_step(state, image, attrs, noises)
    vae = MultimodalVAE(n_latents=n, precision="tf32", dropout_p=0.0)
    vae.load_state_dict(state)
    vae.train()
    opt = torch.optim.Adam(vae.parameters(), lr=1e-3)
    vae.zero_grad()
    img, att = image.cuda(), attrs.cuda()
    r1 = vae(image=img, attrs=att, eps=noises[0])
    r2 = vae(image=img, eps=noises[1])
    r3 = vae(attrs=att, eps=noises[2])
    losses = [loss_function(r[2], r[3], recon_x=r[0], x=img, recon_y=r[1], y=att) for r in (r1, r2, r3)]
    for a, b in zip(losses, ref_losses):
        assert abs(float(a.detach()) - b) <= 1e-3 * abs(b)
    (losses[0] + losses[1] + losses[2]).backward()
    dg = vae.grads_reference()
    for k, v in ref_grads.items():
        if float(v.abs().max()) < 1e-7:
            continue
        assert rel(dg[k], v) < 4e-3, (k, rel(dg[k], v))
    sd = vae.state_dict()
    assert int(sd["image_encoder.features.3.num_batches_tracked"]) == 2 and int(sd["attrs_decoder.net.1.num_batches_tracked"]) == 3
    for k, v in ref_bufs.items():
        if not k.endswith("num_batches_tracked"):
            assert rel(sd[k], v) < 5e-3, k
    before = vae.flat_params.clone()
    opt.step()
    assert float((vae.flat_params - before).abs().max()) > 1e-4       # torch.optim updates the flat leaf in place


def test_autograd_with_optimizer_zero_grad_set_to_none():
    """optimizer.zero_grad() (set_to_none) detaches param.grad from the library's buffer: autograd then delivers the gradient."""
    import celeba_oracle as O
    from mvae_b200.celeba import MultimodalVAE, loss_function
    n, B = 16, 4
    state = O.init_state(n, seed=7)
    image, attrs, noises = O.synthetic_batch(B, n, 3)
    vae = MultimodalVAE(n_latents=n, precision="tf32", dropout_p=0.0)
    vae.load_state_dict(state)
    opt = torch.optim.SGD(vae.parameters(), lr=0.1)
    opt.zero_grad()
    ri, ra, mu, lv = vae(image=image.cuda(), attrs=attrs.cuda(), eps=noises[0])
    loss_function(mu, lv, recon_x=ri, x=image.cuda(), recon_y=ra, y=attrs.cuda()).backward()
    g = vae.param.grad
    assert g is not None and g.data_ptr() != vae.flat_grads.data_ptr() and float(g.abs().sum()) > 0
    work = {k: (v.clone() if O.is_buffer(k) else v.clone().requires_grad_(True)) for k, v in state.items()}
    r = O.forward(work, image, attrs, noises[0], work, True)
    O.loss_function(r[2], r[3], r[0], image, r[1], attrs).backward()
    key = "image_decoder.hallucinate.0.weight"
    l = vae.layouts[key]
    assert rel(l.to_reference(g[l.offset:l.offset + l.numel]), work[key].grad) < 4e-3


@pytest.mark.parametrize("precision,tol", [("tf32", 4e-3), ("bf16", 8e-2)])
def test_ragged_sizes_match_oracle(precision, tol):
    """Batch and latent sizes that are not multiples of any tile / vector width (B = 5, n = 10: padded leading dimensions,
    partial 128-row tiles everywhere)."""
    O, m, tr, state, image, attrs, noises = _device_step(precision, 5, 10, 4)
    losses, grads, _, _ = O.train_step(state, image, attrs, noises)
    for a, b in zip(tr.losses(), losses):
        assert abs(a[0] - b) <= (2e-3 if precision == "tf32" else 3e-2) * abs(b)
    dg = m.grads_reference()
    bad = {k: rel(dg[k], v) for k, v in grads.items() if float(v.abs().max()) > 1e-7 and rel(dg[k], v) > tol}
    assert not bad, bad
