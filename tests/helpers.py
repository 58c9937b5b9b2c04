"""Shared helpers of the GPU parity tests."""
import torch

import mnist_oracle as O

NAMES = ("joint", "image", "text")


def rel_l2(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def device_forward_override(m, B, text, terms=NAMES):
    """{(tag, term): tensor} of the DEVICE's forward intermediates of the last step, for oracle.FORWARD_OVERRIDE."""
    n = m.n_latents
    G = len(terms)
    R = G * B
    f32 = torch.float32
    buf = lambda name, shape, dt=None: m.debug_buffer(name, B, shape, dt).float().cpu().clone()
    ov = {}
    h1pre, h2pre = buf("h1pre", (B, 400)), buf("h2pre", (B, 200))
    enc = buf("enc", (B, 2 * n), f32)
    table = buf("txt_table", (10, 2 * n), f32)
    z, g1pre, g2pre = buf("z", (R, n)), buf("g1pre", (R, 200)), buf("g2pre", (R, 400))
    t1pre = buf("t1pre", (R, 10), f32)
    for g, t in enumerate(terms):
        k = NAMES.index(t)
        sl = slice(g * B, (g + 1) * B)
        if t != "text":
            ov[("image_encoder.net.0", k)] = h1pre
            ov[("image_encoder.net.3", k)] = h2pre
            ov[("image_encoder.net.6", k)] = enc
        if t != "image":
            ov[("text_encoder.net.3", k)] = table[text]
        ov[("z", k)] = z[sl]
        ov[("image_decoder.net.0", k)] = g1pre[sl]
        ov[("image_decoder.net.3", k)] = g2pre[sl]
        ov[("text_decoder.net.0", k)] = t1pre[sl]
    return ov


def oracle_step(state, image, text, noises, terms=NAMES, lambdas=((1., 1.),) * 3, emulate=None, override=None):
    mask = tuple(t in terms for t in NAMES)
    full = [(0., 0.)] * 3
    k = 0
    for i in range(3):
        if mask[i]:
            full[i] = lambdas[k]
            k += 1
    O.MATMUL_EMULATION = emulate
    O.FORWARD_OVERRIDE = override
    try:
        return O.train_step(state, image, text, noises, tuple(full), mask)
    finally:
        O.MATMUL_EMULATION = None
        O.FORWARD_OVERRIDE = None


def run_device_step(state, image, text, noises, n, precision, terms=NAMES, lambdas=((1., 1.),) * 3, update=False):
    import mvae_b200
    m = mvae_b200.MVAE(n, precision=precision)
    m.load_state_dict(state)
    tr = mvae_b200.MVAETrainer(m)
    idx = [NAMES.index(t) for t in terms]
    eps = torch.stack([noises[i] for i in idx]).cuda()
    losses, outs = tr.step(image.cuda(), text.cuda(), eps=eps, terms=terms, lambdas=lambdas, update=update, outputs=True)
    torch.cuda.synchronize()
    return m, tr, losses.cpu(), outs
