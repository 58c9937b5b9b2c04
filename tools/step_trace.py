"""Bring-up / debugging aid: every intermediate tensor of one fused MNIST step on the device (forward activations AND
backward gradients, read out of the step workspace) against an fp64 autograd recomputation of the reference's step
(mnist/model.py:53-185, mnist/train.py:64-81,132-148) on the CPU.

    python tools/step_trace.py B [precision] [n_latents] [seed]      e.g.  python tools/step_trace.py 4096 bf16

Prints one line per buffer: relative L2 error and the norm.  Also used to A/B two device paths (MVAE_CHAIN=0/1).
"""
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
warnings.filterwarnings("ignore")

import torch  # noqa: E402

import mnist_oracle as O  # noqa: E402


def rel(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def bn_train(x, gamma, beta):
    mean = x.mean(0)
    var = x.var(0, unbiased=False)
    return gamma * (x - mean) / torch.sqrt(var + 1e-5) + beta


def reference_trace(state, image, text, noises):
    """fp64 forward + backward with every intermediate kept (retain_grad)."""
    p = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() and not O.is_buffer(k) else v)
         for k, v in state.items()}
    x = image.double()
    keep = {}

    def K(name, t):
        t.retain_grad()
        keep[name] = t
        return t

    h1pre = K("h1pre", x @ p["image_encoder.net.0.weight"].t() + p["image_encoder.net.0.bias"])
    h1 = K("h1", torch.relu(bn_train(h1pre, p["image_encoder.net.1.weight"], p["image_encoder.net.1.bias"])))
    h2pre = K("h2pre", h1 @ p["image_encoder.net.3.weight"].t() + p["image_encoder.net.3.bias"])
    h2 = K("h2", torch.relu(bn_train(h2pre, p["image_encoder.net.4.weight"], p["image_encoder.net.4.bias"])))
    enc = K("enc", h2 @ p["image_encoder.net.6.weight"].t() + p["image_encoder.net.6.bias"])
    n = enc.shape[1] // 2
    e = p["text_encoder.net.0.weight"][text]
    te = torch.relu(bn_train(e, p["text_encoder.net.1.weight"], p["text_encoder.net.1.bias"]))
    tenc = K("tenc", te @ p["text_encoder.net.3.weight"].t() + p["text_encoder.net.3.bias"])
    B = x.shape[0]
    zs, g1pres, g1s, g2pres, g2s, logits, t1pres, mus, lvs = [], [], [], [], [], [], [], [], []
    total = 0.0
    losses = []
    for k in range(3):
        ms, ls = [], []
        if k != 2:
            ms.append(enc[:, :n]); ls.append(enc[:, n:])
        if k != 1:
            ms.append(tenc[:, :n]); ls.append(tenc[:, n:])
        mu, lv = O.product_of_experts(torch.stack(ms), torch.stack(ls))
        z = K("z%d" % k, noises[k].double() * torch.exp(0.5 * lv) + mu)
        g1pre = K("g1pre%d" % k, z @ p["image_decoder.net.0.weight"].t() + p["image_decoder.net.0.bias"])
        g1 = K("g1%d" % k, torch.relu(bn_train(g1pre, p["image_decoder.net.1.weight"], p["image_decoder.net.1.bias"])))
        g2pre = K("g2pre%d" % k, g1 @ p["image_decoder.net.3.weight"].t() + p["image_decoder.net.3.bias"])
        g2 = K("g2%d" % k, torch.relu(bn_train(g2pre, p["image_decoder.net.4.weight"], p["image_decoder.net.4.bias"])))
        lg = K("logits%d" % k, g2 @ p["image_decoder.net.6.weight"].t() + p["image_decoder.net.6.bias"])
        t1pre = K("t1pre%d" % k, z @ p["text_decoder.net.0.weight"].t() + p["text_decoder.net.0.bias"])
        t1 = torch.relu(bn_train(t1pre, p["text_decoder.net.1.weight"], p["text_decoder.net.1.bias"]))
        tl = t1 @ p["text_decoder.net.3.weight"].t() + p["text_decoder.net.3.bias"]
        bce = torch.nn.functional.binary_cross_entropy_with_logits(lg, x, reduction="mean")
        ce = torch.nn.functional.cross_entropy(tl, text, reduction="mean")
        kld = -0.5 * torch.sum(1 + lv - mu.pow(2) - lv.exp()) / (B * (784 / 3))
        loss = bce + ce + kld
        losses.append(float(loss))
        total = total + loss
        mus.append(mu); lvs.append(lv)
    total.backward()
    cat = lambda pre: torch.cat([keep["%s%d" % (pre, k)] for k in range(3)])
    catg = lambda pre: torch.cat([keep["%s%d" % (pre, k)].grad for k in range(3)])
    fwd = {"h1pre": keep["h1pre"], "h1": keep["h1"], "h2pre": keep["h2pre"], "h2": keep["h2"], "enc": keep["enc"],
           "z": cat("z"), "g1pre": cat("g1pre"), "g1": cat("g1"), "g2pre": cat("g2pre"), "g2": cat("g2"), "t1pre": cat("t1pre")}
    dz_img = catg("g1pre") @ p["image_decoder.net.0.weight"].detach()
    bwd = {"dlog": catg("logits"), "dy2": catg("g2pre"), "dy1": catg("g1pre"), "dz": dz_img, "denc": keep["enc"].grad,
           "dye2": keep["h2pre"].grad, "dye1": keep["h1pre"].grad}
    grads = {k: v.grad for k, v in p.items() if torch.is_tensor(v) and v.requires_grad}
    return fwd, bwd, grads, losses


def device_buffers(m, B):
    n, R = m.n_latents, 3 * B
    f32 = torch.float32
    buf = lambda name, shape, dt=None: m.debug_buffer(name, B, shape, dt).float().cpu().clone()
    fwd = {"h1pre": buf("h1pre", (B, 400)), "h1": buf("h1", (B, 400)), "h2pre": buf("h2pre", (B, 200)), "h2": buf("h2", (B, 200)),
           "enc": buf("enc", (B, 2 * n), f32), "z": buf("z", (R, n)), "g1pre": buf("g1pre", (R, 200)), "g1": buf("g1", (R, 200)),
           "g2pre": buf("g2pre", (R, 400)), "g2": buf("g2", (R, 400)), "t1pre": buf("t1pre", (R, 10), f32)}
    bwd = {"dlog": buf("dlog", (R, 784)), "dy2": buf("dy2", (R, 400)), "dy1": buf("dy1", (R, 200)), "dz": buf("dz", (R, n), f32),
           "denc": buf("denc", (B, 2 * n)), "dye2": buf("dye2", (B, 200)), "dye1": buf("dye1", (B, 400))}
    return fwd, bwd


def run_device(state, image, text, noises, n, precision):
    import mvae_b200
    m = mvae_b200.MVAE(n, precision=precision)
    m.load_state_dict(state)
    tr = mvae_b200.MVAETrainer(m)
    dl, _ = tr.step(image.cuda(), text.cuda(), eps=torch.stack(noises).cuda(), update=False)
    torch.cuda.synchronize()
    B = image.shape[0]
    try:
        bars = m.debug_buffer("bars", B, (32,), torch.int32).cpu().tolist()
        print("  chain barrier counters %s  time-out flag 0x%x" % (bars[:16], bars[31] & 0xffffffff))
    except Exception as e:  # noqa: BLE001
        print("  (no barrier counters: %s)" % e)
    fwd, bwd = device_buffers(m, B)
    grads = {k: p.grad.detach().float().cpu().clone() for k, p in m.named_parameters()}
    return fwd, bwd, grads, dl[:, 0].cpu().tolist()


def main():
    B = int(sys.argv[1])
    precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    seed = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    state = O.perturbed_state(n, seed)
    image, text, noises = O.synthetic_batch(B, n, seed)
    rf, rb, rg, rl = reference_trace(state, image, text, noises)
    df, db, dg, dl = run_device(state, image, text, noises, n, precision)
    print("B=%d precision=%s n=%d   losses dev %s ref %s" % (B, precision, n, ["%.6f" % v for v in dl], ["%.6f" % v for v in rl]))
    for name in rf:
        print("  fwd %-6s rel %.3e   |ref| %.3e" % (name, rel(df[name], rf[name]), float(rf[name].norm())))
    for name in rb:
        print("  bwd %-6s rel %.3e   |ref| %.3e" % (name, rel(db[name], rb[name]), float(rb[name].norm())))
    for name in rg:
        if name in O.PRE_BN_BIASES:
            continue
        print("  grad %-30s rel %.3e" % (name, rel(dg[name], rg[name])))


if __name__ == "__main__":
    main()
