"""GPU bring-up: fused MNIST step vs the CPU oracle, per-tensor error report (run on the B200 box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import mnist_oracle as O
import mvae_b200
from mvae_b200 import MVAE, MVAETrainer


def rel(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    d = (a - b).abs().max().item(); s = b.abs().max().item()
    l2 = ((a - b).norm() / (b.norm() + 1e-30)).item()
    return d, s, l2


def run(B, n, precision, terms=("joint", "image", "text"), lambdas=((1., 1.),) * 3, seed=0):
    state = O.perturbed_state(n, seed)
    image, text, noises = O.synthetic_batch(B, n, seed)
    tmask = tuple(t in terms for t in ("joint", "image", "text"))
    full_l = [(0., 0.)] * 3; k = 0
    for i, t in enumerate(("joint", "image", "text")):
        if tmask[i]:
            full_l[i] = lambdas[k]; k += 1
    losses, grads, bufs, outs = O.train_step(state, image, text, noises, tuple(full_l), tmask)
    m = MVAE(n, precision=precision)
    m.load_state_dict(state)
    tr = MVAETrainer(m)
    eps = torch.stack([noises[i] for i in range(3) if tmask[i]]).cuda()
    dl, douts = tr.step(image.cuda(), text.cuda(), eps=eps, terms=terms, lambdas=lambdas, update=False, outputs=True)
    torch.cuda.synchronize()
    dl = dl.cpu()
    print("== B=%d n=%d %s terms=%s" % (B, n, precision, terms))
    act = [i for i in range(3) if tmask[i]]
    for gi, ti in enumerate(act):
        print("  loss[%d] ref=%.6f got=%.6f (bce %.5f ce %.5f kl %.5f)" % (ti, losses[ti], dl[gi, 0], dl[gi, 1], dl[gi, 2], dl[gi, 3]))
    worst = 0
    sd = m.state_dict()
    for name, p in m.named_parameters():
        g = p.grad
        d, s, l2 = rel(g, grads[name])
        flag = ""
        if name in O.PRE_BN_BIASES:
            flag = "(pre-BN bias: true grad 0)"
        else:
            worst = max(worst, l2)
        print("  grad %-32s maxabs_err=%.3e ref_max=%.3e relL2=%.3e %s" % (name, d, s, l2, flag))
    for k, v in bufs.items():
        d, s, l2 = rel(sd[k].float(), v.float())
        print("  buf  %-40s maxabs_err=%.3e ref_max=%.3e" % (k, d, s))
    ri, rt, mu, lv = douts
    for gi, ti in enumerate(act):
        o = outs[ti]
        print("  out[%d] recon_image relL2=%.2e recon_text relL2=%.2e mu relL2=%.2e logvar relL2=%.2e" % (
            ti, rel(ri[gi * B:(gi + 1) * B].float(), o[0])[2], rel(rt[gi * B:(gi + 1) * B], o[1])[2], rel(mu[gi], o[2])[2], rel(lv[gi], o[3])[2]))
    print("  WORST grad relL2 = %.3e" % worst)


if __name__ == "__main__":
    torch.manual_seed(0)
    run(24, 8, "tf32", seed=3)
    run(100, 64, "tf32", seed=0)
    run(32, 20, "tf32", terms=("joint", "text"), lambdas=((1., 1.), (0., 1.)), seed=9)
    run(100, 64, "bf16", seed=0)
    run(512, 64, "tf32", seed=1)
    run(512, 64, "bf16", seed=1)
    # timing of the full step at the benchmark size
    for prec in ("tf32", "bf16"):
        B, n = 4096, 64
        m = MVAE(n, precision=prec); tr = MVAETrainer(m)
        image, text, _ = O.synthetic_batch(B, n, 0)
        x = m.to_act(image.cuda()); y = text.cuda()
        for _ in range(5):
            tr.step(x, y)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10):
                l, _ = tr.step(x, y)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("STEP %s B=%d: %.1f us/step  %.2f M samples/s  loss=%s" % (prec, B, ms * 1e3, B / ms / 1e3, l[:, 0].tolist()))
