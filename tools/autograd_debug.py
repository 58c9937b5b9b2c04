import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import mvae_b200  # noqa
import celeba_oracle as O
from mvae_b200.celeba import MultimodalVAE, loss_function
def rel(a, b):
    a, b = a.double().cpu().reshape(-1), b.double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))
n, B, seed = 16, 8, 2
state = O.init_state(n, seed=1234 + seed)
image, attrs, noises = O.synthetic_batch(B, n, seed)
img, att = image.cuda(), attrs.cuda()
calls = (dict(image=img, attrs=att), dict(image=img), dict(attrs=att))
ocalls = ((image, attrs), (image, None), (None, attrs))
for which in ([0], [1], [2], [0, 1, 2]):
    vae = MultimodalVAE(n_latents=n, precision="tf32", dropout_p=0.0)
    vae.load_state_dict(state); vae.train(); vae.zero_grad()
    tot = 0
    for k in which:
        r = vae(eps=noises[k], **calls[k])
        tot = tot + loss_function(r[2], r[3], recon_x=r[0], x=img, recon_y=r[1], y=att)
    tot.backward()
    work = {k: (v.clone() if O.is_buffer(k) else v.clone().requires_grad_(True)) for k, v in state.items()}
    rt = 0
    for k in which:
        o = O.forward(work, ocalls[k][0], ocalls[k][1], noises[k], work, True)
        rt = rt + O.loss_function(o[2], o[3], o[0], image, o[1], attrs)
    rt.backward()
    dg = vae.grads_reference()
    bad = {k: round(rel(dg[k], work[k].grad), 4) for k in dg if work[k].grad is not None and float(work[k].grad.abs().max()) > 1e-7 and rel(dg[k], work[k].grad) > 4e-3}
    print("terms", which, "loss dev %.6f ref %.6f" % (float(tot), float(rt)), "bad:", bad)
