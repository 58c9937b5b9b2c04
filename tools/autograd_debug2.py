import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import mvae_b200  # noqa
import celeba_oracle as O
from mvae_b200.celeba import MultimodalVAE, CelebATrainer, loss_function
def rel(a, b):
    a, b = a.double().cpu().reshape(-1), b.double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))
n, B, seed = 16, 8, 2
state = O.init_state(n, seed=1234 + seed)
image, attrs, noises = O.synthetic_batch(B, n, seed)
img, att = image.cuda(), attrs.cuda()
keys = ["image_encoder.classifier.3.bias", "attrs_encoder.net.3.bias", "image_decoder.upsample.0.weight"]
def oracle(kl):
    work = {k: (v.clone() if O.is_buffer(k) else v.clone().requires_grad_(True)) for k, v in state.items()}
    o = O.forward(work, image, attrs, noises[0], work, True)
    O.loss_function(o[2], o[3], o[0], image, o[1], attrs, kl_lambda=kl).backward()
    return work
for kl in (1e-3, 0.0, 1.0):
    w = oracle(kl)
    # trainer path, joint term only
    m = MultimodalVAE(n_latents=n, precision="tf32", dropout_p=0.0); m.load_state_dict(state)
    tr = CelebATrainer(m, kl_lambda=kl)
    tr.step(img, att, terms=("joint",), lambdas=((1.0, 1.0),), eps=noises[0][None].cuda(), adam=False)
    dg = m.grads_reference()
    print("kl", kl, "trainer:", {k: round(rel(dg[k], w[k].grad), 5) for k in keys})
    # module path
    v = MultimodalVAE(n_latents=n, precision="tf32", dropout_p=0.0); v.load_state_dict(state); v.train(); v.zero_grad()
    r = v(image=img, attrs=att, eps=noises[0])
    loss_function(r[2], r[3], recon_x=r[0], x=img, recon_y=r[1], y=att, kl_lambda=kl).backward()
    dg = v.grads_reference()
    print("kl", kl, "module :", {k: round(rel(dg[k], w[k].grad), 5) for k in keys})
