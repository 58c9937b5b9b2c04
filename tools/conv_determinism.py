"""Run the same conv-MVAE step twice on one GPU and report per-tensor gradient differences (atomics / races)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import mvae_b200  # noqa
import celeba_oracle as CO, multimnist_oracle as MO
from mvae_b200.celeba import MultimodalVAE as CV, CelebATrainer
from mvae_b200.multimnist import MultimodalVAE as MV, MultiMNISTTrainer

def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))

for name, Model, Trainer, fn in (("celeba", CV, CelebATrainer, CO.synthetic_batch), ("multimnist", MV, MultiMNISTTrainer, MO.synthetic_batch)):
    for prec in ("tf32", "bf16"):
        B, n = 8, 16
        img, oth, noises = fn(B, n, 10)
        m = Model(n_latents=n, precision=prec, dropout_p=0.0)
        tr = Trainer(m)
        gs = []
        for it in range(3):
            m.flat_grads.zero_()
            tr.step(img.cuda(), oth.cuda(), eps=torch.stack(noises).cuda(), adam=False)
            torch.cuda.synchronize()
            gs.append({k: v.clone() for k, v in m.grads_reference().items()})
        tot = rel(torch.cat([v.reshape(-1) for v in gs[1].values()]), torch.cat([v.reshape(-1) for v in gs[0].values()]))
        print("%s %s: whole-gradient run-to-run rel diff %.2e" % (name, prec, tot))
        worst = sorted(((rel(gs[1][k], gs[0][k]), k) for k in gs[0] if float(gs[0][k].abs().max()) > 0), reverse=True)[:6]
        for e, k in worst:
            print("    %-48s %.2e  (2 vs 0: %.2e)" % (k, e, rel(gs[2][k], gs[0][k])))
