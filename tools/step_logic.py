import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, warnings
warnings.filterwarnings("ignore")
import mnist_oracle as O
from helpers import *
for prec in ("tf32", "bf16"):
  for (B, n, seed) in [(24, 8, 3), (100, 64, 0), (130, 24, 5), (512, 64, 1)]:
    state = O.perturbed_state(n, seed); image, text, noises = O.synthetic_batch(B, n, seed)
    m, tr, dl, outs = run_device_step(state, image, text, noises, n, prec)
    ov = device_forward_override(m, B, text)
    l, g, _, o = oracle_step(state, image, text, noises)
    lo, go, _, oo = oracle_step(state, image, text, noises, emulate=prec, override=ov)
    print("== %s B=%d n=%d  losses dev %s given-fwd %s exact %s" % (prec, B, n, dl[:, 0].tolist(), lo, l))
    worst = 0
    for name, p in m.named_parameters():
        if name in O.PRE_BN_BIASES: continue
        a, b = rel_l2(p.grad, go[name]), rel_l2(p.grad, g[name])
        worst = max(worst, a)
        print("   %-32s dev-vs-oracle(given device forward) %.2e   dev-vs-exact %.2e" % (name, a, b))
    print("   WORST given-forward %.2e" % worst)
