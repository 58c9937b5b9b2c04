import os, sys, faulthandler
faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import mnist_oracle as O
import mvae_b200
from mvae_b200 import MVAE, MVAETrainer
B, n = int(sys.argv[1]), int(sys.argv[2]); prec = sys.argv[3]
print("start", flush=True)
m = MVAE(n, precision=prec); tr = MVAETrainer(m)
image, text, noises = O.synthetic_batch(B, n, 0)
eps = torch.stack(noises).cuda()
print("calling step", flush=True)
l, _ = tr.step(image.cuda(), text.cuda(), eps=eps, update=False)
torch.cuda.synchronize()
print("losses", l.cpu().tolist(), flush=True)
