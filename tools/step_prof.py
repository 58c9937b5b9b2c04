"""A few fused MNIST steps at the benchmark size, for ncu (run on the B200 box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import mvae_b200
from mvae_b200 import MVAE, MVAETrainer
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
m = MVAE(64, precision=prec); tr = MVAETrainer(m)
g = torch.Generator().manual_seed(0)
x = m.to_act(torch.rand(B, 784, generator=g).cuda()); y = torch.randint(0, 10, (B,), generator=g).cuda()
for _ in range(steps):
    l, _ = tr.step(x, y)
torch.cuda.synchronize()
print("ok", l[:, 0].tolist())
