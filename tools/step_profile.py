"""Per-launch CUDA-event durations of one fused step (serialised on one stream), B200 box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mvae_b200
from mvae_b200 import MVAE, MVAETrainer
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
m = MVAE(64, precision=prec); tr = MVAETrainer(m)
g = torch.Generator().manual_seed(0)
xs = [m.to_act(torch.rand(B, 784, generator=g).cuda()) for _ in range(4)]
ys = [torch.randint(0, 10, (B,), generator=g).cuda() for _ in range(4)]
tt, klw = tr._norm(("joint", "image", "text"), B, 1.0)
runs = []
for i in range(8):
    runs.append(m.profile(xs[i % 4], ys[i % 4], tt, ((1.0, 1.0),) * 3, klw, backward=True, zero_grad=True, adam=tr.adam))
n = len(runs[0]); tot = 0
for j in range(n):
    ms = sorted(r[j][1] for r in runs[2:])
    med = ms[len(ms) // 2]; tot += med
    print("%-58s %7.1f us" % (runs[0][j][0], med * 1e3))
print("TOTAL serialised %.1f us" % (tot * 1e3))
