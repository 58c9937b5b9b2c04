"""Compare device intermediates of the decoder backward against an emulated-oracle recomputation (bring-up)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, warnings
warnings.filterwarnings("ignore")
import mnist_oracle as O
import mvae_b200
from mvae_b200 import MVAE, MVAETrainer
def rel(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
B, n, seed = int(sys.argv[1]), 64, 0
state = O.perturbed_state(n, seed); image, text, noises = O.synthetic_batch(B, n, seed)
m = MVAE(n, precision="tf32"); m.load_state_dict(state); tr = MVAETrainer(m)
os.environ["MVAE_SIDE_STREAM"] = os.environ.get("MVAE_SIDE_STREAM", "1")
dl, douts = tr.step(image.cuda(), text.cuda(), eps=torch.stack(noises).cuda(), update=False, outputs=True)
torch.cuda.synchronize()
R = 3 * B
buf = lambda name, shape, dt=torch.float32: m.debug_buffer(name, B, shape, dt).clone().cpu()
z, g1pre, g1, g2pre, g2 = buf("z", (R, n)), buf("g1pre", (R, 200)), buf("g1", (R, 200)), buf("g2pre", (R, 400)), buf("g2", (R, 400))
dlog, dy2, dy1, dz = buf("dlog", (R, 784)), buf("dy2", (R, 400)), buf("dy1", (R, 200)), buf("dz", (R, n))
sv_d2 = buf("sv_d2", (2, 3, 400)); sb_d2 = buf("sb_d2", (2, 3, 400))
# recompute the decoder backward on the CPU from the DEVICE's own forward tensors, emulating tf32 operand rounding
rt = O.round_tf32
W3 = state["image_decoder.net.6.weight"]; W2 = state["image_decoder.net.3.weight"]
ga2 = state["image_decoder.net.4.weight"]; be2 = state["image_decoder.net.4.bias"]
dh2 = rt(dlog).double() @ rt(W3).double()                       # [R,400]
out = []
for g in range(3):
    sl = slice(g * B, (g + 1) * B)
    x = g2pre[sl].double(); mean = x.mean(0); var = x.var(0, unbiased=False); rstd = 1 / torch.sqrt(var + 1e-5)
    xh = (x - mean) * rstd
    y = ga2.double() * xh + be2.double()
    dyh = dh2[sl] * (y > 0)
    s0 = dyh.sum(0); s1 = (dyh * xh).sum(0)
    dx = ga2.double() * rstd * (dyh - s0 / B - xh * s1 / B)
    out.append((dyh, dx, mean, rstd, s0, s1))
dx_ref = torch.cat([o[1] for o in out]); 
print("B=%d: dy2(after bn_bwd = dG2pre) dev vs recomputed: %.2e" % (B, rel(dy2, dx_ref)))
print("   saved mean %.2e rstd %.2e   s0 %.2e s1 %.2e" % (rel(sv_d2[0], torch.stack([o[2] for o in out])), rel(sv_d2[1], torch.stack([o[3] for o in out])),
      rel(sb_d2[0], torch.stack([o[4] for o in out])), rel(sb_d2[1], torch.stack([o[5] for o in out]))))
for g in range(3):
    sl = slice(g * B, (g + 1) * B)
    print("   group %d dG2pre rel %.2e   |dyh| %.3e |dx| %.3e" % (g, rel(dy2[sl], out[g][1]), float(out[g][0].norm()), float(out[g][1].norm())))
