"""Which workspace buffer first differs between two identical CelebA steps (tf32)?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import mvae_b200  # noqa
import celeba_oracle as CO
from mvae_b200.celeba import MultimodalVAE as CV, CelebATrainer
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
B, n = 8, 16
img, oth, noises = CO.synthetic_batch(B, n, 10)
m = CV(n_latents=n, precision=prec, dropout_p=0.0)
tr = CelebATrainer(m)
ws = m.workspace(B, 3)
names = ["enc_col", "enc_pre", "enc_sum", "enc_sumsq", "enc_mean", "enc_rstd", "enc_act", "f1pre", "f1", "encA", "t1pre", "t1", "encB", "z", "mu", "u1pre", "u1",
         "dec_pre", "dec_sum", "dec_sumsq", "dec_mean", "dec_rstd", "dec_act", "logits", "s1pre", "s1", "alogits", "dalog", "ds1", "ds1pre", "dz", "dec_s0", "dec_s1", "dec_dpre", "dec_dact", "du1", "du1pre",
         "dencA", "dencB", "dt1", "dt1pre", "df1", "df1pre", "enc_dact", "enc_s0", "enc_s1", "enc_dpre"]
def snap():
    out = {}
    for k in names:
        v = getattr(ws, k)
        if isinstance(v, list):
            for i, t in enumerate(v):
                out["%s[%d]" % (k, i)] = t.clone()
        else:
            out[k] = v.clone()
    return out
snaps = []
for it in range(3):
    m.flat_grads.zero_()
    tr.step(img.cuda(), oth.cuda(), eps=torch.stack(noises).cuda(), adam=False)
    torch.cuda.synchronize()
    snaps.append(snap())
    from mvae_b200 import _ops
    tmp = torch.empty_like(ws.s1pre)
    wad, ldad = m._operand_cached("attrs_decoder.net.0.weight", n)
    _ops.gemm(ws.z, wad, tmp, 3 * B, 64, n, ws.ld_z, ldad, 64, bias=m.P("attrs_decoder.net.0.bias"))
    torch.cuda.synchronize()
    print("run %d: recomputed s1pre == stored: %s   (max diff %.3g)  z checksum %.10f  W checksum %.10f" % (
        it, bool(torch.equal(tmp, ws.s1pre)), float((tmp - ws.s1pre).abs().max()), float(ws.z.double().sum()),
        float(m.P("attrs_decoder.net.0.weight").double().sum())))
def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))
for k in snaps[0]:
    e1, e2 = rel(snaps[1][k], snaps[0][k]), rel(snaps[2][k], snaps[0][k])
    if e1 > 0 or e2 > 0:
        print("%-16s run1-vs-0 %.2e   run2-vs-0 %.2e   run2-vs-1 %.2e   max|x| %.3g" % (k, e1, e2, rel(snaps[2][k], snaps[1][k]), float(snaps[0][k].abs().max())))
