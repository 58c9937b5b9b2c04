"""Bring-up: phase timeline of the slab-persistent chain kernels (csrc/chain.cu) from in-kernel %globaltimer stamps.
    python tools/chain_timeline.py [B]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
import mnist_oracle as O  # noqa: E402
import mvae_b200  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = 64
state = O.perturbed_state(n, 1)
image, text, noises = O.synthetic_batch(B, n, 1)
m = mvae_b200.MVAE(n, precision="bf16")
m.load_state_dict(state)
tr = mvae_b200.MVAETrainer(m)
x, y = m.to_act(image.cuda()), text.cuda()
lib = mvae_b200._lib.load()
buf = torch.zeros(4, 148, 32, dtype=torch.int64, device="cuda")
for _ in range(3):
    tr.step(x, y)
torch.cuda.synchronize()
lib.mvae_debug_chain_times(C.c_void_p(buf.data_ptr()))
tr.step(x, y)
torch.cuda.synchronize()
lib.mvae_debug_chain_times(C.c_void_p(None))
t = buf.cpu()
names = ["enc_fwd", "dec_fwd", "dec_bwd", "enc_bwd"]
slots = {0: "start"}
for i in range(8):
    slots[1 + i] = "producer pass %d issued" % i
    slots[9 + i] = "mma pass %d issued" % i
for il in range(3):
    slots[17 + il * 4] = "epi L%d begin" % il
    slots[18 + il * 4] = "epi L%d pass1 done" % il
    slots[19 + il * 4] = "epi L%d barrier done" % il
    slots[20 + il * 4] = "epi L%d done" % il
slots[31] = "end"
for k, name in enumerate(names):
    tk = t[k]
    used = tk[:, 0] > 0
    if not bool(used.any()):
        continue
    tk = tk[used]
    t0 = int(tk[:, 0].min())
    print("%s: %d CTAs, first start -> last end %.2f us; start skew %.2f us" % (
        name, tk.shape[0], (int(tk[:, 31].max()) - t0) / 1e3, (int(tk[:, 0].max()) - t0) / 1e3))
    for s in sorted(slots):
        col = tk[:, s]
        ok = col > 0
        if not bool(ok.any()):
            continue
        rel = (col[ok] - t0).double() / 1e3
        print("   %-28s cta0 %8.2f   min %8.2f  median %8.2f  max %8.2f us" % (
            slots[s], (int(tk[0, s]) - t0) / 1e3 if int(tk[0, s]) > 0 else float("nan"), float(rel.min()), float(rel.median()), float(rel.max())))
