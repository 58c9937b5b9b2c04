"""%globaltimer phase timeline of one GEMM kind inside the real step (bring-up)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mvae_b200
from mvae_b200 import MVAE, MVAETrainer, _lib
kind = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = 4096
m = MVAE(64, precision="bf16"); tr = MVAETrainer(m)
g = torch.Generator().manual_seed(0)
x = m.to_act(torch.rand(B, 784, generator=g).cuda()); y = torch.randint(0, 10, (B,), generator=g).cuda()
lib = _lib.load()
dbg = torch.zeros(8192, 8, device="cuda", dtype=torch.int64)
os.environ["MVAE_SIDE_STREAM"] = "0"
for i in range(3):
    tr.step(x, y)
torch.cuda.synchronize()
lib.mvae_debug_gemm_times(C.c_void_p(dbg.data_ptr()), kind)
dbg.zero_(); tr.step(x, y); torch.cuda.synchronize()
lib.mvae_debug_gemm_times(None, -1)
d = dbg[dbg[:, 0] > 0].cpu()
names = ["start", "setup", "first_full", "mma_issued", "accum_ready", "tmem2smem", "rowpass", "end"]
t0 = d[:, 0].min()
print("kind %d: ctas(last launch of this kind)=%d span=%.1f us" % (kind, d.shape[0], float(d[:, 7].max() - t0) / 1e3))
dur = (d[:, 1:] - d[:, :-1]).float() / 1e3
print("  CTA start: median %.2f max %.2f us" % (float((d[:, 0] - t0).float().median()) / 1e3, float((d[:, 0] - t0).max()) / 1e3))
for i in range(7):
    print("  %-12s -> %-12s median %.2f max %.2f us" % (names[i], names[i + 1], float(dur[:, i].median()), float(dur[:, i].max())))
print("  CTA lifetime median %.2f us" % float(((d[:, 7] - d[:, 0]).float() / 1e3).median()))
