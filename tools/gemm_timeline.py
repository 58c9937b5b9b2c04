"""Per-CTA phase timeline of the GEMM kernel via %globaltimer stamps (bring-up tool, B200 box)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mvae_b200
from mvae_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0")
conv = [(1, 196608, 512, 64, 0, 1, 0, 0), (1, 786432, 48, 32, 0, 1, 0, 0), (1, 65536, 64, 512, 0, 0, 0, 0)]
small = [(1, 12288, 200, 64, 0, 0, 0, 0), (1, 4096, 200, 400, 0, 0, 0, 0), (1, 12288, 400, 200, 0, 0, 0, 0)]
shapes = conv if "conv" in sys.argv else small if "small" in sys.argv else [(1, 4096, 400, 784, 0, 0, 0, 0), (1, 4096, 400, 784, 0, 0, 0, 208), (0, 4096, 400, 784, 0, 0, 0, 0),
          (1, 12288, 400, 200, 0, 0, 0, 0), (1, 12288, 784, 400, 0, 0, 0, 0), (1, 784, 400, 12288, 1, 1, 1, 0)]
names = ["start", "setup", "first_full", "mma_done_issue", "accum_ready", "tmem2smem", "rowpass", "end"]
for (dt, M, N, K, am, bm, acc, bn) in shapes:
    tdt = torch.float32 if dt == 0 else torch.bfloat16
    A = torch.randn(K, M, device=dev, dtype=tdt) if am else torch.randn(M, K, device=dev, dtype=tdt)
    B = torch.randn(K, N, device=dev, dtype=tdt) if bm else torch.randn(N, K, device=dev, dtype=tdt)
    Cc = torch.zeros(M, N, device=dev, dtype=tdt if "conv" in sys.argv else torch.float32)
    dbg = torch.zeros(16384, 8, device=dev, dtype=torch.int64)
    st0 = torch.zeros(3, N, device=dev); st1 = torch.zeros(3, N, device=dev)
    stats = "small" in sys.argv
    if stats:
        Cc = torch.zeros(M, N, device=dev, dtype=tdt)
    a = _lib.GemmArgs(dt, M, N, K, A.data_ptr(), A.stride(0), am, B.data_ptr(), B.stride(0), bm, Cc.data_ptr(),
                      N, (dt if ("conv" in sys.argv or stats) else 0), None, acc, st0.data_ptr() if stats else None,
                      st1.data_ptr() if stats else None, 4096 if stats else 0, bn, 0, 0, dbg.data_ptr())
    for it in range(3):
        dbg.zero_()
        _lib.check(lib.mvae_gemm(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
    d = dbg[dbg[:, 0] > 0].cpu()
    t0 = d[:, 0].min()
    rel = (d - t0).float()
    print("shape dt=%d %dx%dx%d am=%d bm=%d bn=%d: ctas=%d kernel span=%.1f us" % (dt, M, N, K, am, bm, bn, d.shape[0], float(rel[:, 7].max()) / 1e3))
    dur = (d[:, 1:] - d[:, :-1]).float() / 1e3
    print("   CTA start spread: median %.2f max %.2f us" % (float(rel[:, 0].median()) / 1e3, float(rel[:, 0].max()) / 1e3))
    for i in range(7):
        print("   %-16s -> %-16s median %.2f us  max %.2f us" % (names[i], names[i + 1], float(dur[:, i].median()), float(dur[:, i].max())))
