"""Sweep stages / block_n for a few shapes: kernel span and mainloop time from %globaltimer stamps."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mvae_b200
from mvae_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0")
def run(dt, M, N, K, am, bm, acc, bn, stages, c_bf16=0):
    tdt = torch.float32 if dt == 0 else torch.bfloat16
    A = torch.randn(K, M, device=dev, dtype=tdt) if am else torch.randn(M, K, device=dev, dtype=tdt)
    B = torch.randn(K, N, device=dev, dtype=tdt) if bm else torch.randn(N, K, device=dev, dtype=tdt)
    Cc = torch.zeros(M, N, device=dev, dtype=torch.bfloat16 if c_bf16 else torch.float32)
    dbg = torch.zeros(8192, 8, device=dev, dtype=torch.int64)
    a = _lib.GemmArgs(dt, M, N, K, A.data_ptr(), A.stride(0), am, B.data_ptr(), B.stride(0), bm, Cc.data_ptr(),
                      N, c_bf16, None, acc, None, None, 0, bn, 0, stages, dbg.data_ptr())
    for it in range(3):
        dbg.zero_()
        _lib.check(lib.mvae_gemm(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
    d = dbg[dbg[:, 0] > 0].cpu()
    t0 = d[:, 0].min()
    span = float((d[:, 7].max() - t0)) / 1e3
    ml = float((d[:, 3] - d[:, 2]).float().median()) / 1e3
    ep = float((d[:, 6] - d[:, 4]).float().median()) / 1e3
    print("dt=%d %5dx%4dx%5d am=%d bm=%d bn=%3d st=%d ctas=%4d span=%6.1f us mainloop=%5.2f epi=%5.2f  %.0f TF/s" % (
        dt, M, N, K, am, bm, bn, stages, d.shape[0], span, ml, ep, 2.0 * M * N * K / span / 1e6), flush=True)
for (M, N, K, am, bm, acc) in [(4096, 400, 784, 0, 0, 0), (12288, 784, 400, 0, 0, 0), (12288, 400, 784, 0, 1, 0), (784, 400, 12288, 1, 1, 1)]:
    for bn in ([64, 80, 112, 144, 208] if N == 400 else [64, 112, 160, 208, 256]):
        for stages in (2, 3, 4, 6):
            run(1, M, N, K, am, bm, acc, bn, stages, c_bf16=0 if acc else 1)
