"""Small fixed GEMM workload for ncu (run on the B200 box)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mvae_b200
from mvae_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0")
shapes = [(0, 4096, 400, 784, 0, 0, 0), (1, 4096, 400, 784, 0, 0, 0), (1, 12288, 400, 200, 0, 0, 0),
          (1, 12288, 784, 400, 0, 0, 0), (1, 12288, 400, 784, 0, 1, 0), (1, 784, 400, 12288, 1, 1, 1)]
for it in range(3):
    for (dt, M, N, K, am, bm, acc) in shapes:
        tdt = torch.float32 if dt == 0 else torch.bfloat16
        A = torch.randn(K, M, device=dev, dtype=tdt) if am else torch.randn(M, K, device=dev, dtype=tdt)
        B = torch.randn(K, N, device=dev, dtype=tdt) if bm else torch.randn(N, K, device=dev, dtype=tdt)
        Cc = torch.zeros(M, N, device=dev)
        a = _lib.GemmArgs(dt, M, N, K, A.data_ptr(), A.stride(0), am, B.data_ptr(), B.stride(0), bm, Cc.data_ptr(),
                          N, 0, None, acc, None, None, 0, int(os.environ.get("BN", "0")), 0, 0)
        _lib.check(lib.mvae_gemm(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
print("ok")
