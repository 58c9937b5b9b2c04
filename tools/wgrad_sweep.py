"""Sweep block_n / split_k / stages for the weight-gradient GEMM shapes of the MNIST step (kernel span from %globaltimer stamps,
and CUDA-event time of 50 back-to-back launches)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mvae_b200
from mvae_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0")
def run(M, N, K, bn, sk, stages):
    A = torch.randn(K, M, device=dev, dtype=torch.bfloat16)
    B = torch.randn(K, N, device=dev, dtype=torch.bfloat16)
    Cc = torch.zeros(M, N, device=dev, dtype=torch.float32)
    dbg = torch.zeros(8192, 8, device=dev, dtype=torch.int64)
    a = _lib.GemmArgs(1, M, N, K, A.data_ptr(), A.stride(0), 1, B.data_ptr(), B.stride(0), 1, Cc.data_ptr(),
                      N, 0, None, 1, None, None, 0, bn, sk, stages, dbg.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for it in range(3):
        dbg.zero_()
        _lib.check(lib.mvae_gemm(C.byref(a), st))
        torch.cuda.synchronize()
    d = dbg[dbg[:, 0] > 0].cpu()
    span = float((d[:, 7].max() - d[:, 0].min())) / 1e3
    a.debug_times = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            _lib.check(lib.mvae_gemm(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print("%4dx%4dx%5d bn=%3d split=%2d st=%d ctas=%4d span=%6.1f us  graph %.1f us/launch" % (
        M, N, K, bn, sk, stages, d.shape[0], span, e0.elapsed_time(e1) * 1e3 / 20), flush=True)
shapes = [(400, 784, 4096), (784, 400, 12288)] if len(sys.argv) < 2 else [tuple(int(x) for x in sys.argv[1:4])]
for (M, N, K) in shapes:
    run(M, N, K, 0, 0, 0)
    for bn in (112, 160, 208, 256):
        for sk in (4, 6, 8, 11, 16):
            run(M, N, K, bn, sk, 0)
