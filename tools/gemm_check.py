"""GPU bring-up check of the tcgen05 GEMM through the C ABI (run on the B200 box).

Compares mvae_gemm against a float64 torch matmul for every operand-major combination, both
storage dtypes, ragged shapes, split-K accumulation and the column statistics.
"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mvae_b200
from mvae_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda:0")


def run(dt, M, N, K, a_mn, b_mn, acc=False, stats=False, bias=False, block_n=0, split_k=0, rpg=0, stages=0):
    tdt = torch.float32 if dt == 0 else torch.bfloat16
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, generator=g).to(dev, tdt)
    B = torch.randn(N, K, generator=g).to(dev, tdt)
    A_st = A.t().contiguous() if a_mn else A.contiguous()
    B_st = B.t().contiguous() if b_mn else B.contiguous()
    Cout = torch.zeros(M, N, device=dev, dtype=torch.float32)
    if acc:
        Cout.fill_(1.0)
    bias_t = torch.randn(N, generator=g).to(dev) if bias else None
    groups = 1 if rpg <= 0 else (M + rpg - 1) // rpg
    s0 = torch.zeros(groups, N, device=dev) if stats else None
    s1 = torch.zeros(groups, N, device=dev) if stats else None
    a = _lib.GemmArgs(dt, M, N, K, A_st.data_ptr(), A_st.stride(0), int(a_mn), B_st.data_ptr(), B_st.stride(0),
                      int(b_mn), Cout.data_ptr(), Cout.stride(0), 0, _lib.ptr(bias_t), int(acc), _lib.ptr(s0),
                      _lib.ptr(s1), rpg, block_n, split_k, stages)
    rc = lib.mvae_gemm(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "mvae_gemm")
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    if dt == 0:
        def rt(x):
            i = x.contiguous().view(torch.int32)
            return ((i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF).view(torch.float32)
        ref_e = rt(A).double() @ rt(B).double().t()
        if bias:
            ref_e = ref_e + bias_t.double()
        if acc:
            ref_e = ref_e + 1.0
        emu_err = ((Cout.double() - ref_e).abs().max() / ref_e.abs().max()).item()
    else:
        emu_err = float("nan")
    if bias:
        ref = ref + bias_t.double()
    if acc:
        ref = ref + 1.0
    err = (Cout.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    msg = "dt=%d M=%5d N=%4d K=%5d a_mn=%d b_mn=%d acc=%d bn=%3d sk=%2d  max_abs_err=%.3e (ref max %.2e) rel=%.2e vs_tf32_emulation=%.2e" % (
        dt, M, N, K, a_mn, b_mn, acc, block_n, split_k, err, scale, err / scale, emu_err)
    ok = err / scale < (2e-3 if dt == 0 else 1e-2)
    if stats:
        y = Cout.double()
        pad = groups * max(rpg, 1) - M if rpg > 0 else 0
        if rpg > 0:
            yp = torch.cat([y, torch.zeros(pad, N, device=dev, dtype=torch.float64)]) if pad else y
            r0 = yp.view(groups, rpg, N).sum(1)
            r1 = (yp * yp).view(groups, rpg, N).sum(1)
        else:
            r0 = y.sum(0, keepdim=True)
            r1 = (y * y).sum(0, keepdim=True)
        e0 = ((s0.double() - r0).abs().max() / r0.abs().max()).item()
        e1 = ((s1.double() - r1).abs().max() / r1.abs().max()).item()
        msg += " stat_rel=(%.1e,%.1e)" % (e0, e1)
        ok = ok and e0 < 1e-4 and e1 < 1e-4
    print(("PASS " if ok else "FAIL ") + msg, flush=True)
    return ok


def main():
    print(torch.cuda.get_device_name(0), "TMA tf32 map:", os.environ.get("MVAE_TMA_TF32", "0"))
    _lib.check(lib.mvae_device_check(0), "device_check")
    allok = True
    for dt in (0, 1):
        # forward-style (K-major both), ragged N and K
        allok &= run(dt, 256, 64, 64, 0, 0, block_n=64)
        allok &= run(dt, 256, 64, 256, 0, 0, block_n=64)
        allok &= run(dt, 4096, 400, 784, 0, 0, bias=True, stats=True)
        allok &= run(dt, 4096, 400, 784, 0, 0, bias=True, stats=True, block_n=208)
        allok &= run(dt, 4096, 200, 400, 0, 0, stats=True, rpg=1024)
        allok &= run(dt, 300, 200, 64, 0, 0, bias=True, stats=True, rpg=100)
        allok &= run(dt, 100, 128, 200, 0, 0)
        allok &= run(dt, 12288, 784, 400, 0, 0, block_n=256)
        # dgrad-style: A K-major, B MN-major
        allok &= run(dt, 256, 64, 64, 0, 1, block_n=64)
        allok &= run(dt, 4096, 400, 784, 0, 1)
        allok &= run(dt, 300, 200, 128, 0, 1)
        allok &= run(dt, 12288, 64, 200, 0, 1)
        # wgrad-style: both MN-major, split-K accumulate
        allok &= run(dt, 128, 64, 64, 1, 1, block_n=64)
        allok &= run(dt, 128, 64, 512, 1, 1, block_n=64)
        allok &= run(dt, 784, 400, 12288, 1, 1, acc=True)
        allok &= run(dt, 400, 784, 4096, 1, 1, acc=True)
        allok &= run(dt, 200, 64, 300, 1, 1, acc=True)
        allok &= run(dt, 128, 200, 4096, 1, 1, acc=True, split_k=7)
        allok &= run(dt, 400, 784, 4096, 1, 0)
    # timing of a few shapes (CUDA events, 20 iterations after warm-up)
    for dt in (0, 1):
        for (M, N, K, am, bm, acc) in [(4096, 400, 784, 0, 0, 0), (12288, 784, 400, 0, 0, 0), (12288, 400, 784, 0, 1, 0),
                                      (784, 400, 12288, 1, 1, 1), (12288, 400, 200, 0, 0, 0)]:
            tdt = torch.float32 if dt == 0 else torch.bfloat16
            A = torch.randn(K, M, device=dev, dtype=tdt) if am else torch.randn(M, K, device=dev, dtype=tdt)
            B = torch.randn(K, N, device=dev, dtype=tdt) if bm else torch.randn(N, K, device=dev, dtype=tdt)
            Cc = torch.zeros(M, N, device=dev)
            a = _lib.GemmArgs(dt, M, N, K, A.data_ptr(), A.stride(0), am, B.data_ptr(), B.stride(0), bm, Cc.data_ptr(),
                              N, 0, None, acc, None, None, 0, 0, 0, 0)
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            for _ in range(3):
                lib.mvae_gemm(C.byref(a), st)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                stg = C.c_void_p(torch.cuda.current_stream().cuda_stream)
                for _ in range(20):
                    lib.mvae_gemm(C.byref(a), stg)
            graph.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 20 * 1e3
            print("time(graph) dt=%d %5dx%4dx%5d am=%d bm=%d: %.1f us  %.1f TFLOP/s" % (dt, M, N, K, am, bm, us,
                                                                               2.0 * M * N * K / us / 1e6), flush=True)
    print("ALL PASS" if allok else "SOME FAILED")
    return 0 if allok else 1


if __name__ == "__main__":
    sys.exit(main())
