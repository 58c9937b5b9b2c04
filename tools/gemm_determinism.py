"""Is a store-epilogue GEMM bitwise reproducible call to call?  (small M, K < one k-block, tf32 + bf16)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mvae_b200  # noqa
from mvae_b200 import _ops as ops
g = torch.Generator().manual_seed(0)
for dt in (torch.float32, torch.bfloat16):
    for (M, N, K, bmaj) in ((24, 64, 16, 0), (24, 6400, 16, 0), (600, 2048, 256, 1), (24, 64, 100, 0), (4096, 512, 784, 0)):
        A = torch.randn(M, K, generator=g).to(dt).cuda()
        B = (torch.randn(K, N, generator=g) if bmaj else torch.randn(N, K, generator=g)).to(dt).cuda()
        bias = torch.randn(N, generator=g).cuda()
        outs = []
        for it in range(4):
            C = torch.empty(M, N, device="cuda", dtype=torch.float32)
            ops.gemm(A, B, C, M, N, K, K, N if bmaj else K, N, b_major=bmaj, bias=bias)
            torch.cuda.synchronize()
            outs.append(C.clone())
            junk = torch.randn(1 << 20, device="cuda")  # perturb allocator / caches between calls
        ref = (A.float() @ (B.float() if bmaj else B.float().t())) + bias
        same = [bool(torch.equal(outs[0], o)) for o in outs[1:]]
        err = float((outs[0] - ref).norm() / ref.norm())
        d = max(float((outs[0] - o).abs().max()) for o in outs[1:])
        print("%s M=%d N=%d K=%d bmaj=%d: identical %s  max|diff| %.3g  rel err vs fp32 %.2e" % (str(dt)[6:], M, N, K, bmaj, same, d, err))
