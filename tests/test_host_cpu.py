"""CPU-side tests (no GPU): the C ABI loads and exports every declared symbol, the parameter layout matches the
reference's state_dict, host-side data-parallel logic works over gloo with world_size 2."""
import ctypes as C
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

import mnist_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import mvae_b200
    from mvae_b200 import _lib
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "mvae_b200.h")).read()
    names = set(re.findall(r"\b(mvae_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 14
    for n in sorted(names):
        assert hasattr(lib, n), n
    m = re.search(r"#define MVAE_ABI_VERSION (\d+)", hdr)
    assert lib.mvae_abi_version() == int(m.group(1))


def test_library_sass_is_tcgen05_tma_native():
    """The built library is sm_100a code on the Blackwell tensor path: its SASS carries tcgen05.mma (UTC*MMA), tcgen05.ld
    (LDTM), TMA loads (UTMALDG), the cta_group::2 forms of the pair kernels and the programmatic-launch instructions -
    and nothing falls back to mma.sync (HMMA) tiles.  (cuobjdump only; no GPU.)"""
    import shutil
    import subprocess
    from mvae_b200 import _lib
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    _lib.load()
    so = os.path.join(ROOT, "multimodal-vae_b200", "libmvae_b200.so")
    elf = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in elf, elf
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    count = lambda pat: len(re.findall(pat, sass))
    assert count(r"\bUTC[A-Z]*MMA") >= 64          # every GEMM / chain instantiation issues tcgen05.mma
    assert count(r"\bLDTM") >= 64 and count(r"\bUTMALDG") >= 64
    assert count(r"\bUTC[A-Z]*MMA\.2CTA|\.2CTA") >= 8   # CTA-pair chain kernels
    assert count(r"\bACQBULK") >= 8 and count(r"\bPREEXIT") >= 8
    assert count(r"\bHMMA\.") == 0


def test_layout_matches_reference_state_dict():
    import mvae_b200
    from mvae_b200 import mnist
    for n in (8, 20, 64):
        table = mnist.tensor_table(n)
        shapes = O.param_shapes(n)
        assert [t[0] for t in table] == list(shapes.keys())   # same keys, same order as the reference
        for name, kind, shape, off in table:
            assert tuple(shape) == tuple(shapes[name]), name
            assert (kind != 0) == O.is_buffer(name)
        si = mnist.sizes(n, 128, 0)
        n_params = sum(int(torch.tensor(s).prod()) if len(s) else 1 for k, s in shapes.items() if not O.is_buffer(k))
        assert si.param_floats >= n_params and si.param_floats % 64 == 0
    assert sum(int(torch.tensor(s).prod()) for k, s in O.param_shapes(64).items() if not O.is_buffer(k)) == 838020


def test_errors_are_reported_not_swallowed():
    import mvae_b200
    from mvae_b200 import _lib
    lib = _lib.load()
    si = _lib.MnistSizeInfo()
    assert lib.mvae_mnist_sizes(7, 128, 0, C.byref(si)) != 0
    assert b"n_latents" in lib.mvae_last_error()
    with pytest.raises(_lib.MvaeError):
        _lib.check(lib.mvae_mnist_sizes(64, 1, 0, C.byref(si)), "sizes")


def test_no_cpu_fallback():
    """The product must fail loudly without a CUDA device instead of computing on the CPU."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import mvae_b200
    with pytest.raises(Exception):
        mvae_b200.MVAE(8, device=torch.device("cpu"))
    with pytest.raises(Exception):
        mvae_b200.MVAE(8)


def test_bucket_planner():
    from mvae_b200.parallel import plan_buckets
    assert plan_buckets([10, 10, 10], 25) == [(0, 20), (20, 30)]
    assert plan_buckets([100], 10) == [(0, 100)]
    assert plan_buckets([], 10) == []
    b = plan_buckets([3, 5, 7, 2, 9], 10)
    assert b[0][0] == 0 and b[-1][1] == 26 and all(x[1] == y[0] for x, y in zip(b, b[1:]))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import mvae_b200
    from mvae_b200.parallel import allreduce_flat_, plan_buckets
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    # gradient averaging over equal shards == gradient of the global mean (the identity the DP step relies on)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(8, 5, generator=g)
    w = torch.randn(5, 3, generator=g, requires_grad=True)
    shard = x[rank * 4:(rank + 1) * 4]
    (shard @ w).pow(2).mean().backward()
    flat = w.grad.reshape(-1).clone()
    allreduce_flat_(flat, buckets=plan_buckets([7, 8], 8))
    flat /= world
    w2 = w.detach().clone().requires_grad_(True)
    (x @ w2).pow(2).mean().backward()
    ok = torch.allclose(flat, w2.grad.reshape(-1), atol=1e-6)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_dp_gradient_averaging_gloo_world2():
    import torch.multiprocessing as mp
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference runs on CPU only (the oracle port) and prints ONE JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["value"] > 0 and d["vs_baseline"] is None


def test_conv_models_layout_and_no_cpu_fallback():
    """Host logic of the conv models without a GPU: the reference state_dict layout (keys, shapes, order) matches the
    oracles', the internal <-> reference layout conversions are exact inverses, and construction fails loudly on CPU."""
    import pytest
    import torch
    import mvae_b200  # noqa: F401
    from mvae_b200 import celeba, multimnist
    from mvae_b200.convnet import Layout
    import celeba_oracle as CO
    import multimnist_oracle as MO
    for mod, O in ((celeba, CO), (multimnist, MO)):
        cls = mod.MultimodalVAE
        keys = cls.reference_keys(cls.__new__(cls), 100)
        shapes = O.param_shapes(100)
        assert [k for k, _, _ in keys] == list(shapes.keys())
        assert all(tuple(s) == tuple(shapes[k]) for k, s, _ in keys)
        g = torch.Generator().manual_seed(0)
        for k, s, kind in keys:
            if kind in ("rm", "rv", "nbt"):
                continue
            l = Layout(k, s, kind, cls.FLAT_C, cls.FLAT_HW)
            t = torch.randn(s, generator=g)
            assert torch.equal(l.to_reference(l.to_internal(t).reshape(-1)), t), k
        if not torch.cuda.is_available():
            with pytest.raises(Exception):
                cls(n_latents=16)


@pytest.mark.parametrize("k,s,p,hin", [(4, 2, 1, 8), (4, 2, 0, 2), (5, 2, 1, 12), (4, 1, 0, 5), (3, 3, 1, 4), (4, 2, 1, 5)])
def test_transposed_conv_parity_classes_match_torch(k, s, p, hin):
    """The index map of the implicit transposed convolution (mvae_b200._ops.transposed_conv_classes): every output-parity
    class is a stride-1 gather GEMM; assembled, the classes reproduce F.conv_transpose2d exactly (CPU, float64)."""
    import torch
    import torch.nn.functional as F
    from mvae_b200 import _ops
    g = torch.Generator().manual_seed(k * 100 + s * 10 + p)
    N, Ci, Co = 2, 3, 4
    x = torch.randn(N, Ci, hin, hin, generator=g, dtype=torch.float64)
    w = torch.randn(Ci, Co, k, k, generator=g, dtype=torch.float64)      # nn.ConvTranspose2d weight layout
    ref = F.conv_transpose2d(x, w, stride=s, padding=p)
    hout, classes = _ops.transposed_conv_classes(k, s, p, hin)
    assert hout == ref.shape[-1]
    xl = x.permute(0, 2, 3, 1)                                            # NHWC
    out = torch.zeros(N, hout, hout, Co, dtype=torch.float64)
    covered = torch.zeros(hout, hout, dtype=torch.int64)
    for ca in classes:
        for cb in classes:
            Ta, Tb = len(ca["kh"]), len(cb["kh"])
            if ca["count"] == 0 or cb["count"] == 0:
                continue
            # patch matrix A[(n,u,v), (t'h, t'w, ci)] and weights B[(t'h, t'w, ci), co]
            A = torch.zeros(N, ca["count"], cb["count"], Ta, Tb, Ci, dtype=torch.float64)
            for th in range(Ta):
                for tw in range(Tb):
                    for u in range(ca["count"]):
                        ih = u - ca["pad_lo"] + th
                        if not 0 <= ih < hin:
                            continue
                        for v in range(cb["count"]):
                            iw = v - cb["pad_lo"] + tw
                            if 0 <= iw < hin:
                                A[:, u, v, th, tw] = xl[:, ih, iw]
            Bm = torch.stack([torch.stack([w[:, :, ca["kh"][th], cb["kh"][tw]] for tw in range(Tb)]) for th in range(Ta)])
            y = A.reshape(N * ca["count"] * cb["count"], -1) @ Bm.reshape(Ta * Tb * Ci, Co)
            out[:, ca["a"]::s, cb["a"]::s] = y.view(N, ca["count"], cb["count"], Co)
            covered[ca["a"]::s, cb["a"]::s] += 1
    assert int(covered.min()) == 1 and int(covered.max()) == 1
    torch.testing.assert_close(out.permute(0, 3, 1, 2), ref, rtol=1e-12, atol=1e-12)


def test_gather_multiply_shift_division_is_exact():
    """The implicit-GEMM gather maps pixel index -> (n, ho, wo) with q = (m * magic) >> 40, magic = 2^40 // d + 1
    (csrc/gemm.cu, launch_gemm): exact for m < 2^24 and d < 2^16, the limits the launcher enforces."""
    rng = np.random.default_rng(0)
    for d in [1, 2, 3, 5, 25, 64, 255, 256, 1023, 1024, 4096, 65535] + [int(x) for x in rng.integers(1, 65536, size=40)]:
        magic = np.uint64((1 << 40) // d + 1)
        mult = np.arange(0, (1 << 24) // d + 1, dtype=np.uint64) * np.uint64(d)
        m = np.concatenate([rng.integers(0, 1 << 24, size=20000, dtype=np.uint64), mult[mult < (1 << 24)],
                            mult[(mult > 0) & (mult <= (1 << 24))] - np.uint64(1), np.array([(1 << 24) - 1], dtype=np.uint64)])
        assert ((m * magic) >> np.uint64(40) == m // np.uint64(d)).all(), d


def test_c_parity_class_tables_equal_the_python_index_map():
    """mvae_convt_axis_classes (the tables mvae_convt_gemm puts into the kernel parameters) == _ops.transposed_conv_classes,
    which test_transposed_conv_parity_classes_match_torch pins to F.conv_transpose2d."""
    import mvae_b200  # noqa: F401
    from mvae_b200 import _lib, _ops
    lib = _lib.load()
    I4, I32 = C.c_int * 4, C.c_int * 32
    for k in range(1, 9):
        for s in range(1, 5):
            for p in range(0, 4):
                for n in (1, 2, 5, 8, 13):
                    if (n - 1) * s - 2 * p + k <= 0:
                        continue
                    cnt, taps, pad_lo, kh = I4(), I4(), I4(), I32()
                    out = lib.mvae_convt_axis_classes(k, s, p, n, cnt, taps, pad_lo, kh)
                    size_out, classes = _ops.transposed_conv_classes(k, s, p, n)
                    if max(len(c["kh"]) for c in classes) > 8:
                        assert out == -1
                        continue
                    assert out == size_out, (k, s, p, n)
                    for c in classes:
                        a = c["a"]
                        assert (cnt[a], taps[a], pad_lo[a]) == (c["count"], len(c["kh"]), c["pad_lo"]), (k, s, p, n, a)
                        assert [kh[a * 8 + t] for t in range(taps[a])] == c["kh"], (k, s, p, n, a)
    assert lib.mvae_convt_axis_classes(4, 5, 1, 8, I4(), I4(), I4(), I32()) == -1   # stride > 4
