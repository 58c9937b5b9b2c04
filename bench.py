#!/usr/bin/env python
"""Benchmark of the MVAE training step (BASELINE.json metric: "MVAE train samples/sec (fwd+bwd ELBO)").

    python bench.py --gpus 1 --steps K --warmup W                 # this repo's CUDA path
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps K --warmup W         # the reference's CPU path (oracle port)

One "step" = one full three-term ELBO training step (forward, backward, Adam) on one batch of synthetic
MNIST-shaped data (config[1] of BASELINE.json: MNIST MVAE, n_latents=64, batch 4096 per GPU).  Prints ONE
JSON line (rank 0).  Timing rules followed: >= 3 warm-up steps, device-side CUDA events, barrier + synchronize
on both sides, max over ranks, inputs rotate through a pool larger than L2, clocks sampled during the timed
region.  Nothing here reads /root/reference.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_LATENTS = 64
BATCH_PER_GPU = 4096
L2_BYTES = 126 * 1024 * 1024
# algorithmic work per sample per step for the reference MNIST architecture
F_ALG_PER_SAMPLE = 9_254_920          # FLOP on tensor cores (encoders once, decoders x3, fwd + dgrad + wgrad)
Q_TAIL_PER_SAMPLE = 29_880            # bytes of the fused PoE/reparam/KL + BCE/CE tail, fp32 I/O


def log(msg):
    if os.environ.get("MVAE_BENCH_VERBOSE"):
        sys.stderr.write("[bench r%s %.1fs] %s\n" % (os.environ.get("RANK", "0"), time.time() % 1000, msg))
        sys.stderr.flush()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _sample(self):
        nv = self.nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
            else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        for k, bit in names.items():
            if r & bit:
                self.reasons.add(k)

    def run(self):
        if self.nv is None:
            return
        while not self._stop_evt.is_set():
            try:
                self._sample()
            except Exception:
                break
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if self.nv is not None and not self.samples:
            try:
                self._sample()
            except Exception:
                pass
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def synthetic_pool(batch, n_slots, seed, device, torch):
    """Pool of distinct synthetic batches: uint8 MNIST-shaped images and labels (SURVEY 8d shapes)."""
    g = torch.Generator().manual_seed(seed)
    imgs = torch.randint(0, 256, (n_slots, batch, 784), generator=g, dtype=torch.uint8)
    labels = torch.randint(0, 10, (n_slots, batch), generator=g, dtype=torch.int64)
    return imgs, labels


# ------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU implementation of the step, timed on the host cores.  The reference is a set of
    Python scripts over PyTorch ATen (nothing to compile); /root/reference is absent on the GPU box, so the
    arm runs oracle/mnist_oracle.py - the line-by-line restatement pinned to the reference's golden vectors."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mnist_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = BATCH_PER_GPU
    state = O.init_state(N_LATENTS, seed=1234)
    image, text, noises = O.synthetic_batch(B, N_LATENTS, 0)
    mom = {k: torch.zeros_like(v) for k, v in state.items() if not O.is_buffer(k)}
    vel = {k: torch.zeros_like(v) for k, v in state.items() if not O.is_buffer(k)}
    # bounded sample: at ~0.25 s per step the whole run must end within minutes
    steps = min(args.steps, 40)
    warmup = max(3, min(args.warmup, 50))   # --warmup is honoured (capped: a CPU step takes ~0.1 s)
    p = state
    step_no = 0
    noise_pool = [[torch.randn_like(n) for n in noises] for _ in range(8)]   # drawn outside the timed loop

    def one():
        nonlocal p, step_no
        step_no += 1
        _, grads, bufs, _ = O.train_step(p, image, text, noise_pool[step_no % len(noise_pool)])
        p = O.adam_step(p, grads, mom, vel, step_no)
        p.update(bufs)

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    value = B * steps / dt
    line = {
        "impl": "reference", "metric": "MVAE train samples/sec (fwd+bwd ELBO)", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "MNIST MVAE n_latents=64 batch=4096 3-term ELBO step (fwd+bwd+Adam), CPU oracle port",
                   "batch": B, "n_latents": N_LATENTS},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "%d steps of batch %d after %d warm-up" % (steps, B, warmup)},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def cpu_baseline(torch, budget_s=12.0):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mnist_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = BATCH_PER_GPU
    state = O.init_state(N_LATENTS, seed=1234)
    image, text, noises = O.synthetic_batch(B, N_LATENTS, 0)
    mom = {k: torch.zeros_like(v) for k, v in state.items() if not O.is_buffer(k)}
    vel = {k: torch.zeros_like(v) for k, v in state.items() if not O.is_buffer(k)}
    p, n, t_used = state, 0, 0.0
    for i in range(2):
        _, grads, bufs, _ = O.train_step(p, image, text, noises)
    while t_used < budget_s and n < 40:
        t0 = time.perf_counter()
        _, grads, bufs, _ = O.train_step(p, image, text, noises)
        p = O.adam_step(p, grads, mom, vel, n + 1)
        p.update(bufs)
        t_used += time.perf_counter() - t0
        n += 1
    return {"value": B * n / t_used, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d steps of batch %d (oracle/mnist_oracle.py, fp32, %d threads)" % (n, B, torch.get_num_threads())}


def torch_gpu_baseline(torch, dev, budget_steps=30):
    """BASELINE.md section 2 comparator: the reference's step as stock PyTorch on the same B200 (the oracle port's
    modules / autograd / Adam moved to the device, library kernels only - cuBLAS, ATen), fp32 and TF32, eager and
    replayed as one CUDA graph.  A baseline leg: nothing of this repo's CUDA path is on it."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mnist_oracle as O
    B = BATCH_PER_GPU
    image, text, noises = O.synthetic_batch(B, N_LATENTS, 0)
    image, text = image.to(dev), text.to(dev)
    noises = [n.to(dev) for n in noises]
    out = {"batch": B, "what": "oracle/mnist_oracle.py train_step + adam_step on cuda (ATen / cuBLAS), same synthetic batch"}

    def fresh():
        st = {k: v.to(dev) for k, v in O.init_state(N_LATENTS, seed=1234).items()}
        mom = {k: torch.zeros_like(v) for k, v in st.items() if not O.is_buffer(k)}
        vel = {k: torch.zeros_like(v) for k, v in st.items() if not O.is_buffer(k)}
        return st, mom, vel

    def one(st, mom, vel, i):
        # O.train_step without its float(loss) host read-backs (they would serialise the eager run and break graph capture)
        work = {k: (v.clone() if O.is_buffer(k) else v.detach().clone().requires_grad_(True)) for k, v in st.items()}
        losses, _ = O.train_step_losses(work, image, text, noises, work)
        names = [k for k in work if not O.is_buffer(k)]
        gs = torch.autograd.grad(losses[0] + losses[1] + losses[2], [work[k] for k in names], allow_unused=True)
        grads = {k: (torch.zeros_like(work[k]) if g is None else g) for k, g in zip(names, gs)}
        new = O.adam_step(st, grads, mom, vel, i)
        new.update({k: v for k, v in work.items() if O.is_buffer(k)})
        return new

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return B * n / (e0.elapsed_time(e1) * 1e-3)

    prev = torch.backends.cuda.matmul.allow_tf32
    try:
        for name, tf32 in (("fp32_eager", False), ("tf32_eager", True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            st, mom, vel = fresh()
            box = {"st": st}
            for i in range(3):
                box["st"] = one(box["st"], mom, vel, i + 1)

            def step(i, box=box, mom=mom, vel=vel):
                box["st"] = one(box["st"], mom, vel, i + 4)
            out[name] = timed(step, budget_steps)
        # one CUDA graph per step (static inputs: every replay recomputes the same step - the work is identical)
        try:
            torch.backends.cuda.matmul.allow_tf32 = True
            st, mom, vel = fresh()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for i in range(3):
                    one(st, mom, vel, i + 1)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                one(st, mom, vel, 4)
            out["tf32_cuda_graph"] = timed(lambda i: graph.replay(), budget_steps)
        except Exception as exc:   # capture of the autograd step is best effort
            out["tf32_cuda_graph"] = None
            out["cuda_graph_error"] = repr(exc)[:200]
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    out["unit"] = "samples/s"
    return out


def other_configs(torch, dev, log):
    """Short single-GPU rows for the other BASELINE.json configurations (the driver only runs this script): MNIST in tf32
    and with the weak-supervision term flips (configs[4] rule), CelebA (configs[3]) and MultiMNIST (configs[2]) at
    batch 256.  Same timing rules as the headline (CUDA events, warm-up, CUDA-graph replay, rotating inputs)."""
    import numpy as np
    from mvae_b200 import MVAE, MVAETrainer
    out = {}

    def timed(step, warm, n):
        for i in range(warm):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            step(warm + i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    B = BATCH_PER_GPU
    g = torch.Generator().manual_seed(5)
    slots = 24
    try:
        xs32 = [torch.rand(B, 784, generator=g).to(dev) for _ in range(slots)]
        ys = [torch.randint(0, 10, (B,), generator=g).to(dev) for _ in range(slots)]
        m = MVAE(N_LATENTS, precision="tf32", device=dev, seed=1)
        tr = MVAETrainer(m, lr=1e-3, use_cuda_graph=True)
        ms = timed(lambda i: tr.step(xs32[i % slots], ys[i % slots]), 5, 40)
        out["mnist_tf32_b4096"] = {"ms_per_step": ms, "samples_per_s": B / (ms * 1e-3), "precision": "tf32 (per-layer tcgen05 GEMM path)"}
        del m, tr
        m = MVAE(N_LATENTS, precision="bf16", device=dev, seed=1)
        tr = MVAETrainer(m, lr=1e-3, use_cuda_graph=True)
        xs16 = [m.to_act(x) for x in xs32]
        flips = np.random.RandomState(42)

        def weak(i):
            terms, lams = ["joint"], [(1.0, 1.0)]
            if flips.random_sample() < 0.5:
                terms.append("image"); lams.append((1.0, 1.0))
            if flips.random_sample() < 0.5:
                terms.append("text"); lams.append((0.0, 1.0))
            tr.step(xs16[i % slots], ys[i % slots], terms=tuple(terms), lambdas=tuple(lams))
        ms = timed(weak, 24, 60)
        out["mnist_weak_0.5_0.5_b4096"] = {"ms_per_step": ms, "samples_per_s": B / (ms * 1e-3), "precision": "bf16",
                                           "rule": "mnist/modal_weak.py:60-97, per-batch flips from np.random.seed(42)"}
        del m, tr, xs16
        # north-star MLP instantiation (policy 3: "second benchmark row"): Linear+Swish 784-512-512-2n, no
        # normalisation, precision PoE with the prior expert; per-layer tcgen05 GEMMs with bias / Swish / sigmoid-BCE epilogues
        from mvae_b200.mlp import MVAE as SwishMVAE, MVAETrainer as SwishTrainer
        ms_model = SwishMVAE(N_LATENTS, hidden=512, precision="bf16", device=dev, seed=1)
        ms_tr = SwishTrainer(ms_model, use_cuda_graph=True)
        ms = timed(lambda i: ms_tr.step(xs32[i % slots], ys[i % slots]), 5, 40)
        flops = 6.0 * sum(o * i * r for o, i, r in ms_model.linear_shapes(3, 2)) * B
        out["mnist_swish_mlp_512_b4096"] = {"ms_per_step": ms, "samples_per_s": B / (ms * 1e-3), "precision": "bf16",
                                            "model": "784-512-512-2n Linear+Swish, PRECISION PoE + prior expert, n_latents=64",
                                            "gpu_launches_per_step": ms_tr.last_graph_launches,
                                            "gemm_tflops": flops / (ms * 1e-3) / 1e12}
        # configs[4]: per-SAMPLE missing-modality masks inside the captured step (70 % of the rows have the image, 50 % the label)
        hi_m = [(torch.rand(B, generator=g) < 0.7).to(dev) for _ in range(slots)]
        ht_m = [(torch.rand(B, generator=g) < 0.5).to(dev) for _ in range(slots)]
        ms = timed(lambda i: ms_tr.step(xs32[i % slots], ys[i % slots], has_image=hi_m[i % slots], has_text=ht_m[i % slots]), 5, 40)
        out["mnist_swish_mlp_512_masked_b4096"] = {"ms_per_step": ms, "samples_per_s": B / (ms * 1e-3), "precision": "bf16",
                                                   "masks": "per-sample has_image p=0.7 / has_text p=0.5, weights computed on the device "
                                                            "inside the CUDA graph (no host sync)",
                                                   "gpu_launches_per_step": ms_tr.last_graph_launches}
        del ms_model, ms_tr, xs32
    except Exception as exc:
        out["mnist_error"] = repr(exc)[:300]
    for name in ("celeba", "multimnist"):
        try:
            Bc, n = 256, 100
            if name == "celeba":
                from mvae_b200.celeba import MultimodalVAE, CelebATrainer as Trainer
                shape = (3, 64, 64)
                other = lambda: (torch.rand(Bc, 18, generator=g) > 0.5).float()
            else:
                from mvae_b200.multimnist import MultimodalVAE, MultiMNISTTrainer as Trainer
                shape = (1, 50, 50)
                other = lambda: torch.randint(0, 12, (Bc, 4), generator=g)
            model = MultimodalVAE(n_latents=n, precision="bf16", device=dev, seed=1)
            tr = Trainer(model, use_cuda_graph=True)
            nb = max(4, (2 * L2_BYTES) // (Bc * shape[0] * shape[1] * shape[2] * 4) + 1)
            pool = [(torch.rand(Bc, *shape, generator=g).to(dev), other().to(dev)) for _ in range(nb)]
            ms = timed(lambda i: tr.step(*pool[i % nb]), 5, 30)
            out["%s_b256" % name] = {"ms_per_step": ms, "samples_per_s": Bc / (ms * 1e-3), "precision": "bf16", "n_latents": n,
                                     "gpu_launches_per_step": tr.last_graph_launches}
            del model, tr, pool
        except Exception as exc:
            out["%s_error" % name] = repr(exc)[:300]
        log("other_configs %s done" % name)
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------- device arm
def run_device(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d does not match WORLD_SIZE %d" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        log("init_process_group")
        dist.init_process_group("nccl", device_id=dev)
        log("init done")

    import mvae_b200
    from mvae_b200 import MVAE, MVAETrainer
    from mvae_b200.parallel import DataParallelTrainer

    B = args.batch
    model = MVAE(N_LATENTS, precision=args.precision, device=dev, seed=1234 + rank)
    if world > 1:
        trainer = DataParallelTrainer(model, lr=1e-3, use_cuda_graph=not args.no_graph, overlap=not args.no_overlap,
                                      fused=os.environ.get("MVAE_DP_FUSED", "1") != "0")
    else:
        trainer = MVAETrainer(model, lr=1e-3, use_cuda_graph=not args.no_graph)

    es = 2 if args.precision == "bf16" else 4
    n_slots = max(4, (2 * L2_BYTES + B * 784 * es - 1) // (B * 784 * es))
    imgs_u8, labels = synthetic_pool(B, n_slots, 100 + rank, dev, torch)
    pool_x = [model.to_act(imgs_u8[i].to(dev)) for i in range(n_slots)]   # resident in HBM, storage dtype
    pool_y = [labels[i].to(dev) for i in range(n_slots)]
    host_x = [imgs_u8[i].pin_memory() for i in range(min(n_slots, 8))]
    host_y = [labels[i].pin_memory() for i in range(min(n_slots, 8))]
    torch.cuda.synchronize()
    log("pools ready")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    _align_buf = torch.zeros(8, device=dev) if world > 1 else None

    def align_ranks():
        """Device-side alignment right before a timed region: every rank's stream first spins for ~3 ms (so that the host
        has the whole timed loop enqueued before the device gets there), then meets the others in a tiny all-reduce.  The
        start event recorded behind it fires at the same moment on every rank - host-side skew between the ranks (process
        scheduling, the clock sampler's start-up) cannot land inside the timed window."""
        if world > 1:
            torch.cuda._sleep(6_000_000)
            dist.all_reduce(_align_buf)

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    lib = mvae_b200._lib.load()
    # Weak supervision (BASELINE.json configs[4], mnist/modal_weak.py:60-97): the joint term always, the image-only term with
    # probability P1 and the text-only term (lambdas (0,1)) with probability P2 per batch, flips from np.random.seed(42) -
    # identical on every rank, so the data-parallel graphs (one per term set) stay in lock step.
    import numpy as np
    flips = np.random.RandomState(42)

    def step_kwargs():
        if args.weak is None:
            return {}
        terms, lams = ["joint"], [(1.0, 1.0)]
        if flips.random_sample() < args.weak[0]:
            terms.append("image"); lams.append((1.0, 1.0))
        if flips.random_sample() < args.weak[1]:
            terms.append("text"); lams.append((0.0, 1.0))
        return {"terms": tuple(terms), "lambdas": tuple(lams)}

    if args.weak is not None:   # capture the (up to four) term-set graphs before timing
        for terms, lams in ((("joint",), ((1.0, 1.0),)), (("joint", "image"), ((1.0, 1.0),) * 2),
                            (("joint", "text"), ((1.0, 1.0), (0.0, 1.0))), (("joint", "image", "text"), ((1.0, 1.0), (1.0, 1.0), (0.0, 1.0)))):
            trainer.step(pool_x[0], pool_y[0], terms=terms, lambdas=lams)
    # ---- device-resident throughput ("value")
    for i in range(max(args.warmup, 3)):
        losses, _ = trainer.step(pool_x[i % n_slots], pool_y[i % n_slots], ready=True, **step_kwargs())
    log("warmup done")
    barrier()
    log("barrier done")
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = lib.mvae_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    align_ranks()
    e0.record()
    for i in range(args.steps):
        losses, _ = trainer.step(pool_x[i % n_slots], pool_y[i % n_slots], ready=True, **step_kwargs())
    e1.record()
    log("timed loop enqueued")
    barrier()
    clocks = sampler.stop()
    log("timed loop done")
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    eager_launches = int(lib.mvae_launch_count() - l0)
    per_step_launches = trainer.last_graph_launches if not args.no_graph else eager_launches // max(args.steps, 1)
    final_loss = [float(v) for v in losses[:, 0].tolist()]
    value = B * world * args.steps / (ms_total * 1e-3)

    # ---- end to end through the public API with HOST buffers: pinned uint8 images + int64 labels are uploaded
    # every step (copy stream, overlapped with the previous step's compute) and every step's loss tensor is read
    # back to the host (mvae_b200.HostPipeline is the public entry a training script would use).
    from mvae_b200 import HostPipeline
    e2e_steps = max(3, min(args.steps, 400))
    pipe = HostPipeline(trainer)
    for l_host in pipe.run(((host_x[i % len(host_x)], host_y[i % len(host_y)]) for i in range(4))):
        pass
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    align_ranks()
    t0.record()
    n_read = 0
    for l_host in pipe.run(((host_x[i % len(host_x)], host_y[i % len(host_y)]) for i in range(e2e_steps))):
        n_read += 1
        last_host_loss = float(l_host[0, 0])
    t1.record()
    barrier()
    assert n_read == e2e_steps
    e2e_ms = max_over_ranks(t0.elapsed_time(t1))
    log("e2e done")
    e2e_value = B * world * e2e_steps / (e2e_ms * 1e-3)

    def teardown():
        """Release the captured graphs (they hold the NCCL communicator) before destroying the process group;
        a watchdog makes sure a rank can never hang at exit once its JSON line is out."""
        if world == 1:
            return
        threading.Timer(20.0, lambda: os._exit(0)).start()
        try:
            trainer._dp_graphs.clear()
            trainer._graphs.clear()
            import gc
            gc.collect()
            torch.cuda.synchronize()
            dist.destroy_process_group()
        except Exception:
            pass

    if rank != 0:
        teardown()
        return 0

    # ---- live per-kernel durations (CUDA events around every launch, rank 0)
    peaks = load_peaks()
    tt, klw = trainer._norm(("joint", "image", "text"), B, 1.0)
    prof_runs = []
    for i in range(5):
        prof_runs.append(model.profile(pool_x[i % n_slots], pool_y[i % n_slots], tt, ((1.0, 1.0),) * 3, klw, backward=True,
                                       zero_grad=True, adam=trainer.adam))
    agg = {}
    for run in prof_runs[1:]:
        for label, ms in run:
            kind = label.split(":")[0].split("#")[0]
            agg[kind] = agg.get(kind, 0.0) + ms / (len(prof_runs) - 1)
    # per-launch table: algorithmic work of every launch against the roofline that bounds it (tensor peak for the GEMMs,
    # HBM peak for the streaming kernels), durations = CUDA events around the launch (they include ~4 us launch latency)
    n_lat, es_ = N_LATENTS, (2 if args.precision == "bf16" else 4)
    R3 = 3 * B
    gemm_shapes = {  # (rows, out, in) of the Linear behind each label; fwd / dgrad / wgrad are all 2*rows*out*in FLOP
        "image_encoder.net.0.weight": (B, 400, 784), "image_encoder.net.3.weight": (B, 200, 400),
        "image_encoder.net.6.weight": (B, 2 * n_lat, 200), "image_decoder.net.0.weight": (R3, 200, n_lat),
        "image_decoder.net.3.weight": (R3, 400, 200), "image_decoder.net.6.weight": (R3, 784, 400)}
    bn_feats = {"image_encoder.net.1.weight": (B, 400), "image_encoder.net.4.weight": (B, 200),
                "image_decoder.net.1.weight": (R3, 200), "image_decoder.net.4.weight": (R3, 400)}
    # slab-persistent chain kernels (csrc/chain.cu): algorithmic FLOPs of the Linears each launch contains
    enc_mac = 784 * 400 + 400 * 200 + 200 * 2 * n_lat
    dec_mac = n_lat * 200 + 200 * 400 + 400 * 784
    chain_flops = {"chain_enc_fwd": 2.0 * B * enc_mac, "chain_dec_fwd": 2.0 * R3 * dec_mac, "chain_dec_bwd": 2.0 * R3 * dec_mac,
                   "chain_enc_bwd": 2.0 * B * (200 * 2 * n_lat + 400 * 200)}
    tc_burst = peaks["bf16_tflops"] if args.precision == "bf16" else peaks["bf16_tflops"] / 2   # launches are timed in isolation
    hbm_peak = peaks["hbm_gbs"]
    tc_peak = peaks["bf16_tflops_sustained"] if args.precision == "bf16" else peaks["bf16_tflops_sustained"] / 2
    per_launch = []
    n_runs = len(prof_runs) - 1
    for j, (label, _) in enumerate(prof_runs[-1]):
        us = sum(run[j][1] for run in prof_runs[1:]) / n_runs * 1e3
        kind, _, rest = label.partition(":")
        kind = kind.split("#")[0]
        name = rest.split("#")[0]
        row = {"launch": label, "us": round(us, 2)}
        if kind.startswith("gemm") and name in gemm_shapes:
            r_, o_, i_ = gemm_shapes[name]
            fl = 2.0 * r_ * o_ * i_
            row.update(bound="tensor", gflop=round(fl / 1e9, 3), tflops=round(fl / (us * 1e-6) / 1e12, 1),
                       frac=round(fl / (us * 1e-6) / 1e12 / tc_peak, 4))
        elif kind in ("launch_bn_forward", "launch_bn_backward") and name in bn_feats:
            r_, f_ = bn_feats[name]
            by = r_ * f_ * es_ * (2 if kind == "launch_bn_forward" else 3)
            row.update(bound="hbm", mbytes=round(by / 1e6, 2), gbs=round(by / (us * 1e-6) / 1e9, 1),
                       frac=round(by / (us * 1e-6) / 1e9 / hbm_peak, 4))
        elif kind in chain_flops:
            fl = chain_flops[kind]
            row.update(bound="tensor", gflop=round(fl / 1e9, 3), tflops=round(fl / (us * 1e-6) / 1e12, 1),
                       frac=round(fl / (us * 1e-6) / 1e12 / tc_burst, 4))
        elif kind == "launch_tail_forward":
            by = 4096.0 * B
            row.update(bound="hbm", mbytes=round(by / 1e6, 2), gbs=round(by / (us * 1e-6) / 1e9, 1), frac=round(by / (us * 1e-6) / 1e9 / hbm_peak, 4))
        elif kind == "launch_tail_backward":
            by = 3584.0 * B
            row.update(bound="hbm", mbytes=round(by / 1e6, 2), gbs=round(by / (us * 1e-6) / 1e9, 1), frac=round(by / (us * 1e-6) / 1e9 / hbm_peak, 4))
        elif kind == "launch_adam":
            by = 4.0 * 7 * model.flat_params.numel() + (2 * model.flat_params.numel() if args.precision == "bf16" else 0)
            row.update(bound="hbm", mbytes=round(by / 1e6, 2), gbs=round(by / (us * 1e-6) / 1e9, 1), frac=round(by / (us * 1e-6) / 1e9 / hbm_peak, 4))
        per_launch.append(row)
    gemm_ms = sum(v for k, v in agg.items() if k.startswith("gemm") or k.startswith("chain"))
    gemm_launches = sum(1 for label, _ in prof_runs[-1] if label.startswith("gemm") or label.startswith("chain"))
    serial_ms = sum(agg.values())
    flops = F_ALG_PER_SAMPLE * B
    family_tf = flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    peak_tf = peaks["bf16_tflops_sustained"] if args.precision == "bf16" else peaks["bf16_tflops_sustained"] / 2
    # dominant kernel = the single most expensive launch: the last decoder Linear with the fused sigmoid/BCE/dlogits
    # epilogue (gemm_kernel<bf16, BCE>), [3B,400] x [784,400]^T
    chained = "chain_dec_fwd" in agg
    if chained:
        # the decoder's forward chain: three Linears with in-kernel BatchNorm grid barriers + sigmoid/BCE/dlogits epilogue
        dom_label, dom_ms, dom_flops = "chain_dec_fwd", agg["chain_dec_fwd"], chain_flops["chain_dec_fwd"]
        dom_name = ("mvae::chain_kernel<FWD_BN, FWD_BN, BCE, pair> (slab-persistent decoder forward: Linear n-200, 200-400, 400-784 on "
                    "tcgen05 cta_group::2 / TMA with in-kernel BatchNorm grid barriers + fused sigmoid/BCE/dlogits epilogue; the most "
                    "expensive launch)")
    else:
        dom_label, dom_ms, dom_flops = "gemm_fwd_bce", agg.get("gemm_fwd_bce", 0.0), 2.0 * 3 * B * 784 * 400
        dom_name = ("mvae::gemm_kernel<BCE> (tcgen05/TMA GEMM [3B,400]x[784,400]^T + fused sigmoid/BCE/dlogits epilogue; the most "
                    "expensive launch of the per-layer path)")
    achieved_tf = dom_flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    traffic, traffic_source = None, "no ncu summary for this kernel under profiles/"
    try:   # DRAM traffic of the same kernel from the committed ncu --set full summary (stamped with the commit it was taken at)
        with open(os.path.join(ROOT, "profiles", "r02_ncu_chain_summary.json")) as f:
            prof = json.load(f)
        k = prof["kernels"][dom_label]
        traffic = float(k["dram_bytes_read"]) + float(k["dram_bytes_write"])
        traffic_source = "profiles/r02_ncu_chain_summary.json (ncu --set full, commit %s)" % prof.get("commit", "?")
    except Exception:
        pass
    t_roof_us = flops / (peaks["bf16_tflops"] * 1e12) * 1e6 + Q_TAIL_PER_SAMPLE * B / (peaks["hbm_gbs"] * 1e9) * 1e6
    ms_per_step = ms_total / args.steps

    line = {
        "metric": "MVAE train samples/sec (fwd+bwd ELBO)",
        "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "tf32", "data": "synthetic",
        "config": {"workload": "MNIST MVAE (784-400-200-2n MLP + label text, BatchNorm+ReLU), n_latents=64, "
                               "batch %d per GPU, fused PoE + 3-term subsampled ELBO, fwd+bwd+Adam" % B,
                   "batch_per_gpu": B, "global_batch": B * world, "n_latents": N_LATENTS, "parallelism": "dp%d" % world,
                   "gradient_exchange": ("fused NVLink reduce-scatter + all-gather + Adam kernel (csrc/dp.cu)" if getattr(trainer, "fused", False)
                                         else ("NCCL all-reduce + Adam kernel" if world > 1 else "none")),
                   "l2_policy": "inputs rotate through a pool of %d distinct batches (%.0f MB > 2x L2)" % (
                       n_slots, n_slots * B * 784 * es / 1e6),
                   "cuda_graph": not args.no_graph, "precision": args.precision,
                   **({"weak_supervision": {"p_image_only": args.weak[0], "p_text_only": args.weak[1],
                                            "rule": "mnist/modal_weak.py:60-97, np.random.seed(42) flips per batch"}}
                      if args.weak is not None else {})},
        "final_loss_terms": final_loss,
        "gpu_launches": per_step_launches * args.steps,
        "gpu_launches_per_step": per_step_launches,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": B * 784 + B * 8,
                "d2h_bytes_per_step": 3 * 4 * 4, "steps": e2e_steps,
                "path": "HostPipeline(MVAETrainer).run(pinned uint8 images, int64 labels) -> pinned host losses, "
                        "copies overlapped with compute"},
        "roofline": {"bound": "tensor",
                     "kernel": dom_name,
                     "achieved": achieved_tf, "peak": tc_burst, "unit": "TFLOP/s", "frac": achieved_tf / tc_burst,
                     "traffic": traffic, "traffic_source": traffic_source,
                     "peak_source": peaks["source"] + " (burst bf16 GEMM peak: the launch is timed in isolation; tf32 = half)",
                     "flops_per_launch": dom_flops, "launch_ms": dom_ms,
                     "note": "duration = CUDA events around the launch on its stream (includes ~5 us launch latency)"},
        "gemm_family": {"launches_per_step": gemm_launches, "flops_per_step": flops, "ms_per_step_serialised": gemm_ms,
                        "achieved_tflops": family_tf, "frac_of_peak": family_tf / peak_tf},
        "step_roofline": {"t_roof_us": t_roof_us, "t_measured_us": ms_per_step * 1e3,
                          "frac": t_roof_us / (ms_per_step * 1e3),
                          "definition": "SURVEY 8d: F_alg/bf16 burst peak + Q_tail/HBM peak"},
        "per_launch": per_launch,
        "kernel_ms_per_step_serialised": {k: round(v, 5) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])},
        "serialised_ms_per_step": serial_ms,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(torch)
    if world == 1 and not args.no_extra:
        trainer._graphs.clear()
        line["torch_gpu_baseline"] = torch_gpu_baseline(torch, dev)
        line["other_configs"] = other_configs(torch, dev, log)
    print(json.dumps(line), flush=True)
    teardown()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="data parallel: one all-reduce after the whole backward")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the torch_gpu_baseline and other_configs legs")
    ap.add_argument("--weak", type=float, nargs=2, default=None, metavar=("P_IMAGE", "P_TEXT"),
                    help="MNIST weak supervision (mnist/modal_weak.py): per-batch probabilities of the image-only / text-only terms")
    ap.add_argument("--workload", default="mnist", choices=["mnist", "celeba", "multimnist"],
                    help="mnist = the headline config (BASELINE.json configs[1]); celeba / multimnist = configs[3] / [2]")
    args = ap.parse_args()
    args.batch_set = any(a == "--batch" or a.startswith("--batch=") for a in sys.argv[1:])
    if args.workload != "mnist":
        import bench_conv
        if args.impl == "reference":
            return bench_conv.run_reference(args)
        return bench_conv.run(args, log, ClockSampler, load_peaks)
    if args.impl == "reference":
        return run_reference(args)
    return run_device(args)


def _protect_stdout():
    """Libraries (NCCL prints its version banner) write to fd 1; the contract is ONE JSON line on stdout.  Route
    fd 1 to stderr for the whole run and keep a private handle on the real stdout for the final line."""
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w")


if __name__ == "__main__":
    _protect_stdout()
    rc = main()
    sys.stdout.flush()
    sys.stderr.flush()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        os._exit(rc)   # NCCL + CUDA-graph teardown order is not worth a hang at interpreter exit
    sys.exit(rc)
