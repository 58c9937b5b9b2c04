"""bench.py --workload celeba|multimnist: train-step throughput of the convolutional MVAEs (BASELINE.json configs 3/4).

Same contract as bench.py's device arm: W untimed warm-up steps, K timed steps bracketed by barrier + synchronize, CUDA
events, max over ranks, ONE JSON line from rank 0.  Inputs rotate through a pool larger than 2x L2.  `e2e` uploads
pinned host batches (fp32 image + second modality) every step and reads the loss accumulators back every step.
The roofline is the step-level one of SURVEY.md 8d: algorithmic GEMM FLOPs (encoders once, decoders per term; forward +
dgrad + wgrad, no dgrad into the input image) over the measured bf16 tensor peak.
"""
from __future__ import annotations

import json
import os
import sys
import threading

ROOT = os.path.dirname(os.path.abspath(__file__))
L2_BYTES = 126 * 1024 * 1024


def algorithmic_flops(model, n_terms=3, n_img_terms=2) -> float:
    """GEMM FLOPs per sample per step any correct implementation must do."""
    from mvae_b200 import _ops
    n = model.n_latents
    f = 0.0
    for li, (pre, ci, co, k, s, p, hin, bn) in enumerate(model.ENC_CONVS):
        ho = _ops.out_size(hin, k, s, p)
        mac = ho * ho * co * k * k * ci
        f += 2 * mac * (2 if li == 0 else 3)          # fwd + wgrad (+ dgrad except into the image)
    for pre, ci, co, k, s, p, hout, bn in model.DEC_CONVS:
        hin = _ops.out_size(hout, k, s, p)
        f += n_terms * 2 * hin * hin * ci * k * k * co * 3
    for (o, i, reps) in model.linear_shapes(n_terms, n_img_terms):
        f += reps * 2 * o * i * 3
    return f


def run_reference(args):
    """CPU arm for the conv workloads: the oracle port (oracle/{celeba,multimnist}_oracle.py, pinned to the reference's
    fixtures) on all host cores, a bounded sample of steps of the same batch size, Adam included."""
    import time
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mnist_oracle as MN
    O = __import__("celeba_oracle" if args.workload == "celeba" else "multimnist_oracle")
    torch.set_num_threads(os.cpu_count() or 1)
    B, n = (args.batch if args.batch_set else 256), 100
    state = O.init_state(n, seed=1234)
    image, other, noises = O.synthetic_batch(B, n, 0)
    mom = {k: torch.zeros_like(v) for k, v in state.items() if not O.is_buffer(k)}
    vel = {k: torch.zeros_like(v) for k, v in state.items() if not O.is_buffer(k)}
    steps, warmup = min(args.steps, 10), min(args.warmup, 2)
    p, step_no = state, 0

    def one():
        nonlocal p, step_no
        step_no += 1
        _, grads, bufs, _ = O.train_step(p, image, other, [torch.randn_like(x) for x in noises])
        p = MN.adam_step(p, grads, mom, vel, step_no)
        p.update(bufs)

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    value = B * steps / dt
    print(json.dumps({
        "impl": "reference", "metric": "MVAE train samples/sec (fwd+bwd ELBO)", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s MVAE n_latents=100 batch=%d 3-term ELBO step (fwd+bwd+Adam), CPU oracle port (Dropout off)" % (args.workload, B)},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "%d steps of batch %d after %d warm-up" % (steps, B, warmup)},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
    return 0


def run(args, log, ClockSampler, load_peaks):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import mvae_b200
    lib = mvae_b200._lib.load()
    B = args.batch if args.batch_set else 256
    n = 100
    g = torch.Generator().manual_seed(100 + rank)
    if args.workload == "celeba":
        from mvae_b200.celeba import MultimodalVAE, CelebATrainer
        model = MultimodalVAE(n_latents=n, precision=args.precision, device=dev, seed=1234 + rank)
        trainer = CelebATrainer(model, use_cuda_graph=not args.no_graph, overlap=not args.no_overlap)
        img_shape, sample_bytes = (3, 64, 64), 3 * 64 * 64 * 4
        make_other = lambda: (torch.rand(B, 18, generator=g) > 0.5).float()
        workload = ("CelebA MVAE (conv 3-32-64-128-256 + FC 6400-1024-2n, ConvT decoder, 18 attributes), n_latents=100, "
                    "batch %d per GPU, 3-term ELBO fwd+bwd+Adam, Dropout 0.1" % B)
    else:
        from mvae_b200.multimnist import MultimodalVAE, MultiMNISTTrainer
        model = MultimodalVAE(n_latents=n, precision=args.precision, device=dev, seed=1234 + rank)
        trainer = MultiMNISTTrainer(model, use_cuda_graph=not args.no_graph, overlap=not args.no_overlap)
        img_shape, sample_bytes = (1, 50, 50), 50 * 50 * 4
        make_other = lambda: torch.randint(0, 12, (B, 4), generator=g)
        workload = ("MultiMNIST MVAE (conv 1-32-64-128-256 + FC, ConvT decoder, biGRU text encoder, 2-layer GRU text decoder), "
                    "n_latents=100, batch %d per GPU, 3-term ELBO fwd+bwd+Adam" % B)
    n_slots = max(4, (2 * L2_BYTES + B * sample_bytes - 1) // (B * sample_bytes))
    host = [(torch.rand(B, *img_shape, generator=g).pin_memory(), make_other().pin_memory()) for _ in range(min(n_slots, 8))]
    pool = [(torch.rand(B, *img_shape, generator=g).to(dev), make_other().to(dev)) for _ in range(n_slots)]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for i in range(max(args.warmup, 3)):
        trainer.step(*pool[i % n_slots])
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = lib.mvae_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        trainer.step(*pool[i % n_slots])
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    eager = int(lib.mvae_launch_count() - l0)
    per_step = trainer.last_graph_launches if not args.no_graph else eager // max(args.steps, 1)
    losses = trainer.losses()
    value = B * world * args.steps / (ms_total * 1e-3)

    # ---- end to end: pinned host batches uploaded on a copy stream one step ahead, accumulators read back every step
    e2e_steps = max(3, min(args.steps, 100))
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    dbuf = [(torch.empty(B, *img_shape, device=dev), torch.empty_like(pool[0][1])) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    hacc = torch.empty(3, 4).pin_memory()

    def upload(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[s])
            dbuf[s][0].copy_(host[i % len(host)][0], non_blocking=True)
            dbuf[s][1].copy_(host[i % len(host)][1], non_blocking=True)
            ready[s].record(copy_stream)

    for s in range(2):
        freed[s].record(main)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    upload(0)
    for i in range(e2e_steps):
        if i + 1 < e2e_steps:
            upload(i + 1)
        s = i % 2
        main.wait_event(ready[s])
        acc = trainer.step(*dbuf[s])
        freed[s].record(main)
        hacc.copy_(acc, non_blocking=True)
        main.synchronize()
        _ = float(hacc[0, 0])
    t1.record()
    barrier()
    e2e_ms = max_over_ranks(t0.elapsed_time(t1))
    e2e_value = B * world * e2e_steps / (e2e_ms * 1e-3)
    h2d = int(B * sample_bytes + host[0][1].numel() * host[0][1].element_size())

    def teardown():
        if world == 1:
            return
        threading.Timer(20.0, lambda: os._exit(0)).start()
        try:
            trainer.teardown()
            import gc
            gc.collect()
            torch.cuda.synchronize()
            dist.destroy_process_group()
        except Exception:
            pass

    if rank != 0:
        teardown()
        return 0
    peaks = load_peaks()
    tc_peak = float(peaks.get("bf16_tflops_sustained", 1411.1)) * (1.0 if args.precision == "bf16" else 0.5)
    flops = algorithmic_flops(model) * B
    t_step = ms_total / args.steps * 1e-3
    out = {
        "metric": "MVAE train samples/sec (fwd+bwd ELBO)", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": workload, "batch_per_gpu": B, "global_batch": B * world, "n_latents": n,
                   "parallelism": "dp%d" % world, "cuda_graph": not args.no_graph,
                   "l2_policy": "inputs rotate through a pool of %d batches (> 2x L2)" % n_slots},
        "final_loss_terms": [l[0] for l in losses],
        "gpu_launches": per_step * args.steps, "gpu_launches_per_step": per_step,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 48, "steps": e2e_steps,
                "path": "Trainer.step(pinned fp32 images + second modality uploaded one step ahead on a copy stream) -> "
                        "pinned host loss accumulators read every step"},
        "roofline": {"bound": "tensor", "kernel": "whole step (all GEMM launches; see profiles/ for the per-kernel list)",
                     "achieved": flops / t_step / 1e12, "peak": tc_peak, "unit": "TFLOP/s",
                     "frac": flops / t_step / 1e12 / tc_peak, "traffic": None,
                     "flops_per_step": flops, "peak_source": "MEASURED_PEAKS.json sustained bf16 (tf32 = half)"},
    }
    print(json.dumps(out), flush=True)
    teardown()
    return 0
