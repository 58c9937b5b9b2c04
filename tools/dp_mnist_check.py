"""Data-parallel MNIST trainer on N GPUs: fused NVLink exchange + Adam (csrc/dp.cu) against the NCCL all-reduce path.
Checks: (1) replicas stay bit-identical, (2) the fused path and the NCCL path agree to fp32 summation order, (3) both
match a single-process run on the concatenated batch's per-shard gradients averaged (via the NCCL path).
    torchrun --nproc-per-node 2 tools/dp_mnist_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import mvae_b200  # noqa: E402
from mvae_b200 import MVAE  # noqa: E402
from mvae_b200.parallel import DataParallelTrainer  # noqa: E402

B, n, steps = 512, 64, 5
g = torch.Generator().manual_seed(100 + rank)
xs = [torch.rand(B, 784, generator=g) for _ in range(steps)]
ys = [torch.randint(0, 10, (B,), generator=g) for _ in range(steps)]
eps = [torch.randn(3, B, n, generator=g) for _ in range(steps)]


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def run(fused, graph, nsteps):
    m = MVAE(n, precision="bf16", device=dev, seed=7)
    tr = DataParallelTrainer(m, lr=1e-3, use_cuda_graph=graph, fused=fused)
    assert tr.fused == fused, getattr(tr, "fused_error", None)
    for i in range(nsteps):
        tr.step(m.to_act(xs[i].to(dev)), ys[i].to(dev), eps=eps[i].to(dev))
    torch.cuda.synchronize()
    return m.flat_params.clone(), m.flat_grads.clone(), tr


# one step: the summed gradient left in the buffer and the updated parameters, fused vs NCCL (and NCCL vs NCCL = the
# run-to-run noise of the bf16 step's atomics)
p_f1, g_f1, tr = run(True, False, 1)
p_n1, g_n1, _ = run(False, False, 1)
p_n1b, g_n1b, _ = run(False, False, 1)
# several steps, eager and graph-replayed: replicas must stay bit-identical
p_f, _, tr5 = run(True, True, steps)
ref = p_f.clone()
dist.broadcast(ref, 0)
same = torch.tensor([int(bool(torch.equal(ref, p_f)))], device=dev)
dist.all_reduce(same, op=dist.ReduceOp.MIN)
err = int(tr5._symm[1][21]) | int(tr._symm[1][21])
if rank == 0:
    noise_g, noise_p = rel(g_n1b, g_n1), rel(p_n1b, p_n1)
    d_g, d_p = rel(g_f1, g_n1), rel(p_f1, p_n1)
    print("DP_MNIST replicas_identical=%d grad fused-vs-nccl %.3e (nccl run-to-run %.3e) params %.3e (run-to-run %.3e) flag_err=%d" % (
        int(same.item()), d_g, noise_g, d_p, noise_p, err))
    ok = int(same.item()) == 1 and err == 0 and d_g < max(5 * noise_g, 1e-5) and d_p < max(5 * noise_p, 1e-6)
    print("DP_MNIST", "OK" if ok else "FAILED")
dist.barrier()
dist.destroy_process_group()
