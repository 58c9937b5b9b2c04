"""torchrun --nproc-per-node 2 tools/dp_conv_check.py: data-parallel conv trainers (CelebA, MultiMNIST) on 2 GPUs.
Checks (i) the all-reduced gradient equals the sum of the per-rank gradients computed without DP, (ii) replicas stay
bit-identical after Adam steps, (iii) overlap on/off give the same gradient."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
import mvae_b200  # noqa
ok = True


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def check(name, Model, Trainer, batch_fn, n=16, B=8):
    global ok
    import importlib
    data = [batch_fn(B, n, 10 + r) for r in range(world)]          # every rank can rebuild every shard
    m = Model(n_latents=n, precision="tf32", dropout_p=0.0)
    grads = {}
    for overlap in (True, False):
        tr = Trainer(m, overlap=overlap)                            # broadcasts rank 0's parameters
        img, oth, noises = data[rank]
        tr.step(img.to(dev), oth.to(dev), eps=torch.stack(noises).to(dev), adam=False)
        torch.cuda.synchronize()
        grads[overlap] = m.flat_grads.clone()
        m.flat_grads.zero_()
    # reference: the same replica runs every shard locally without DP and sums the gradients
    tot = torch.zeros_like(m.flat_grads)
    tr1 = Trainer.__new__(Trainer)
    Trainer.__init__(tr1, m)
    tr1.world, tr1.comm_stream = 1, None
    state = (m.flat_buffers.clone(), m.flat_nbt.clone())
    for r in range(world):
        img, oth, noises = data[r]
        m.flat_grads.zero_()
        tr1.step(img.to(dev), oth.to(dev), eps=torch.stack(noises).to(dev), adam=False)
        torch.cuda.synchronize()
        tot += m.flat_grads
    m.flat_grads.zero_()
    e1, e2 = rel(grads[True], tot), rel(grads[False], grads[True])
    # replicas stay identical through Adam steps
    tr = Trainer(m, overlap=True)
    for it in range(3):
        img, oth, noises = data[rank]
        tr.step(img.to(dev), oth.to(dev))
    torch.cuda.synchronize()
    mine = m.flat_params.clone()
    ref = mine.clone()
    dist.broadcast(ref, 0)
    same = bool(torch.equal(mine, ref))
    good = e1 < 3e-3 and e2 < 3e-3 and same   # run-to-run noise of the tf32 path itself is ~5e-4 (reproducibility)
    ok = ok and good
    print("[rank %d] %s: allreduce-vs-local-sum %.2e, overlap-vs-not %.2e, replicas identical %s -> %s" %
          (rank, name, e1, e2, same, "OK" if good else "FAIL"), flush=True)


import celeba_oracle as CO, multimnist_oracle as MO
from mvae_b200.celeba import MultimodalVAE as CV, CelebATrainer
from mvae_b200.multimnist import MultimodalVAE as MV, MultiMNISTTrainer
check("celeba", CV, CelebATrainer, CO.synthetic_batch)
check("multimnist", MV, MultiMNISTTrainer, MO.synthetic_batch)
dist.barrier()
torch.cuda.synchronize()
print("DP_CONV_CHECK", "PASS" if ok else "FAIL", flush=True)
os._exit(0 if ok else 1)
