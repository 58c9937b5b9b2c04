"""Chain kernels vs the per-layer path at batch sizes the parity suite does not visit (losses and gradients after one step, same
seeded inputs and injected noise): python tools/chain_sizes_check.py  (spawns itself with MVAE_CHAIN=0 / 1)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
SIZES = [1024, 1536, 2048, 8192]
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import mnist_oracle as O
    import mvae_b200
    out = {}
    for B in SIZES:
        n = 64
        state = O.perturbed_state(n, 1)
        image, text, noises = O.synthetic_batch(B, n, 1)
        m = mvae_b200.MVAE(n, precision="bf16"); m.load_state_dict(state)
        tr = mvae_b200.MVAETrainer(m, use_cuda_graph=False) if "use_cuda_graph" in mvae_b200.MVAETrainer.__init__.__code__.co_varnames else mvae_b200.MVAETrainer(m)
        eps = torch.stack(noises).cuda()
        l, _ = tr.step(m.to_act(image.cuda()), text.cuda(), eps=eps, update=False)
        torch.cuda.synchronize()
        out[B] = (l[:, 0].cpu(), m.flat_grads.clone().cpu())
    torch.save(out, sys.argv[2])
    sys.exit(0)
import torch
res = {}
for flag in ("0", "1"):
    path = "/tmp/chain_sizes_%s.pt" % flag
    subprocess.run([sys.executable, __file__, "child", path], check=True, env=dict(os.environ, MVAE_CHAIN=flag))
    res[flag] = torch.load(path)
ok = True
for B in SIZES:
    l0, g0 = res["0"][B]; l1, g1 = res["1"][B]
    rel = float((g1 - g0).norm() / g0.norm())
    dl = float((l1 - l0).abs().max() / l0.abs().max())
    print("B=%5d  loss rel diff %.2e  flat-gradient rel-L2 %.2e" % (B, dl, rel))
    ok = ok and dl < 2e-3 and rel < 5e-2
print("CHAIN_SIZES", "OK" if ok else "FAILED")
