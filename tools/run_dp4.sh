# four-GPU check of the data-parallel MNIST path (gpurun --gpus 4): replica consistency, then the driver's own command line
N=4
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 tools/dp_mnist_check.py > gpurun_out/r02_dp_check4.log 2>&1; grep "DP_MNIST" gpurun_out/r02_dp_check4.log | head -4
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_dp_bench4.json 2> gpurun_out/r02_dp_bench4.err
python - <<P
import json
d = json.loads(open("gpurun_out/r02_dp_bench4.json").read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "ms_per_step", round(d["ms_per_step"], 4), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["config"].get("gradient_exchange"))
P
