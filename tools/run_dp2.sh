# two-GPU checks of the data-parallel MNIST path (run under gpurun --gpus 2): consistency, then scaling bench lines
set -x
N=${NGPU:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_mnist_check.py > gpurun_out/r2_dp_check.log 2>&1; grep "DP_MNIST\|Error\|error" gpurun_out/r2_dp_check.log | head -8
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_dp_bench1.json 2> gpurun_out/r2_dp_bench1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 10 > gpurun_out/r2_dp_bench${N}.json 2> gpurun_out/r2_dp_bench${N}.err
MVAE_DP_FUSED=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 10 > gpurun_out/r2_dp_bench${N}_nccl.json 2> gpurun_out/r2_dp_bench${N}_nccl.err
python - <<P
import json
for f in ("r2_dp_bench1", "r2_dp_bench$N", "r2_dp_bench${N}_nccl"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "n_gpus", d["n_gpus"], "ms_per_step", round(d["ms_per_step"], 4), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["config"].get("parallelism"), d["config"].get("gradient_exchange"))
    except Exception as e:
        print(f, "FAILED", e)
P
tail -3 gpurun_out/r2_dp_bench${N}.err
