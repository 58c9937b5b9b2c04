# 8-GPU data-parallel runs (gpurun --gpus 8): the driver's own command line (20 steps), a longer window, replica consistency
set -x
N=${1:-8}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:3}" > gpurun_out/$2.json 2> gpurun_out/$2.err; python - <<P
import json
try:
    d = json.loads(open("gpurun_out/$2.json").read().strip().splitlines()[-1])
    print("$2", "n_gpus", d["n_gpus"], "ms_per_step", round(d["ms_per_step"], 4), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["config"].get("gradient_exchange"))
except Exception as e:
    print("$2 FAILED", e)
P
}
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r02_dp_bench1.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_dp_bench1.json').read().strip().splitlines()[-1]); print('1 gpu ms', d['ms_per_step'], 'value', d['value'])"
run 29601 r02_dp_bench${N} --steps 20 --warmup 5
run 29602 r02_dp_bench${N}_500 --steps 500 --warmup 20 --no-cpu-baseline
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/dp_mnist_check.py > gpurun_out/r02_dp_check${N}.log 2>&1; grep "DP_MNIST" gpurun_out/r02_dp_check${N}.log | head -4
