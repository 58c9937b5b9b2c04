# two-GPU bench lines with / without the programmatic-launch marks (MVAE_PDL) under the data-parallel phases
# (consistency: tools/dp_mnist_check.py under torchrun, see run_dp2.sh)
N=2
port=29820
for v in 14 0 14 0; do
port=$((port + 1))
MVAE_PDL=$v timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MVAE_PDL=$v n_gpus', d['n_gpus'], 'ms_per_step', round(d['ms_per_step'] * 1e3, 2), 'us  value', round(d['value']), 'e2e', round(d['e2e']['value']))"
done
