set -x
N=${1:-8}
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:3}" > gpurun_out/$2.json 2> gpurun_out/$2.err; tail -1 gpurun_out/$2.json | cut -c1-230; }
run 29603 r01_scale_${N}gpu_celeba --workload celeba --steps 40 --warmup 5
run 29604 r01_scale_${N}gpu_multimnist --workload multimnist --steps 40 --warmup 5
