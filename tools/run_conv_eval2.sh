set -x
timeout 600 python -m pytest tests/test_celeba_gpu.py tests/test_multimnist_gpu.py -q -x 2>&1 | tail -4 > gpurun_out/conv_tests2.log; cat gpurun_out/conv_tests2.log
timeout 300 python bench.py --workload multimnist --steps 50 --warmup 5 > gpurun_out/bench_mm2.json 2> gpurun_out/bench_mm2.err; tail -2 gpurun_out/bench_mm2.err; cut -c1-400 gpurun_out/bench_mm2.json
timeout 300 python bench.py --workload celeba --steps 50 --warmup 5 > gpurun_out/bench_celeba3.json 2> gpurun_out/bench_celeba3.err; tail -2 gpurun_out/bench_celeba3.err; cut -c1-400 gpurun_out/bench_celeba3.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_conv_check.py > gpurun_out/dp_conv_check.log 2>&1; tail -8 gpurun_out/dp_conv_check.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload celeba --steps 50 --warmup 5 > gpurun_out/bench_celeba_2gpu.json 2> gpurun_out/bench_celeba_2gpu.err; tail -2 gpurun_out/bench_celeba_2gpu.err; cut -c1-300 gpurun_out/bench_celeba_2gpu.json
