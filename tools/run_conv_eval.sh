set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/full_gpu_tests.log
cat gpurun_out/full_gpu_tests.log
timeout 300 python bench.py --workload multimnist --steps 50 --warmup 5 > gpurun_out/bench_mm1.json 2> gpurun_out/bench_mm1.err; tail -2 gpurun_out/bench_mm1.err; cat gpurun_out/bench_mm1.json
timeout 300 python bench.py --workload celeba --steps 50 --warmup 5 > gpurun_out/bench_celeba2.json 2> gpurun_out/bench_celeba2.err; tail -2 gpurun_out/bench_celeba2.err; cat gpurun_out/bench_celeba2.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_celeba_launches.csv python bench.py --workload celeba --steps 2 --warmup 3 --no-graph > gpurun_out/ncu_celeba.log 2>&1; tail -2 gpurun_out/ncu_celeba.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_multimnist_launches.csv python bench.py --workload multimnist --steps 2 --warmup 3 --no-graph > gpurun_out/ncu_mm.log 2>&1; tail -2 gpurun_out/ncu_mm.log
