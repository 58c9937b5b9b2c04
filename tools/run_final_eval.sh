set -x
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/final_gpu_tests.log; cat gpurun_out/final_gpu_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; tail -6 gpurun_out/final_smoke.log
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -2 gpurun_out/final_bench.err; cut -c1-600 gpurun_out/final_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; cut -c1-300 gpurun_out/final_bench_ref.json
timeout 300 python bench.py --workload celeba --impl reference --steps 5 --warmup 1 > gpurun_out/final_bench_celeba_ref.json 2>/dev/null; cut -c1-200 gpurun_out/final_bench_celeba_ref.json
timeout 300 python bench.py --workload multimnist --impl reference --steps 5 --warmup 1 > gpurun_out/final_bench_mm_ref.json 2>/dev/null; cut -c1-200 gpurun_out/final_bench_mm_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_celeba_launches.csv python bench.py --workload celeba --steps 2 --warmup 3 --no-graph > gpurun_out/ncu_celeba.log 2>&1; tail -1 gpurun_out/ncu_celeba.log | cut -c1-100
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_multimnist_launches.csv python bench.py --workload multimnist --steps 2 --warmup 3 --no-graph > gpurun_out/ncu_mm.log 2>&1; tail -1 gpurun_out/ncu_mm.log | cut -c1-100
MVAE_PDL=1 timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('PDL=1 ms_per_step', d['ms_per_step'])"
