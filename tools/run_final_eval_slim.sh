# end-of-round verification with the frozen code: full GPU suite, smoke, the three bench lines
mkdir -p gpurun_out
timeout 170 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/final2_gpu_tests.log; cat gpurun_out/final2_gpu_tests.log
timeout 40 python __graft_entry__.py smoke > gpurun_out/final2_smoke.log 2>&1; tail -4 gpurun_out/final2_smoke.log
timeout 25 python bench.py --workload celeba --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/final2_bench_celeba.json 2>/dev/null; cut -c1-160 gpurun_out/final2_bench_celeba.json
timeout 25 python bench.py --workload multimnist --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/final2_bench_multimnist.json 2>/dev/null; cut -c1-160 gpurun_out/final2_bench_multimnist.json
timeout 60 python bench.py > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.err; cut -c1-300 gpurun_out/final2_bench.json
