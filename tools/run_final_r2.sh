# end-of-round-2 verification with the frozen code: full GPU suite, smoke, default bench line + reference arm
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r02_final_gpu_tests.log; cat gpurun_out/r02_final_gpu_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_final_smoke.log 2>&1; tail -3 gpurun_out/r02_final_smoke.log
timeout 200 python bench.py > gpurun_out/r02_final_bench_1gpu.json 2> gpurun_out/r02_final_bench_1gpu.err; cut -c1-260 gpurun_out/r02_final_bench_1gpu.json
timeout 100 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02_final_bench_reference_arm.json 2>/dev/null; cut -c1-200 gpurun_out/r02_final_bench_reference_arm.json
