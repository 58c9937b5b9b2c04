# round-2 evidence run on the B200 box: launch list of one step, ncu --set full of the step's own kernels, bench line.
set -x
python tools/step_prof.py bf16 4096 3 > gpurun_out/r02_sp.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_step_launches_bf16.csv python tools/step_prof.py bf16 4096 3 > gpurun_out/r02_ncu_launches.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"chain_kernel|tail_|adam_kernel|textdec|textenc" --launch-skip 26 --launch-count 13 -o gpurun_out/r02_step_full -f python tools/step_prof.py bf16 4096 3 > gpurun_out/r02_ncu_full.log 2>&1; tail -2 gpurun_out/r02_ncu_full.log
if [ -z "$SKIP_BENCH" ]; then
timeout 600 python bench.py --steps 300 --warmup 20 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -2 gpurun_out/r02_bench_1gpu.err; cut -c1-400 gpurun_out/r02_bench_1gpu.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; cut -c1-300 gpurun_out/r02_bench_reference_arm.json
fi
