# round-2 evidence run on the B200 box: launch list of one step, ncu --set full of the step's own kernels, sanitizer pass
set -x
python tools/step_prof.py bf16 4096 3 > gpurun_out/r02_sp.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_step_launches_bf16.csv python tools/step_prof.py bf16 4096 3 > gpurun_out/r02_ncu_launches.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"chain_kernel|tail_|adam_kernel|textdec|textenc" --launch-skip 22 --launch-count 11 -o gpurun_out/r02_step_full -f python tools/step_prof.py bf16 4096 3 > gpurun_out/r02_ncu_full.log 2>&1; tail -2 gpurun_out/r02_ncu_full.log
timeout 600 compute-sanitizer --tool memcheck python tools/step_prof.py bf16 256 2 > gpurun_out/r02_sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/r02_sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python tools/step_prof.py bf16 256 2 > gpurun_out/r02_sanitizer_racecheck.log 2>&1; tail -4 gpurun_out/r02_sanitizer_racecheck.log
