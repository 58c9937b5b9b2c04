# the specialised tail kernels: parity tests, then ncu durations + stall reasons of the tail kernels only, then the bench line
set -x
timeout 900 python -m pytest tests/test_mnist_step_gpu.py tests/test_module_surface_gpu.py -q -x 2>&1 | tail -5
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tail_" --launch-skip 2 --launch-count 4 -o gpurun_out/r02_tail_fast -f python tools/step_prof.py bf16 4096 3 > gpurun_out/r02_tail_fast.log 2>&1; tail -2 gpurun_out/r02_tail_fast.log
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-extra > gpurun_out/r2_bench_tail_fast.json 2>/dev/null
python - <<P
import json
d = json.loads(open("gpurun_out/r2_bench_tail_fast.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["value"], json.dumps({k: v for k, v in d.get("kernel_ms_per_step_serialised", {}).items()}))
P
