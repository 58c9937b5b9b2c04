# chain kernels: parity tests that exercise them, phase timeline, bench line (run on the B200 box)
set -x
timeout 900 python -m pytest tests/test_mnist_step_gpu.py tests/test_module_surface_gpu.py -q -x ${PYTEST_K:+-k "$PYTEST_K"} 2>&1 | tail -15 > gpurun_out/r2_chain_tests.log; cat gpurun_out/r2_chain_tests.log
timeout 120 python tools/chain_timeline.py > gpurun_out/r2_chain_timeline.log 2>&1; grep -c . gpurun_out/r2_chain_timeline.log
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/r2_bench_chain.json 2> gpurun_out/r2_bench_chain.err; tail -3 gpurun_out/r2_bench_chain.err; python - <<'P'
import json
d = json.loads(open("gpurun_out/r2_bench_chain.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "value", d["value"], "launches/step", d.get("gpu_launches_per_step"), "e2e", d["e2e"]["value"])
print(json.dumps(d.get("kernel_ms_per_step_serialised")))
P
