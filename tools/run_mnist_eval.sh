set -x
timeout 600 python -m pytest tests/test_mnist_step_gpu.py tests/test_module_surface_gpu.py -q -x -k "not curve" 2>&1 | tail -12 > gpurun_out/mnist_tests.log; cat gpurun_out/mnist_tests.log
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/bench_fz.json 2> gpurun_out/bench_fz.err; tail -3 gpurun_out/bench_fz.err; python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_fz.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "value", d["value"], "launches/step", d.get("gpu_launches_per_step"))
print(json.dumps(d.get("kernel_ms_per_step_serialised")))
P
MVAE_FUSE_BN_EPI=0 timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/bench_nofz.json 2> gpurun_out/bench_nofz.err; python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_nofz.json").read().strip().splitlines()[-1])
print("UNFUSED ms_per_step", d["ms_per_step"], "value", d["value"], "launches/step", d.get("gpu_launches_per_step"))
P
