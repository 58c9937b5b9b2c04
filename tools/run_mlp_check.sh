# north-star MLP instantiation: parity tests (fused Swish / BCE epilogues, step vs oracle) and its bench row
set -x
timeout 600 python -m pytest tests/test_mlp_gpu.py -q 2>&1 | tail -30 > gpurun_out/r2_mlp_tests.log; cat gpurun_out/r2_mlp_tests.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_bench_mlp.json 2> gpurun_out/r2_bench_mlp.err; tail -5 gpurun_out/r2_bench_mlp.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2_bench_mlp.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
print(json.dumps(d.get("other_configs"), indent=1))
P
