#!/bin/bash
# implicit-GEMM conv path: bf16 parity subset + step time with and without it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 280 python -m pytest tests/test_celeba_gpu.py tests/test_multimnist_gpu.py -x -q -m gpu -k "step_matches_oracle or ragged or graph_replay" 2>&1 | tail -5
for w in celeba multimnist; do
  for imp in 1 0; do
    MVAE_IMPLICIT_CONV=$imp timeout 200 python bench.py --workload $w --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/imp_${w}_${imp}.json 2> gpurun_out/imp_${w}_${imp}.err
    python - <<P
import json
try:
    d=json.loads(open("gpurun_out/imp_${w}_${imp}.json").read().strip().splitlines()[-1])
    print("$w implicit=$imp", round(d["value"]), "samples/s", round(d["ms_per_step"],4), "ms", d["gpu_launches_per_step"], "launches")
except Exception as e:
    print("$w implicit=$imp FAILED", e); print(open("gpurun_out/imp_${w}_${imp}.err").read()[-1500:])
P
  done
done
