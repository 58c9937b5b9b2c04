#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in celeba multimnist; do
  for st in 2 4 6; do
    MVAE_GATHER_STAGES=$st timeout 200 python bench.py --workload $w --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/gs.json 2> gpurun_out/gs.err
    python - <<P
import json
try:
    d=json.loads(open("gpurun_out/gs.json").read().strip().splitlines()[-1])
    print("$w stages=$st", round(d["value"]), "samples/s", round(d["ms_per_step"],4), "ms")
except Exception as e:
    print("$w stages=$st FAILED", e); print(open("gpurun_out/gs.err").read()[-800:])
P
  done
done
