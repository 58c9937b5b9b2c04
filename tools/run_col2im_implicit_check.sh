#!/bin/bash
# First GPU job of the next round: the hosts' use of mvae_convt_class_gemm (MVAE_IMPLICIT_COL2IM=1) - parity subset of the
# two conv suites, then step time with and without it, then a full ncu capture of the three gather instantiations.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
MVAE_TEST_CONVT_MERGED=1 timeout 60 python -m pytest tests/test_celeba_gpu.py -x -q -m gpu -k "transposed_conv" 2>&1 | tail -5
MVAE_TEST_UNVERIFIED=1 timeout 60 python -m pytest tests/test_multimnist_gpu.py tests/test_module_surface_gpu.py -x -q -m gpu -k "generate or uint8" 2>&1 | tail -5
MVAE_IMPLICIT_COL2IM=1 timeout 280 python -m pytest tests/test_celeba_gpu.py tests/test_multimnist_gpu.py -x -q -m gpu \
  -k "step_matches_oracle or ragged or graph_replay or fixture or autograd" 2>&1 | tail -5
for w in celeba multimnist; do
  for imp in 2 1 0; do
    MVAE_CONVT_MERGED=$([ $imp = 2 ] && echo 1 || echo 0) MVAE_IMPLICIT_COL2IM=$([ $imp = 0 ] && echo 0 || echo 1) timeout 200 python bench.py --workload $w --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/c2i_${w}_${imp}.json 2> gpurun_out/c2i_${w}_${imp}.err
    python - <<P
import json
try:
    d=json.loads(open("gpurun_out/c2i_${w}_${imp}.json").read().strip().splitlines()[-1])
    print("$w implicit_col2im=$imp", round(d["value"]), "samples/s", round(d["ms_per_step"],4), "ms", d["gpu_launches_per_step"], "launches")
except Exception as e:
    print("$w implicit_col2im=$imp FAILED", e); print(open("gpurun_out/c2i_${w}_${imp}.err").read()[-1500:])
P
  done
done
# demangled names carry "(int)1, (int)0, (int)1": match on the gather template argument
#MVAE_IMPLICIT_COL2IM=1 timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
#  -k "regex:gemm_kernel<\(int\)1, \(int\)., \(int\)[12]>" -s 24 -c 6 -o gpurun_out/r02_celeba_gather_gemm -f \
#  python bench.py --workload celeba --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_gather.log 2>&1
#tail -2 gpurun_out/ncu_gather.log | cut -c1-160
