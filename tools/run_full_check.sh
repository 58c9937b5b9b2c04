# full GPU suite + smoke (1 GPU)
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_full_gpu_tests.log; cat gpurun_out/r2_full_gpu_tests.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; tail -6 gpurun_out/r2_smoke.log
