set -x
timeout 600 python -m pytest tests/test_mlp_gpu.py tests/test_masked_conv_gpu.py -q 2>&1 | tail -40 > gpurun_out/r2_mask_tests.log; cat gpurun_out/r2_mask_tests.log
