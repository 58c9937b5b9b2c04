set -x
for k in col2im_vec_kernel im2col_vec_kernel col_reduce_kernel bn_act_bwd_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -o gpurun_out/r01_celeba_$k -f python bench.py --workload celeba --steps 1 --warmup 3 --no-graph > gpurun_out/ncu_$k.log 2>&1
  tail -2 gpurun_out/ncu_$k.log | cut -c1-120
done
# the store-bound transposed-conv GEMM (grid 3 x 1536): 7th STORE-epilogue launch of a step -> pick by launch index among gemm kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 60 -c 12 -o gpurun_out/r01_celeba_gemm_kernels -f python bench.py --workload celeba --steps 1 --warmup 3 --no-graph > gpurun_out/ncu_gemm_conv.log 2>&1
tail -2 gpurun_out/ncu_gemm_conv.log | cut -c1-120
ls -la gpurun_out/*.ncu-rep | tail -6
