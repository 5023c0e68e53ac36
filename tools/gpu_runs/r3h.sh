mkdir -p gpurun_out
timeout 420 python tools/plan_ab.py > gpurun_out/r3h_plan_ab.log 2>&1
cat gpurun_out/r3h_plan_ab.log | cut -c1-200
ATOMAI_B200_NA4=3 ATOMAI_B200_SMEM_KB=224 ATOMAI_B200_MIN_NR=4 timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -q -x 2>&1 | tail -4
