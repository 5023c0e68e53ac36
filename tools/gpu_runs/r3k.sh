mkdir -p gpurun_out
PLAN_AB_SET=res timeout 150 python tools/plan_ab.py layers > gpurun_out/r3k_plan_res.log 2>&1
cat gpurun_out/r3k_plan_res.log | cut -c1-200
ATOMAI_B200_RES_NA4=1 timeout 120 python -m pytest tests/test_unet_gpu.py -q -x -k "test_forward_backward_vs_reference and tf32x3 and (unet_default_3c_512 or unet_default_3c_128)" 2>&1 | tail -3
