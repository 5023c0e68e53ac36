mkdir -p gpurun_out
timeout 85 python -m pytest tests/test_unet_gpu.py -q -x -k "tf32x3 and test_forward_backward_vs_reference" 2>&1 | tail -3
PLAN_AB_SET=res timeout 45 python tools/plan_ab.py step 2>&1 | grep -E "^step" | head -8
