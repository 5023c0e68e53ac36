mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -q > gpurun_out/r2o_pytest.log 2>&1
grep -E "^E   .*(assert|Error)|passed|failed|^FAILED|^ERROR" gpurun_out/r2o_pytest.log | head -20
timeout 300 python tools/bench_layers.py wgrad --math tf32 > gpurun_out/r2o_wgrad_tf32.log 2>&1; cat gpurun_out/r2o_wgrad_tf32.log
ATOMAI_B200_WGRAD_LOADER=0 timeout 300 python tools/bench_layers.py wgrad --math tf32 > gpurun_out/r2o_wgrad_tf32_old.log 2>&1; tail -1 gpurun_out/r2o_wgrad_tf32_old.log
timeout 300 python tools/bench_layers.py wgrad --math tf32x3 > gpurun_out/r2o_wgrad_x3.log 2>&1; tail -1 gpurun_out/r2o_wgrad_x3.log
