mkdir -p gpurun_out
timeout 300 python tools/debug_x3_net2.py > gpurun_out/r2c_debug3.log 2>&1
cat gpurun_out/r2c_debug3.log | tail -45
timeout 600 python -m pytest tests/test_dkl_gpu.py -q > gpurun_out/r2b_dkl.log 2>&1
grep -E "^E   .*(assert|Error)|passed|failed|^FAILED" gpurun_out/r2b_dkl.log | head -30
