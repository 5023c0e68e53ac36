mkdir -p gpurun_out
timeout 300 python tools/debug_x3.py > gpurun_out/r2b_debug.log 2>&1
timeout 600 python -m pytest tests/test_dkl_gpu.py -x -q > gpurun_out/r2b_dkl.log 2>&1
tail -12 gpurun_out/r2b_debug.log; tail -15 gpurun_out/r2b_dkl.log
