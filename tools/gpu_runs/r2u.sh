mkdir -p gpurun_out
timeout 300 python tools/bench_layers.py wgrad --math tf32 > gpurun_out/r2u_wgrad_pf1.log 2>&1; cat gpurun_out/r2u_wgrad_pf1.log
ATOMAI_B200_WGRAD_PF=0 timeout 300 python tools/bench_layers.py wgrad --math tf32 > gpurun_out/r2u_wgrad_pf0.log 2>&1; tail -1 gpurun_out/r2u_wgrad_pf0.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -q 2>&1 | tail -3
timeout 200 python tools/gpu_check1.py selftest_next > gpurun_out/r2u_selftest_next.log 2>&1; grep -E "tma_m4|tma_m0_off0_3|sw128_v0" gpurun_out/r2u_selftest_next.log | head -12; tail -3 gpurun_out/r2u_selftest_next.log
