mkdir -p gpurun_out
timeout 300 python tools/bench_layers.py wgrad --math tf32 > gpurun_out/r2w_wgrad_ct.log 2>&1; cat gpurun_out/r2w_wgrad_ct.log
ATOMAI_B200_WGRAD_LOADER=1 timeout 300 python tools/bench_layers.py wgrad --math tf32 > gpurun_out/r2w_wgrad_rt.log 2>&1; tail -1 gpurun_out/r2w_wgrad_rt.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -q 2>&1 | tail -3
timeout 600 python -m pytest tests/test_unet_gpu.py -q -k "forward_backward" 2>&1 | tail -3
