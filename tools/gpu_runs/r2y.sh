mkdir -p gpurun_out
timeout 300 python tools/bench_layers.py wgrad --math tf32 > gpurun_out/r2y_wgrad.log 2>&1; cat gpurun_out/r2y_wgrad.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_vae_gpu.py tests/test_dkl_gpu.py -q 2>&1 | tail -3
