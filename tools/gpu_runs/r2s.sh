mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -q > gpurun_out/r2s_pytest.log 2>&1
grep -E "^E   .*(assert|Error)|passed|failed|^FAILED|^ERROR" gpurun_out/r2s_pytest.log | head -20
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2s_launches.csv python tools/profile_step.py 32 tf32x3 > gpurun_out/r2s_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r2s_launches.csv gpurun_out/r2s_launches.md | tail -30
