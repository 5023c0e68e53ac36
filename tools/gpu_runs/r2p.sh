mkdir -p gpurun_out
timeout 300 python tools/bench_layers.py wgrad --math tf32 > gpurun_out/r2p_wgrad_tf32.log 2>&1; cat gpurun_out/r2p_wgrad_tf32.log
ATOMAI_B200_WGRAD_ORDER=0 timeout 300 python tools/bench_layers.py wgrad --math tf32 > gpurun_out/r2p_wgrad_tf32_o0.log 2>&1; tail -1 gpurun_out/r2p_wgrad_tf32_o0.log
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_dkl_gpu.py -q > gpurun_out/r2p_pytest.log 2>&1
grep -E "^E   .*(assert|Error)|passed|failed|^FAILED|^ERROR" gpurun_out/r2p_pytest.log | head -20
for m in tf32x3 tf32; do timeout 300 python bench.py --workload gram --math $m --steps 10 --warmup 3 --no-baselines > gpurun_out/r2p_gram_$m.json 2>gpurun_out/r2p_gram.err; head -c 250 gpurun_out/r2p_gram_$m.json; echo; done
tail -3 gpurun_out/r2p_gram.err
