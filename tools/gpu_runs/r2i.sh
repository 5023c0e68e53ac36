mkdir -p gpurun_out
timeout 300 python tools/debug_x3.py > gpurun_out/r2i_debug.log 2>&1; tail -14 gpurun_out/r2i_debug.log
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_kernels_gpu.py tests/test_vae_gpu.py -q -x > gpurun_out/r2i_pytest.log 2>&1
grep -E "^E   .*(assert|Error)|passed|failed|^FAILED" gpurun_out/r2i_pytest.log | head -10
for tma in 1 0; do
 for m in tf32 tf32x3; do
  echo "== TMA=$tma math=$m"
  ATOMAI_B200_TMA=$tma timeout 300 python tools/bench_layers.py fwd dgrad --math $m > gpurun_out/r2i_layers_${m}_tma$tma.log 2>&1; cat gpurun_out/r2i_layers_${m}_tma$tma.log
 done
done
timeout 300 python tools/bench_layers.py wgrad --math tf32 > gpurun_out/r2i_wgrad_tf32.log 2>&1; cat gpurun_out/r2i_wgrad_tf32.log
timeout 300 python tools/bench_layers.py wgrad --math tf32x3 > gpurun_out/r2i_wgrad_x3.log 2>&1; cat gpurun_out/r2i_wgrad_x3.log
