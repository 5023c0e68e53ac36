mkdir -p gpurun_out
timeout 300 python tools/simt_ab.py > gpurun_out/r3i_simt_ab.log 2>&1; echo "simt_ab rc=$?"
cat gpurun_out/r3i_simt_ab.log | cut -c1-200
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x 2>&1 | tail -3
timeout 200 python tools/plan_ab.py step 2>&1 | grep -E "^step" | head -12
