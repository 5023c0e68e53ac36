mkdir -p gpurun_out
timeout 300 python tools/debug_x3.py > gpurun_out/r2d_debug.log 2>&1
tail -14 gpurun_out/r2d_debug.log
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_dkl_gpu.py tests/test_kernels_gpu.py -q -x > gpurun_out/r2d_pytest.log 2>&1
grep -E "^E   .*(assert|Error)|passed|failed|^FAILED" gpurun_out/r2d_pytest.log | head -20
timeout 600 python bench.py --steps 8 --warmup 3 --no-baselines > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
tail -3 gpurun_out/r2d_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2d_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k,v in d['math_modes'].items():
    print(k, v['images_per_s'], v['ms_per_step'], v['train_step_only_ms'], {kk:(vv['ms'],vv['algorithmic_tflops']) for kk,vv in v['kernels'].items()})
PY
