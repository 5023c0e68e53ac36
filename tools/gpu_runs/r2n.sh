mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_unet_gpu.py tests/test_kernels_gpu.py -q -x > gpurun_out/r2n_pytest.log 2>&1
grep -E "^E   .*(assert|Error)|passed|failed|^FAILED|^ERROR" gpurun_out/r2n_pytest.log | head -20
timeout 600 python bench.py --steps 8 --warmup 3 --no-baselines > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
tail -c 300 gpurun_out/r2n_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2n_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('traffic'))
for k,v in d['math_modes'].items():
    print(k, round(v['images_per_s'],1), round(v['ms_per_step'],2), round(v['train_step_only_ms'],2), {kk:(vv['ms'],vv['algorithmic_tflops']) for kk,vv in v['kernels'].items()})
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -o gpurun_out/r2n_wgrad_tf32 -f python tools/profile_conv.py wgrad c6 c51 bn3 > gpurun_out/r2n_ncu2.log 2>&1; tail -3 gpurun_out/r2n_ncu2.log
