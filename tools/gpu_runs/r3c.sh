mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_c1|conv_pix|wgrad_c1|wgrad_small" -o gpurun_out/r3c_thin -f python tools/profile_thin.py > gpurun_out/r3c_ncu.log 2>&1; tail -2 gpurun_out/r3c_ncu.log
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_segmentor_gpu.py -q -x 2>&1 | tail -3
timeout 600 python bench.py --steps 8 --warmup 3 --no-baselines --math tf32x3 > gpurun_out/r3c_bench.json 2> gpurun_out/r3c_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3c_bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'])
PY
