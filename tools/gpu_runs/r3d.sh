mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet_gpu.py -q -x -k "forward_backward or adam" 2>&1 | tail -2
timeout 600 python bench.py --steps 8 --warmup 3 --no-baselines --math tf32x3 > gpurun_out/r3d_bench.json 2> gpurun_out/r3d_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3d_bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'train-only', round(d['config']['train_step_only_ms'],2), 'launches', d['gpu_launches'])
PY
