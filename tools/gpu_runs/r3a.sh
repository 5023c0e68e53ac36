mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_vae_gpu.py tests/test_unet_gpu.py -q -x 2>&1 | tail -3
timeout 300 python bench.py --workload rvae --steps 10 --warmup 3 --no-baselines > gpurun_out/r3a_bench_rvae.json 2> gpurun_out/r3a_bench_rvae.err; head -c 200 gpurun_out/r3a_bench_rvae.json; echo
timeout 600 python bench.py --steps 8 --warmup 3 --no-baselines --math tf32x3 > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3a_bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1))
for k,v in d['math_modes'].items():
    print(k, {kk:(vv['ms'],vv['launches']) for kk,vv in v['kernels'].items()})
PY
