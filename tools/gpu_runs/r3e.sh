mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r3e_pytest.log 2>&1
grep -E "^E   .*(assert|Error)|passed|failed|^FAILED|^ERROR" gpurun_out/r3e_pytest.log | head -20
timeout 300 python tools/gpu_check1.py simt > gpurun_out/r3e_simt.log 2>&1; tail -1 gpurun_out/r3e_simt.log | cut -c1-120
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r3e_bench.json 2> gpurun_out/r3e_bench.err
tail -c 200 gpurun_out/r3e_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3e_bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'roof', d['roofline']['frac'], d['roofline'].get('frac_issued'), 'gpu_base', (d.get('gpu_baseline') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
for k,v in d['math_modes'].items():
    print(k, round(v['images_per_s'],1), round(v['ms_per_step'],2), round(v['train_step_only_ms'],2), {kk:(vv['ms'],vv['algorithmic_tflops']) for kk,vv in v['kernels'].items()})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r3e_launches.csv python tools/profile_step.py 32 tf32x3 > gpurun_out/r3e_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r3e_launches.csv gpurun_out/r3e_launches.md | tail -26
for wl in rvae imspec gram seg256; do timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-baselines > gpurun_out/r3e_bench_$wl.json 2> gpurun_out/r3e_bench_$wl.err; head -c 160 gpurun_out/r3e_bench_$wl.json; echo; done
