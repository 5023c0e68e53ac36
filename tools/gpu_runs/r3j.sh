mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3j_pytest.log 2>&1
grep -E "^E   .*(assert|Error)|passed|failed|^FAILED|^ERROR" gpurun_out/r3j_pytest.log | head -20
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r3j_bench.json 2> gpurun_out/r3j_bench.err
tail -c 300 gpurun_out/r3j_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3j_bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'roof', d['roofline']['frac'], d['roofline'].get('frac_issued'), 'gpu_base', (d.get('gpu_baseline') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), d.get('clocks'))
for k,v in d['math_modes'].items():
    print(k, round(v['images_per_s'],1), round(v['ms_per_step'],2), round(v['train_step_only_ms'],2), {kk:(vv['ms'],vv['algorithmic_tflops']) for kk,vv in v['kernels'].items()})
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r3j_launches.csv python tools/profile_step.py 32 tf32x3 > gpurun_out/r3j_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r3j_launches.csv gpurun_out/r3j_launches.md | tail -28
for wl in rvae imspec gram seg256; do timeout 200 python bench.py --workload $wl --steps 10 --warmup 3 --no-baselines > gpurun_out/r3j_bench_$wl.json 2> gpurun_out/r3j_bench_$wl.err; head -c 170 gpurun_out/r3j_bench_$wl.json; echo; done
