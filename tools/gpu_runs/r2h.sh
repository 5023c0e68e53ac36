mkdir -p gpurun_out
timeout 300 python tools/bench_layers.py --math tf32x3 > gpurun_out/r2h_layers_x3.log 2>&1; cat gpurun_out/r2h_layers_x3.log
timeout 300 python tools/bench_layers.py --math tf32 > gpurun_out/r2h_layers_tf32.log 2>&1; cat gpurun_out/r2h_layers_tf32.log
timeout 300 python bench.py --workload gram --steps 10 --warmup 3 > gpurun_out/r2h_gram_x3.json 2>gpurun_out/r2h_gram.err; head -c 300 gpurun_out/r2h_gram_x3.json; echo
timeout 300 python bench.py --workload gram --math tf32 --steps 10 --warmup 3 > gpurun_out/r2h_gram_tf32.json 2>>gpurun_out/r2h_gram.err; head -c 300 gpurun_out/r2h_gram_tf32.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2h_launches_x3.csv python tools/profile_step.py 32 tf32x3 > gpurun_out/r2h_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r2h_launches_x3.csv gpurun_out/r2h_launches_x3.md | tail -32
timeout 600 python bench.py --steps 8 --warmup 3 --no-baselines > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2h_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k,v in d['math_modes'].items():
    print(k, v['images_per_s'], v['ms_per_step'], v['train_step_only_ms'], {kk:(vv['ms'],vv['algorithmic_tflops']) for kk,vv in v['kernels'].items()})
PY
