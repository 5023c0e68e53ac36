mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_segmentor_gpu.py -q > gpurun_out/r2q_pytest.log 2>&1
grep -E "^E   .*(assert|Error)|passed|failed|^FAILED|^ERROR" gpurun_out/r2q_pytest.log | head -20
for b in 64 128 256; do for m in tf32x3 tf32; do ATOMAI_B200_GRAM_BLOCK=$b timeout 300 python bench.py --workload gram --math $m --steps 10 --warmup 3 --no-baselines > gpurun_out/r2q_gram_${m}_$b.json 2>gpurun_out/r2q_gram.err; echo "block $b $m: $(head -c 120 gpurun_out/r2q_gram_${m}_$b.json | cut -d, -f2)"; done; done
for f in 1 0; do ATOMAI_B200_FUSE_UP=$f timeout 600 python bench.py --steps 8 --warmup 3 --no-baselines --math tf32x3 > gpurun_out/r2q_bench_up$f.json 2> gpurun_out/r2q_bench.err; tail -c 300 gpurun_out/r2q_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2q_bench_up$f.json').read().strip().splitlines()[-1])
print('fuse_up=$f', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'train-only ms', round(d['config']['train_step_only_ms'],2))
for k,v in d['math_modes'].items():
    print(k, {kk:(vv['ms'],vv['launches']) for kk,vv in v['kernels'].items()})
PY
done
