mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1
grep -E "^E   .*(assert|Error)|passed|failed|^FAILED|^ERROR" gpurun_out/r2f_pytest.log | head -40
for wl in rvae imspec gram; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/r2f_bench_$wl.json 2> gpurun_out/r2f_bench_$wl.err
  tail -c 600 gpurun_out/r2f_bench_$wl.err; head -c 1200 gpurun_out/r2f_bench_$wl.json; echo
done
timeout 600 python bench.py --workload seg256 --steps 10 --warmup 3 --no-baselines > gpurun_out/r2f_bench_seg256.json 2> gpurun_out/r2f_bench_seg256.err
head -c 600 gpurun_out/r2f_bench_seg256.json; echo
