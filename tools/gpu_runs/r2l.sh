mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2l_pytest.log 2>&1
grep -E "^E   .*(assert|Error)|passed|failed|^FAILED|^ERROR" gpurun_out/r2l_pytest.log | head -40
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
tail -c 400 gpurun_out/r2l_bench.err; head -c 400 gpurun_out/r2l_bench.json; echo
timeout 400 python bench.py --impl torch-cuda --steps 8 --warmup 3 > gpurun_out/r2l_bench_torchcuda.json 2> gpurun_out/r2l_bench_torchcuda.err
head -c 600 gpurun_out/r2l_bench_torchcuda.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2l_launches_x3.csv python tools/profile_step.py 32 tf32x3 > gpurun_out/r2l_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r2l_launches_x3.csv gpurun_out/r2l_launches_x3.md | tail -40
