mkdir -p gpurun_out
timeout 300 python tools/gpu_check1.py simt > gpurun_out/r2x_simt.log 2>&1; tail -2 gpurun_out/r2x_simt.log | cut -c1-1500
timeout 900 python -m pytest tests/test_vae_gpu.py tests/test_kernels_gpu.py tests/test_dkl_gpu.py -q 2>&1 | tail -5
for wl in rvae imspec; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-baselines > gpurun_out/r2x_bench_$wl.json 2> gpurun_out/r2x_bench_$wl.err
  tail -c 300 gpurun_out/r2x_bench_$wl.err; head -c 330 gpurun_out/r2x_bench_$wl.json; echo
done
