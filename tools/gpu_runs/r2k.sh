mkdir -p gpurun_out
timeout 300 python tools/exp_epilogue.py > gpurun_out/r2k_epi3.log 2>&1; cat gpurun_out/r2k_epi3.log
timeout 300 python tools/bench_layers.py --math tf32x3 > gpurun_out/r2k_layers_x3.log 2>&1; tail -17 gpurun_out/r2k_layers_x3.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
