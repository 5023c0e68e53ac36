mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -c 3 -o gpurun_out/r2v_wgrad_c6 -f python tools/profile_conv.py wgrad c6 > gpurun_out/r2v_ncu.log 2>&1; tail -2 gpurun_out/r2v_ncu.log
for wl in rvae imspec; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2v_launches_$wl.csv python bench.py --workload $wl --steps 3 --warmup 1 --no-baselines > gpurun_out/r2v_$wl.log 2>&1
python tools/summarize_launches.py gpurun_out/r2v_launches_$wl.csv gpurun_out/r2v_launches_$wl.md | head -24
done
