mkdir -p gpurun_out
timeout 300 python tools/debug_wgrad_mix.py > gpurun_out/r2m_mix.log 2>&1; cat gpurun_out/r2m_mix.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -o gpurun_out/r2m_conv_x3 -f python tools/profile_conv.py x3 c6 c51 c40 bn3 > gpurun_out/r2m_ncu1.log 2>&1; tail -2 gpurun_out/r2m_ncu1.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -o gpurun_out/r2m_wgrad_x3 -f python tools/profile_conv.py x3 wgrad c6 c51 bn3 > gpurun_out/r2m_ncu2.log 2>&1; tail -2 gpurun_out/r2m_ncu2.log
