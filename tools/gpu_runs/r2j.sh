mkdir -p gpurun_out
ATOMAI_B200_TMA=1 timeout 300 python tools/debug_tma.py > gpurun_out/r2j_tma.log 2>&1; cat gpurun_out/r2j_tma.log | tail -20
ATOMAI_B200_TMA=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -o gpurun_out/r2j_c51 python tools/profile_conv.py c51 > gpurun_out/r2j_ncu.log 2>&1
tail -3 gpurun_out/r2j_ncu.log
