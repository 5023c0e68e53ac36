mkdir -p gpurun_out
timeout 200 python tools/debug_determinism.py fp32 2>&1 | grep -E "max rel|Error" 
ATOMAI_B200_NO_POOLSTATS=1 timeout 200 python tools/debug_determinism.py fp32 2>&1 | grep -E "max rel|Error"
timeout 200 python -m pytest tests/test_segmentor_gpu.py -q 2>&1 | tail -2
