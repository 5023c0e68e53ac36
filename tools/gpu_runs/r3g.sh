mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_multigpu.py -q -s 2>&1 | grep -E "^\{|passed|failed|assert" | cut -c1-700
