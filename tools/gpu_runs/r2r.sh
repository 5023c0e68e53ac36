mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
nvidia-smi -L | wc -l
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29741 tools/dp_equivalence.py > gpurun_out/r2r_dp8.log 2>&1
grep -E "^\{|Error|error:|atomai_b200:" gpurun_out/r2r_dp8.log | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus 8 --steps 6 --warmup 3 --math tf32x3 --no-baselines > gpurun_out/r2r_bench8.json 2> gpurun_out/r2r_bench8.err
grep -E "Error|error:|atomai_b200:" gpurun_out/r2r_bench8.err | tail -3; head -c 400 gpurun_out/r2r_bench8.json; echo
ATOMAI_B200_P2P=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29743 bench.py --gpus 8 --steps 6 --warmup 3 --math tf32x3 --no-baselines > gpurun_out/r2r_bench8_nccl.json 2> gpurun_out/r2r_bench8_nccl.err
head -c 300 gpurun_out/r2r_bench8_nccl.json; echo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29744 bench.py --gpus 8 --steps 6 --warmup 3 --math tf32x3 --no-baselines --scaling strong > gpurun_out/r2r_bench8_strong.json 2> gpurun_out/r2r_bench8_strong.err
head -c 300 gpurun_out/r2r_bench8_strong.json; echo
timeout 200 python -m pytest tests/test_multigpu.py -q 2>&1 | tail -3
