mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
timeout 200 python -m pytest tests/test_multigpu.py -q 2>&1 | tail -4
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus 4 --workload imspec --steps 10 --warmup 3 --no-baselines > gpurun_out/r3b_imspec4.json 2> gpurun_out/r3b_imspec4.err
grep -E "Error|error:" gpurun_out/r3b_imspec4.err | tail -3; head -c 500 gpurun_out/r3b_imspec4.json; echo
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29752 bench.py --gpus 4 --workload seg256 --steps 8 --warmup 3 --math tf32x3 --no-baselines > gpurun_out/r3b_seg256_4.json 2> gpurun_out/r3b_seg256_4.err
grep -E "Error|error:" gpurun_out/r3b_seg256_4.err | tail -3; head -c 400 gpurun_out/r3b_seg256_4.json; echo
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29753 bench.py --gpus 4 --workload gram --steps 5 --warmup 2 --no-baselines > gpurun_out/r3b_gram4.json 2> gpurun_out/r3b_gram4.err
grep -E "Error|error:" gpurun_out/r3b_gram4.err | tail -3; head -c 400 gpurun_out/r3b_gram4.json; echo
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29754 bench.py --gpus 4 --steps 6 --warmup 3 --math tf32x3 --no-baselines > gpurun_out/r3b_seg512_4.json 2> gpurun_out/r3b_seg512_4.err
head -c 300 gpurun_out/r3b_seg512_4.json; echo
