mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
for p2p in 0 1; do
  ATOMAI_B200_P2P=$p2p timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tools/dp_equivalence.py > gpurun_out/r2g_dp_p2p$p2p.log 2>&1
  grep -E "^\{|Error|error:" gpurun_out/r2g_dp_p2p$p2p.log | tail -4
done
for p2p in 1 0; do
  ATOMAI_B200_P2P=$p2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 10 --warmup 3 --math tf32x3 > gpurun_out/r2g_bench2_p2p$p2p.json 2> gpurun_out/r2g_bench2_p2p$p2p.err
  grep -E "Error|error:|atomai_b200:" gpurun_out/r2g_bench2_p2p$p2p.err | tail -3; head -c 300 gpurun_out/r2g_bench2_p2p$p2p.json; echo
done
ATOMAI_B200_P2P=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 10 --warmup 3 --math tf32x3 --scaling strong > gpurun_out/r2g_bench2_strong.json 2> gpurun_out/r2g_bench2_strong.err
head -c 300 gpurun_out/r2g_bench2_strong.json; echo
