set -x
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/r01ab_gpus.txt
free -g > gpurun_out/r01ab_mem.txt
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r01ab_pytest.log
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 2 --warmup 1 --no-cpu > gpurun_out/r01ab_n$n.json 2> gpurun_out/r01ab_n$n.err
done
