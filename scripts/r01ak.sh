set -x
n=8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/r01ak_n$n.json 2> gpurun_out/r01ak_n$n.err
