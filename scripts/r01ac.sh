set -x
n=4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r01ac_n$n.json 2> gpurun_out/r01ac_n$n.err
