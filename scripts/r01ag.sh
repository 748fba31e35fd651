set -x
n=4
for p in 16 32; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$p bench.py --gpus $n --steps 2 --warmup 1 --no-cpu --no-e2e --pipe-batches $p > gpurun_out/r01ag_n${n}_p$p.json 2> gpurun_out/r01ag_n${n}_p$p.err
done
