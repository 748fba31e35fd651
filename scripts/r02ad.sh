set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup > gpurun_out/r02ad_n8.json 2> gpurun_out/r02ad_n8.err
tail -c 3000 gpurun_out/r02ad_n8.err > gpurun_out/r02ad_n8.err.tail; rm -f gpurun_out/r02ad_n8.err
