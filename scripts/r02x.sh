set -x
nvidia-smi -L > gpurun_out/r02x_gpus.txt
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02x_n8.json 2> gpurun_out/r02x_n8.err
tail -c 5000 gpurun_out/r02x_n8.err > gpurun_out/r02x_n8.err.tail; rm -f gpurun_out/r02x_n8.err
timeout 300 python -m pytest tests/test_multigpu.py -m gpu -q -x -k "8-env or 4-env0" 2>&1 | tail -8 > gpurun_out/r02x_pytest.log
