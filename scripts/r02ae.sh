set -x
timeout 300 python -m pytest tests/test_multigpu.py -m gpu -q -x -k "2-env0" 2>&1 | tail -6 > gpurun_out/r02ae_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu --no-e2e --no-checks > gpurun_out/r02ae_2gpu.json 2> gpurun_out/r02ae_2gpu.err
tail -c 3000 gpurun_out/r02ae_2gpu.err > gpurun_out/r02ae_2gpu.err.tail; rm -f gpurun_out/r02ae_2gpu.err
