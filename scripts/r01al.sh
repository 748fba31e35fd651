set -x
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -n 4 > gpurun_out/r01al_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu > gpurun_out/r01al_n2.json 2> gpurun_out/r01al_n2.err
