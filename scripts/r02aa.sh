set -x
timeout 500 python -m pytest tests/test_multigpu.py -m gpu -q -x -k "2-env0 or 2-env1" 2>&1 | tail -8 > gpurun_out/r02aa_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup > gpurun_out/r02aa_2gpu.json 2> gpurun_out/r02aa_2gpu.err
tail -c 5000 gpurun_out/r02aa_2gpu.err > gpurun_out/r02aa_2gpu.err.tail; rm -f gpurun_out/r02aa_2gpu.err
