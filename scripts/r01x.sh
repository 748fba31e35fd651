set -x
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r01x_pytest_ce.log
KMN_PUSH=kernel timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r01x_pytest_kernel.log
TR="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu --no-e2e"
$TR > gpurun_out/r01x_ce.json 2> gpurun_out/r01x_ce.err
KMN_PUSH=kernel $TR > gpurun_out/r01x_kernel.json 2> gpurun_out/r01x_kernel.err
