set -x
nvidia-smi -L > gpurun_out/r02p_gpus.txt
timeout 1500 python -m pytest tests/test_multigpu.py tests/test_host_cli.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r02p_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu > gpurun_out/r02p_2gpu.json 2> gpurun_out/r02p_2gpu.err
tail -c 5000 gpurun_out/r02p_2gpu.err > gpurun_out/r02p_2gpu.err.tail; rm -f gpurun_out/r02p_2gpu.err
