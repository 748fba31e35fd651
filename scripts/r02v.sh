set -x
timeout 400 python -m pytest tests/test_multigpu.py -m gpu -q -x -k "2-env" 2>&1 | tail -8 > gpurun_out/r02v_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu > gpurun_out/r02v_2gpu.json 2> gpurun_out/r02v_2gpu.err
tail -c 5000 gpurun_out/r02v_2gpu.err > gpurun_out/r02v_2gpu.err.tail; rm -f gpurun_out/r02v_2gpu.err
