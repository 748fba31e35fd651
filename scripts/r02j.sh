set -x
nvidia-smi -L > gpurun_out/r02j_gpus.txt
timeout 1800 python -m pytest tests/test_multigpu.py tests/test_host_cli.py "tests/test_gpu_parity.py::test_count_batch_2na_equals_ascii" -m gpu -q 2>&1 | tail -40 > gpurun_out/r02j_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02j_2gpu.json 2> gpurun_out/r02j_2gpu.err
tail -c 3000 gpurun_out/r02j_2gpu.err > gpurun_out/r02j_2gpu.err.tail; rm -f gpurun_out/r02j_2gpu.err
