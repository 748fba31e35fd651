set -x
n=4
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r01ad_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r01ad_n$n.json 2> gpurun_out/r01ad_n$n.err
KMN_ROUND_SPLIT=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r01ad_n${n}_split1.json 2> gpurun_out/r01ad_n${n}_split1.err
