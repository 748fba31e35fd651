set -x
timeout 400 python -m pytest tests/test_multigpu.py tests/test_host_cli.py tests/test_gpu_parity.py -m gpu -x -q -k "owner_sharded or two_ranks or many_small or count_table_synthetic or large_scale" > gpurun_out/r01aq_pytest.log 2>&1
tail -n 30 gpurun_out/r01aq_pytest.log
