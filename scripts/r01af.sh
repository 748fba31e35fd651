set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01af_pytest.log 2>&1
KMN_PUSH=kernel timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/r01af_pytest_kernel.log 2>&1
KMN_P2P=0 timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/r01af_pytest_nccl.log 2>&1
