set -x
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/r01ae_pytest_s4.log 2>&1
KMN_ROUND_SPLIT=1 timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/r01ae_pytest_s1.log 2>&1
