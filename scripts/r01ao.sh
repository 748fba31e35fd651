set -x
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/r01ao_pytest.log 2>&1
tail -n 30 gpurun_out/r01ao_pytest.log
