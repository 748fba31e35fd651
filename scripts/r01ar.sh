python -m pytest tests -m gpu -x -q 2>&1 | tail -n 6 > gpurun_out/r01ar_pytest.log
