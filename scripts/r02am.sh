set -x
timeout 600 python -m pytest tests/test_host_cli.py -m gpu -q 2>&1 | tail -6 > gpurun_out/r02am_pytest.log
