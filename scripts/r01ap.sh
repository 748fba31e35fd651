set -x
timeout 300 python -m pytest tests/test_host_cli.py -m gpu -x -q -k "two_ranks" > gpurun_out/r01ap_pytest.log 2>&1
tail -n 40 gpurun_out/r01ap_pytest.log
