set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r02ak_pytest.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest "tests/test_gpu_parity.py::test_weight_bound_mixed_qualities" -m gpu -q -x -k "32 or 47" 2>&1 | tail -6 > gpurun_out/r02ak_memcheck.log
