set -x
timeout 500 compute-sanitizer --tool memcheck python -m pytest "tests/test_gpu_parity.py::test_weight_bound_mixed_qualities" "tests/test_gpu_parity.py::test_hot_kmers_tiny_genome" -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r02aj_memcheck.log
