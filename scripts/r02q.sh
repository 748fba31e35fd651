set -x
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r02q_pytest.log
B="timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup --no-checks"
$B > gpurun_out/r02q_def.json 2> gpurun_out/r02q_def.err
KMN_COUNT_DB=0 $B > gpurun_out/r02q_nodb.json 2> gpurun_out/r02q_nodb.err
$B --pipe-batches 3 > gpurun_out/r02q_pb3.json 2> gpurun_out/r02q_pb3.err
KMN_SCATTER_STEPS=2 $B --pipe-batches 2 > gpurun_out/r02q_pb2.json 2> gpurun_out/r02q_pb2.err
timeout 600 compute-sanitizer --tool memcheck python -m pytest "tests/test_gpu_parity.py::test_weight_bound_mixed_qualities" "tests/test_gpu_parity.py::test_hot_kmers_tiny_genome" -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02q_memcheck.log
for f in gpurun_out/r02q_*.err; do tail -c 4000 $f > $f.tail; rm -f $f; done
