set -x
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r02o_pytest.log
KMN_COUNT_WS=1 KMN_SCATTER_STEPS=2 timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -20 > gpurun_out/r02o_pytest_ws.log
B="timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup --no-checks"
$B > gpurun_out/r02o_def.json 2> gpurun_out/r02o_def.err
KMN_COUNT_WS=1 $B > gpurun_out/r02o_ws.json 2> gpurun_out/r02o_ws.err
KMN_SCATTER_STEPS=2 $B > gpurun_out/r02o_st2.json 2> gpurun_out/r02o_st2.err
$B --pipe-batches 3 > gpurun_out/r02o_pb3.json 2> gpurun_out/r02o_pb3.err
KMN_COUNT_WS=1 KMN_SCATTER_STEPS=2 $B --pipe-batches 2 > gpurun_out/r02o_all_pb2.json 2> gpurun_out/r02o_all_pb2.err
for f in gpurun_out/r02o_*.err; do tail -c 4000 $f > $f.tail; rm -f $f; done
