set -x
B="timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup --no-checks"
KMN_SPLIT_TPB=512 $B > gpurun_out/r02af_split512.json 2> gpurun_out/r02af_split512.err
KMN_COUNT_WS=3 $B > gpurun_out/r02af_ws3.json 2> gpurun_out/r02af_ws3.err
KMN_SPLIT_TPB=512 KMN_COUNT_WS=3 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02af_pytest.log
for f in gpurun_out/r02af_*.err; do tail -c 2000 $f > $f.tail; rm -f $f; done
