set -x
KMN_SMEM_COUNT_W2=1 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02ah_pytest_w2.log
B="timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --workload c5a"
KMN_SMEM_COUNT_W2=1 $B > gpurun_out/r02ah_c5a_w2.json 2> gpurun_out/r02ah_c5a_w2.err
for f in gpurun_out/r02ah_*.err; do tail -c 3000 $f > $f.tail; rm -f $f; done
