set -x
KMN_COUNT_WS=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -10 > gpurun_out/r02s_pytest_ws2.log
B="timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup --no-checks"
KMN_COUNT_WS=2 $B > gpurun_out/r02s_ws2.json 2> gpurun_out/r02s_ws2.err
KMN_COUNT_WS=2 $B --pipe-batches 2 > gpurun_out/r02s_ws2_pb2.json 2> gpurun_out/r02s_ws2_pb2.err
P="timeout 600 python bench.py --reads 25000000 --genome 62500000 --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup --no-checks"
mkdir -p /tmp/ncu
k=k_count_slices_ws
KMN_COUNT_WS=2 ncu --set full --clock-control none --import-source on -k regex:$k -s 9 -c 1 -f -o /tmp/ncu/$k $P > gpurun_out/r02s_ncu_$k.log 2>&1
ncu -i /tmp/ncu/$k.ncu-rep --page raw --csv > gpurun_out/r02s_${k}_raw.csv 2>/dev/null
ncu -i /tmp/ncu/$k.ncu-rep --page source --csv > gpurun_out/r02s_${k}_src.csv 2>/dev/null
for f in gpurun_out/r02s_*.err; do tail -c 4000 $f > $f.tail; rm -f $f; done
