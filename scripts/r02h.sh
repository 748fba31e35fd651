set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r02h_pytest.log
B="timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup --no-checks"
$B > gpurun_out/r02h_tma.json 2> gpurun_out/r02h_tma.err
$B --pipe-batches 2 > gpurun_out/r02h_pb2.json 2> gpurun_out/r02h_pb2.err
$B --pipe-batches 3 > gpurun_out/r02h_pb3.json 2> gpurun_out/r02h_pb3.err
P="timeout 600 python bench.py --reads 25000000 --genome 62500000 --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup --no-checks"
mkdir -p /tmp/ncu
for k in k_count_slices_tma; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 9 -c 1 -f -o /tmp/ncu/$k $P > gpurun_out/r02h_ncu_$k.log 2>&1
  ncu -i /tmp/ncu/$k.ncu-rep --page raw --csv > gpurun_out/r02h_${k}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$k.ncu-rep --page source --csv > gpurun_out/r02h_${k}_src.csv 2>/dev/null
done
