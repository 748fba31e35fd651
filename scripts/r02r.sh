set -x
P="timeout 600 python bench.py --reads 25000000 --genome 62500000 --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup --no-checks"
mkdir -p /tmp/ncu
for k in k_count_slices_db; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 9 -c 1 -f -o /tmp/ncu/$k $P > gpurun_out/r02r_ncu_$k.log 2>&1
  ncu -i /tmp/ncu/$k.ncu-rep --page raw --csv > gpurun_out/r02r_${k}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$k.ncu-rep --page source --csv > gpurun_out/r02r_${k}_src.csv 2>/dev/null
done
