set -x
B="timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup --no-checks"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/r02final_launches.csv $B > gpurun_out/r02final_launches.log 2>&1
