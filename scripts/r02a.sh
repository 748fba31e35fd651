set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt
./bench/micro/bin/lsu > gpurun_out/r02a_lsu.csv 2> gpurun_out/r02a_lsu.err
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup --no-checks"
for mb in 64 32 16 8; do
  KMN_PIPELINE=0 $B --slice-mb $mb > gpurun_out/r02a_s${mb}_serial.json 2> gpurun_out/r02a_s${mb}_serial.err
done
$B --slice-mb 16 > gpurun_out/r02a_s16_pipe.json 2> gpurun_out/r02a_s16_pipe.err
