set -x
B="timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup --no-checks"
$B --slice-mb 128 > gpurun_out/r02ac_s128.json 2> gpurun_out/r02ac_s128.err
$B --slice-mb 32 > gpurun_out/r02ac_s32.json 2> gpurun_out/r02ac_s32.err
KMN_SCATTER_STEPS=3 $B > gpurun_out/r02ac_st3.json 2> gpurun_out/r02ac_st3.err
for f in gpurun_out/r02ac_*.err; do tail -c 2000 $f > $f.tail; rm -f $f; done
