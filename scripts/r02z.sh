set -x
B="timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup --no-checks"
$B > gpurun_out/r02z_def.json 2> gpurun_out/r02z_def.err
KMN_LIB_VARIANT=c2048n2 $B > gpurun_out/r02z_c2048n2.json 2> gpurun_out/r02z_c2048n2.err
KMN_LIB_VARIANT=c2048n3 $B > gpurun_out/r02z_c2048n3.json 2> gpurun_out/r02z_c2048n3.err
KMN_SPLIT_S=8 $B > gpurun_out/r02z_s8.json 2> gpurun_out/r02z_s8.err
KMN_SPLIT_S=2 $B > gpurun_out/r02z_s2.json 2> gpurun_out/r02z_s2.err
for f in gpurun_out/r02z_*.err; do tail -c 3000 $f > $f.tail; rm -f $f; done
