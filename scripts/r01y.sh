set -x
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
KMN_PIPELINE=0 $B > gpurun_out/r01y_s64_serial.json 2> gpurun_out/r01y_s64_serial.err
KMN_PIPELINE=0 $B --slice-mb 128 --no-lookup > gpurun_out/r01y_s128_serial.json 2> gpurun_out/r01y_s128_serial.err
KMN_PIPELINE=0 $B --slice-mb 256 --no-lookup > gpurun_out/r01y_s256_serial.json 2> gpurun_out/r01y_s256_serial.err
$B --slice-mb 128 --no-lookup > gpurun_out/r01y_s128_pipe.json 2> gpurun_out/r01y_s128_pipe.err
KMN_PIPELINE=0 ncu --set full --clock-control none --import-source on -k regex:k_kmer_scatter -s 2 -c 1 -o gpurun_out/r01y_scatter python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-lookup > gpurun_out/r01y_ncu.log 2>&1
