B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
for pb in 3 4 6 8 12; do
$B --pipe-batches $pb > gpurun_out/r01h_c$pb.json 2> gpurun_out/r01h_c$pb.err
done
$B --pipe-batches 4 --slice-mb 48 > gpurun_out/r01h_s48.json 2> gpurun_out/r01h_s48.err
$B --pipe-batches 4 --slice-mb 24 > gpurun_out/r01h_s24.json 2> gpurun_out/r01h_s24.err
