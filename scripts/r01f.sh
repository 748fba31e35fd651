set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "synthetic or saturation" 2>&1 | tail -5 > gpurun_out/r01f_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
for pb in 1 2 4 8; do
$B --pipe-batches $pb > gpurun_out/r01f_c$pb.json 2> gpurun_out/r01f_c$pb.err
done
KMN_PART_MAJOR=1 $B --pipe-batches 4 > gpurun_out/r01f_p4.json 2> gpurun_out/r01f_p4.err
KMN_PART_MAJOR=1 $B --pipe-batches 8 > gpurun_out/r01f_p8.json 2> gpurun_out/r01f_p8.err
