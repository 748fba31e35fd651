set -x
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02d_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup"
$B > gpurun_out/r02d_def.json 2> gpurun_out/r02d_def.err
KMN_SPLIT_S=2 $B --no-checks > gpurun_out/r02d_s2.json 2> gpurun_out/r02d_s2.err
KMN_SPLIT_S=8 $B --no-checks > gpurun_out/r02d_s8.json 2> gpurun_out/r02d_s8.err
$B --no-checks --pipe-batches 2 > gpurun_out/r02d_pb2.json 2> gpurun_out/r02d_pb2.err
$B --no-checks --pipe-batches 3 > gpurun_out/r02d_pb3.json 2> gpurun_out/r02d_pb3.err
KMN_SPLIT_TPB=512 KMN_SPLIT_CTAS=2 $B --no-checks > gpurun_out/r02d_t512.json 2> gpurun_out/r02d_t512.err
