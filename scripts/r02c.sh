set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02c_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup"
$B > gpurun_out/r02c_def.json 2> gpurun_out/r02c_def.err
KMN_SMEM_COUNT=0 $B --no-checks > gpurun_out/r02c_nosmem.json 2> gpurun_out/r02c_nosmem.err
KMN_SPLIT_S=8 $B --no-checks > gpurun_out/r02c_s8.json 2> gpurun_out/r02c_s8.err
KMN_SPLIT_S=2 $B --no-checks > gpurun_out/r02c_s2.json 2> gpurun_out/r02c_s2.err
$B --no-checks --pipe-batches 2 > gpurun_out/r02c_pb2.json 2> gpurun_out/r02c_pb2.err
$B --no-checks --pipe-batches 8 > gpurun_out/r02c_pb8.json 2> gpurun_out/r02c_pb8.err
KMN_SPLIT_TPB=512 KMN_SPLIT_CTAS=2 $B --no-checks > gpurun_out/r02c_t512.json 2> gpurun_out/r02c_t512.err
