set -x
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02e_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup --no-checks"
$B > gpurun_out/r02e_def.json 2> gpurun_out/r02e_def.err
KMN_SPLIT_TPB=512 KMN_SPLIT_CTAS=2 $B > gpurun_out/r02e_t512.json 2> gpurun_out/r02e_t512.err
KMN_SPLIT_S=2 $B > gpurun_out/r02e_s2.json 2> gpurun_out/r02e_s2.err
$B --pipe-batches 2 > gpurun_out/r02e_pb2.json 2> gpurun_out/r02e_pb2.err
KMN_SMEM_COUNT=0 $B > gpurun_out/r02e_nosmem.json 2> gpurun_out/r02e_nosmem.err
python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --workload c5b > gpurun_out/r02e_c5b.json 2> gpurun_out/r02e_c5b.err
python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --workload c5a > gpurun_out/r02e_c5a.json 2> gpurun_out/r02e_c5a.err
python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --workload c4 > gpurun_out/r02e_c4.json 2> gpurun_out/r02e_c4.err
