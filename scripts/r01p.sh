set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r01p_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
$B --pipe-batches 8 > gpurun_out/r01p_p8.json 2> gpurun_out/r01p_p8.err
$B --pipe-batches 4 > gpurun_out/r01p_p4.json 2> gpurun_out/r01p_p4.err
$B --pipe-batches 3 > gpurun_out/r01p_p3.json 2> gpurun_out/r01p_p3.err
KMN_FAST=0 $B --pipe-batches 8 > gpurun_out/r01p_old.json 2> gpurun_out/r01p_old.err
