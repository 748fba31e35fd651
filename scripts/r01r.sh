set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r01r_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
$B --pipe-batches 4 > gpurun_out/r01r_p4.json 2> gpurun_out/r01r_p4.err
KMN_TILES=1 $B --pipe-batches 4 > gpurun_out/r01r_p4t.json 2> gpurun_out/r01r_p4t.err
$B --pipe-batches 3 > gpurun_out/r01r_p3.json 2> gpurun_out/r01r_p3.err
KMN_TILES=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r01r_pytest_tiles.log
