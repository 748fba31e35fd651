set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r01aa_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup"
for p in 8 4 3 2; do
KMN_PIPELINE=0 $B --pipe-batches $p > gpurun_out/r01aa_serial_p$p.json 2> gpurun_out/r01aa_serial_p$p.err
done
$B --pipe-batches 4 > gpurun_out/r01aa_pipe_p4.json 2> gpurun_out/r01aa_pipe_p4.err
