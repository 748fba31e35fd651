set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r01g_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --pipe-batches 2"
$B > gpurun_out/r01g_a.json 2> gpurun_out/r01g_a.err
KMN_INSERT_PRE=0 $B > gpurun_out/r01g_b.json 2> gpurun_out/r01g_b.err
KMN_INSERT_PRE=0 KMN_INSERT_CTAS=6 $B > gpurun_out/r01g_c.json 2> gpurun_out/r01g_c.err
KMN_INSERT_CTAS=4 $B > gpurun_out/r01g_d.json 2> gpurun_out/r01g_d.err
ncu --set full --clock-control none --import-source on -k regex:"k_kmer_scatter" -c 1 -f -o gpurun_out/r01g_scatter_full python bench.py --reads 20000000 --genome 50000000 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r01g_ncu.log 2>&1
