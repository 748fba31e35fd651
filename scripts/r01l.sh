B="python bench.py --steps 2 --warmup 1 --no-cpu"
KMN_PIPELINE=1 $B --pipe-batches 8 > gpurun_out/r01l_p8.json 2> gpurun_out/r01l_p8.err
KMN_PIPELINE=1 KMN_INSERT_CTAS=2 $B --pipe-batches 8 > gpurun_out/r01l_p8c2.json 2> gpurun_out/r01l_p8c2.err
KMN_PIPELINE=1 $B --pipe-batches 8 --e2e-reads 100000000 --e2e-batch 4000000 > gpurun_out/r01l_p8e100.json 2> gpurun_out/r01l_p8e100.err
$B --e2e-reads 100000000 --e2e-batch 4000000 > gpurun_out/r01l_e100.json 2> gpurun_out/r01l_e100.err
