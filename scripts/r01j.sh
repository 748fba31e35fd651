python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r01j_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
$B > gpurun_out/r01j_a.json 2> gpurun_out/r01j_a.err
KMN_NO_L2_HINTS=1 $B > gpurun_out/r01j_b.json 2> gpurun_out/r01j_b.err
$B --pipe-batches 2 > gpurun_out/r01j_c.json 2> gpurun_out/r01j_c.err
$B --pipe-batches 8 > gpurun_out/r01j_d.json 2> gpurun_out/r01j_d.err
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed
ncu --metrics $M --clock-control none -k regex:"k_kmer_scatter|k_insert_staged|k_weight_mask" -c 3 --csv --log-file gpurun_out/r01j_ncu100.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r01j_ncu100.log 2>&1
