set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r01q_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
$B --pipe-batches 8 > gpurun_out/r01q_p8.json 2> gpurun_out/r01q_p8.err
$B --pipe-batches 4 > gpurun_out/r01q_p4.json 2> gpurun_out/r01q_p4.err
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
ncu --metrics $M --clock-control none -k regex:"k_kmer_tiles|k_subpartition|k_count_slices|k_weight_mask" -c 4 --csv --log-file gpurun_out/r01q_ncu.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --pipe-batches 4 > gpurun_out/r01q_ncu.log 2>&1
