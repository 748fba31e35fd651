B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --pipe-batches 4"
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,lts__t_sectors.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_op_write.sum,lts__t_sectors_srcunit_ltcfabric.sum,dram__sectors_write.sum,dram__sectors_read.sum,lts__t_sectors_srcnode_gpc_op_write.sum
ncu --metrics $M --clock-control none -k regex:"k_kmer_scatter|k_insert_staged" -c 4 --csv --log-file gpurun_out/r01i_ncu100.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --pipe-batches 4 > gpurun_out/r01i_ncu100.log 2>&1
$B --slice-mb 64 > gpurun_out/r01i_s64.json 2> gpurun_out/r01i_s64.err
$B --slice-mb 96 > gpurun_out/r01i_s96.json 2> gpurun_out/r01i_s96.err
cd kmernator_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DKMN_ONLY_W1 -DKMN_SCATTER_TPB=1024 -DKMN_SCATTER_CTAS=1 -o ../libkmernator_b200.so kmn_api.cu && cd ../..
$B > gpurun_out/r01i_t1024.json 2> gpurun_out/r01i_t1024.err
$B --slice-mb 64 > gpurun_out/r01i_t1024s64.json 2> gpurun_out/r01i_t1024s64.err
