set -x
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02g_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup --no-checks"
$B > gpurun_out/r02g_tma.json 2> gpurun_out/r02g_tma.err
KMN_COUNT_TMA=0 $B > gpurun_out/r02g_notma.json 2> gpurun_out/r02g_notma.err
P="python bench.py --reads 25000000 --genome 62500000 --steps 1 --warmup 1 --no-cpu --no-e2e --no-lookup --no-checks"
mkdir -p /tmp/ncu
for k in k_count_slices_tma k_slice_split k_kmer_scatter; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o /tmp/ncu/$k $P > gpurun_out/r02g_ncu_$k.log 2>&1
  ncu -i /tmp/ncu/$k.ncu-rep --page raw --csv > gpurun_out/r02g_${k}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$k.ncu-rep --page source --csv > gpurun_out/r02g_${k}_src.csv 2>/dev/null
done
KMN_COUNT_TMA=0 ncu --set full --clock-control none --import-source on -k regex:k_count_slices -s 4 -c 1 -f -o /tmp/ncu/k_count_slices_old $P > gpurun_out/r02g_ncu_k_count_old.log 2>&1
ncu -i /tmp/ncu/k_count_slices_old.ncu-rep --page raw --csv > gpurun_out/r02g_k_count_old_raw.csv 2>/dev/null
ncu -i /tmp/ncu/k_count_slices_old.ncu-rep --page source --csv > gpurun_out/r02g_k_count_old_src.csv 2>/dev/null
ls -la gpurun_out/ | tail -20 > gpurun_out/r02g_ls.txt
