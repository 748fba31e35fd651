set -x
B="timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup --no-checks"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02final_launches.csv $B > gpurun_out/r02final_launches.log 2>&1
mkdir -p /tmp/ncu
for spec in k_kmer_scatter:2 k_slice_split:4 k_count_slices_ws:4 k_weight_mask:2; do
  k=${spec%%:*}; sk=${spec##*:}
  ncu --set full --clock-control none --import-source on -k regex:$k -s $sk -c 1 -f -o /tmp/ncu/$k $B > gpurun_out/r02final_ncu_$k.log 2>&1
  ncu -i /tmp/ncu/$k.ncu-rep --page raw --csv > gpurun_out/r02final_${k}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$k.ncu-rep --page source --csv > gpurun_out/r02final_${k}_src.csv 2>/dev/null
done
timeout 900 python bench.py > gpurun_out/r02final_bench.json 2> gpurun_out/r02final_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r02final_ref.json 2> gpurun_out/r02final_ref.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02final_smoke.log 2>&1
for f in gpurun_out/r02final_*.err; do tail -c 3000 $f > $f.tail; rm -f $f; done
