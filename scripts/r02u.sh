set -x
timeout 900 python bench.py > gpurun_out/r02u_full.json 2> gpurun_out/r02u_full.err
for w in c4 c5a c5b; do
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --workload $w > gpurun_out/r02u_$w.json 2> gpurun_out/r02u_$w.err
done
for f in gpurun_out/r02u_*.err; do tail -c 4000 $f > $f.tail; rm -f $f; done
