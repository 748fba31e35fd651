set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -10 > gpurun_out/r02t_pytest.log
timeout 900 python bench.py > gpurun_out/r02t_full.json 2> gpurun_out/r02t_full.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv > gpurun_out/r02t_mem.txt
B="timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-lookup --no-checks"
KMN_SPLIT_BATCHES=1 $B > gpurun_out/r02t_b1.json 2> gpurun_out/r02t_b1.err
KMN_SPLIT_BATCHES=8 $B > gpurun_out/r02t_b8.json 2> gpurun_out/r02t_b8.err
KMN_SCATTER_STEPS=2 $B > gpurun_out/r02t_st2.json 2> gpurun_out/r02t_st2.err
for f in gpurun_out/r02t_*.err; do tail -c 4000 $f > $f.tail; rm -f $f; done
