set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02ag_pytest.log
timeout 600 python scripts/fastq_e2e.py 2000000 > gpurun_out/r02ag_fastq_e2e.json 2> gpurun_out/r02ag_fastq_e2e.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ag_smoke.log 2>&1
