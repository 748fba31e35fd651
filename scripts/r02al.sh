set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02al_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02al_smoke.log 2>&1
