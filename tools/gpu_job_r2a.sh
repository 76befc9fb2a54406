set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
( time python -m pytest tests -q -m gpu -x 2>&1 | tail -15 ) > gpurun_out/r2a_pytest.log 2>&1
tail -20 gpurun_out/r2a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/r2a_smoke.log
