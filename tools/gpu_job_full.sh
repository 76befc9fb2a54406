set -x
( time timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/full_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/full_smoke.log
for a in "2048 128 9 1024 64 64" "512 35 650 650 300 300" "8192 24 77 256 32 32"; do
timeout 300 python tools/time_r2.py $a 3 2>&1 | tail -3 | tee -a gpurun_out/r2b_time.log
done
