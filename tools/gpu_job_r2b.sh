set -x
( timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/r2b_pytest.log
for a in "2048 128 9 1024 64 64" "8192 24 77 256 32 32"; do
timeout 300 python tools/time_r2.py $a 3 2>&1 | tail -3 | tee -a gpurun_out/r2b_time.log
done
bash tools/gpu_job_prof_r02.sh
