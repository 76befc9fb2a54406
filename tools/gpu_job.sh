python -m pytest tests -q -m gpu -x 2>&1 | tail -4
python tools/time_fwd.py 2>&1 | grep -E "clocks under|^train"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"seq_|grad_rows|reduce_partials|xproj" -s 8 -c 5 --csv --log-file gpurun_out/launches_mma.csv python tools/prof_step.py 8192 3 > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_mma.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    print(r['Kernel Name'][:70], r['Metric Value'])
"
build/test_gemm_tc | grep -E "rel|ms per|OK|FAIL"
