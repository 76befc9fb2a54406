timeout 600 python -m pytest tests -q -m gpu -x -k "generic_regime or lm or group_g4" 2>&1 | tail -3
for B in 20 512; do python tools/time_lm.py $B; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm|seq_|grad|reduce|colred|dpre|transpose|zero_pad|splitk" -s 100 -c 400 --csv --log-file gpurun_out/launches_lm.csv python tools/prof_lm.py 512 > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_lm.csv | python -c "
import csv,sys,collections,re
agg=collections.OrderedDict()
for r in csv.DictReader(sys.stdin):
    n=re.sub(r'\(.*','',r['Kernel Name']); n=re.sub(r'^void ','',n)[:70]
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=float(r['Metric Value'])
tot=sum(v[1] for v in agg.values())
for n,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]: print(f'{n:72s} x{c:3d} {t/1e3:9.1f} us {100*t/tot:5.1f}%')
"
