timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-ksweep --no-hbm-spmv 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
cat > /tmp/prof_h.py <<'PY'
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from mac_b200 import synth
from mac_b200.solvers import MAC
fixed, cand, n, k, x0 = synth.headline()
mac = MAC(fixed, cand, n)
w, u, info = mac.frank_wolfe(k, x0, 6, 0.0, 0.0)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2c_launches.csv python /tmp/prof_h.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r2c_launches.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[hi+2:]:
    if len(r)>mv:
        try: agg[r[kn][:50]].append(float(r[mv].replace(',','')))
        except: pass
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    print(f"{k:50s} n={len(v):4d} mean={sum(v)/len(v)/1000:9.1f} us")
PY
