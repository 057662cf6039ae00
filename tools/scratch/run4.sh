timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -14
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench4.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['us_per_lanczos_step'], d['roofline']['share_of_timed_region'], d['config']['lanczos_steps_per_solve'], d.get('parity_check'))
PY
python - <<'PY'
import numpy as np, time
from mac_b200 import synth
from mac_b200.solvers import MAC
fixed, cand, n, k, x0 = synth.headline()
mac = MAC(fixed, cand, n)
w,u,info = mac.frank_wolfe(k, x0, 10, 0.0, 0.0)
print("rr stats", mac._h.device_rr_stats(), mac._h.counters())
PY
