timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
cat > /tmp/prof_h.py <<'PY'
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from mac_b200 import synth
from mac_b200.solvers import MAC
fixed, cand, n, k, x0 = synth.headline()
mac = MAC(fixed, cand, n)
w, u, info = mac.frank_wolfe(k, x0, 10, 0.0, 0.0)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c_launches.csv python /tmp/prof_h.py > /dev/null 2>&1
bash tools/scratch/run19.sh
