set -x
cat > /tmp/prof_h.py <<'PY'
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from mac_b200 import synth
from mac_b200.solvers import MAC
fixed, cand, n, k, x0 = synth.headline()
mac = MAC(fixed, cand, n)
iters = int(os.environ.get("PROF_ITERS", "2"))
w, u, info = mac.frank_wolfe(k, x0, iters, 0.0, 0.0)
print("done", info["f_hist"], mac._h.counters(), mac._h.device_rr_stats()["fallbacks"])
PY
PROF_ITERS=10 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_pipe.csv python /tmp/prof_h.py > gpurun_out/r2_prof_launch.log 2>&1
tail -2 gpurun_out/r2_prof_launch.log
PROF_ITERS=10 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lanczos_pipe -s 9 -c 1 -o gpurun_out/r2_pipe python /tmp/prof_h.py > gpurun_out/r2_prof_full.log 2>&1
tail -3 gpurun_out/r2_prof_full.log
ls -la gpurun_out/r2_pipe.ncu-rep
