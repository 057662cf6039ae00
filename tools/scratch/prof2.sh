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
PROF_ITERS=10 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2b_launches.csv python /tmp/prof_h.py > gpurun_out/r2b_prof_launch.log 2>&1
tail -2 gpurun_out/r2b_prof_launch.log
PROF_ITERS=10 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lanczos_pipe -s 9 -c 1 -o gpurun_out/r2b_pipe python /tmp/prof_h.py > gpurun_out/r2b_prof_full.log 2>&1
tail -3 gpurun_out/r2b_prof_full.log
PROF_ITERS=4 timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_sel2_hist|k_sel2_refine|k_ritz|k_spmv|k_sel2_compact' -s 10 -c 6 -o gpurun_out/r2b_small python /tmp/prof_h.py > gpurun_out/r2b_prof_small.log 2>&1
tail -3 gpurun_out/r2b_prof_small.log
ls -la gpurun_out/*.ncu-rep
