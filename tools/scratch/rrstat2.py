import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from mac_b200 import synth
from mac_b200.solvers import MAC
fixed, cand, n, k, x0 = synth.headline()
mac = MAC(fixed, cand, n)
for x in (x0, np.full(len(x0), 0.2)):
    lam, v = mac.fiedler_pair(x)
    st = mac._h.device_rr_stats()
    print(st["last_k"], st["last_checks"], st["lag_steps_at_decision"], "wait", st["cycles_waiting"], "compute", st["cycles_computing"], st["cycles_by_stage"], mac.last_info["steps"])
