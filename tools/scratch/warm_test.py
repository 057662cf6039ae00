import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mac_b200 import synth
from mac_b200.solvers import MAC
fixed, cand, n, k, x0 = synth.headline()
mac = MAC(fixed, cand, n)
for warm in (False, True):
    mac._h.reset_counters()
    t0 = time.perf_counter(); w, u, info = mac.frank_wolfe(k, x0, 50, 0.0, 0.0, use_cache=warm); dt = time.perf_counter() - t0
    c = mac._h.counters()
    print("warm", warm, "%.1f ms" % (dt * 1e3), "%.1f it/s" % (50 / dt), "steps/solve", c["lanczos_steps"] / c["fiedler_solves"], "f_last", info["f_hist"][-1], "u", u)
    if not warm: w_cold = w
print("max |w_warm - w_cold|", np.abs(w - w_cold).max())
