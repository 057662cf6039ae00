import numpy as np, time, sys, os
sys.path.insert(0, os.getcwd())
from mac_b200 import synth
from mac_b200.solvers import MAC
for seed in (1, 2, 3):
    fixed, cand, n, k, x0 = synth.headline(seed=seed)
    mac = MAC(fixed, cand, n)
    mac.frank_wolfe(k, x0, 3, 0.0, 0.0)
    mac._h.reset_counters()
    t = time.perf_counter(); w, u, info = mac.frank_wolfe(k, x0, 20, 0.0, 0.0); dt = time.perf_counter() - t
    st = mac._h.device_rr_stats(); c = mac._h.counters()
    print("seed", seed, "ms/iter %.3f" % (dt * 50), "fallbacks", st["fallbacks"], "steps/solve", c["lanczos_steps"] / c["fiedler_solves"], "solves", c["fiedler_solves"], "lag", st["lag_steps_at_decision"])
    mac.close()
