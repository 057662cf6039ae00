import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from mac_b200 import synth
from mac_b200.solvers import MAC
fixed, cand, n, k, x0 = synth.headline()
mac = MAC(fixed, cand, n)
h = mac._h
w, u, info = mac.frank_wolfe(k, x0, 5, 0.0, 0.0)
for flush in (True, False):
    h.set_bench(True, flush); h.reset_counters()
    w, u, info = mac.frank_wolfe(k, x0, 20, 0.0, 0.0)
    print("flush", flush, "iter ms", np.round(h.iter_ms(), 2), "sum %.1f" % h.iter_ms().sum())
    lk = h.lanczos_kernel_time(); print("   kernel ms %.1f phases %d -> %.2f us/step" % (lk["ms"], lk["phases"], lk["ms"] * 1e3 / lk["phases"]))
for x in (x0, np.full(len(x0), 0.2)):
    h.set_bench(True, False); h.reset_counters()
    lam, v = mac.fiedler_pair(x)
    lk = h.lanczos_kernel_time(); print("single solve: kernel ms %.2f phases %d -> %.2f us/step" % (lk["ms"], lk["phases"], lk["ms"] * 1e3 / max(lk["phases"], 1)), mac.last_info["steps"])
