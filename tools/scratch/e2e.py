import numpy as np, time, sys, os
sys.path.insert(0, os.getcwd())
from mac_b200 import synth
from mac_b200.solvers import MAC
fixed, cand, n, k, x0 = synth.headline()
mac = MAC(fixed, cand, n); h = mac._h
mac.frank_wolfe(k, x0, 3, 0.0, 0.0)
for K in (1, 5, 20, 40):
    h.device_sync(); t = time.perf_counter(); w, u, info = mac.frank_wolfe(k, x0, K, 0.0, 0.0); h.device_sync(); dt = time.perf_counter() - t
    print("K", K, "wall ms %.2f" % (dt * 1e3), "per iter %.3f" % (dt * 1e3 / K))
h.set_bench(True, False)
h.device_sync(); t = time.perf_counter(); w, u, info = mac.frank_wolfe(k, x0, 20, 0.0, 0.0); h.device_sync(); dt = time.perf_counter() - t
print("bench-mode (events, no flush): wall %.2f  sum iter_ms %.2f" % (dt * 1e3, h.iter_ms().sum()), h.iter_ms()[:5])
os.environ["MACB_FW_SYNC"] = "1"
h.set_bench(False, False)
h.device_sync(); t = time.perf_counter(); w, u, info = mac.frank_wolfe(k, x0, 20, 0.0, 0.0); h.device_sync(); dt = time.perf_counter() - t
print("synchronous loop: wall %.2f" % (dt * 1e3))
