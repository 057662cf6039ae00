import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mac_b200.solvers import MAC, NaiveGreedy
from mac_b200.g2o import split_edges
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for name, k in (("intel", 157), ("sphere2500", 1225), ("city10000", 1068)):
    z = np.load(os.path.join(G, f"g2o_{name}.npz")); fixed, cand = split_edges(z["i"], z["j"], z["kappa"]); n = int(z["n"])
    mac = MAC(fixed, cand, n); x0 = NaiveGreedy(cand[2]).subset(k)
    mac.fiedler_pair(x0)
    t0 = time.perf_counter(); mac._h.set_x(x0); t1 = time.perf_counter(); lam, v, info = mac._h.fiedler(); t2 = time.perf_counter()
    print(name, "set_x %.0f us, fiedler %.0f us" % ((t1 - t0) * 1e6, (t2 - t1) * 1e6), info, flush=True)
