import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
from mac_b200.g2o import split_edges
from mac_b200.solvers import MAC
name = sys.argv[1]
z = np.load(f"tests/golden/g2o_{name}.npz")
fixed, cand = split_edges(z["i"], z["j"], z["kappa"]); n = int(z["n"]); m = len(cand[0])
mac = MAC(fixed, cand, n)
k = int(0.4 * m); x0 = np.zeros(m); x0[:k] = 1.0
for rep in range(3):
    mac._h.reset_counters()
    t = time.perf_counter()
    r, w, u = mac.solve(k, x0, max_iters=20)
    dt = time.perf_counter() - t
c = mac._h.counters()
print(name, "RR smem", os.environ.get("MACB_RR_SMEM_KB"), "%.1f ms" % (dt * 1e3), "steps", c["lanczos_steps"], mac._h.lanczos_footprint())
