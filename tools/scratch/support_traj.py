"""Support (active off-diagonal slots) of L(x_t) along the headline FW trajectory + Lanczos steps per solve."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mac_b200 import synth
from mac_b200.solvers import MAC
fixed, cand, n, k, x = synth.headline()
mac = MAC(fixed, cand, n)
for i in range(50):
    f, g = mac.problem(x)
    s = mac.solve_lp(k)
    print(i, "nnz_active", mac._h.sizes()["nnz_active"], "steps", mac.last_info["steps"], "f %.6f" % f, flush=True)
    x = x + 2.0 / (i + 2.0) * (s - x)
