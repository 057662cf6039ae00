import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from mac_b200 import synth
from mac_b200.solvers import MAC
for tag, n, m, k in (("small", 600, 3000, 600), ("multi-CTA", 6000, 40000, 8000)):
    fixed, cand, n = synth.chain_plus_random(n, m, seed=5, weighted=True)
    mac = MAC(fixed, cand, n)
    r, w, u = mac.solve(k, synth.first_k_init(m, k), max_iters=2)
    print(tag, mac._h.lanczos_kernel_name(), u, int(r.sum()))
    mac.close()
