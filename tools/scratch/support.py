import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from mac_b200 import synth
from mac_b200.solvers import MAC
fixed, cand, n, k, x0 = synth.headline()
mac = MAC(fixed, cand, n)
for it in (1, 2, 5, 10, 20, 40):
    w, u, info = mac.frank_wolfe(k, x0, max_iters=it)
    print(it, "support", int((w > 1e-10).sum()), "of", len(w), "nnz_active", mac._h.sizes()["nnz_active"], flush=True)
