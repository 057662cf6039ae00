"""ncu target: one cold Fiedler solve + gradient + top-k on the headline graph (scratch tool)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mac_b200 import synth
from mac_b200.solvers import MAC
n = int(os.environ.get("PROF_N", 100000)); m = int(os.environ.get("PROF_M", 1000000))
fixed, cand, n, k, x0 = synth.headline(n=n, m=m)
mac = MAC(fixed, cand, n)
iters = int(os.environ.get("PROF_ITERS", 1))
w, u, info = mac.frank_wolfe(k, x0, iters, 0.0, 0.0)
print(info, mac._h.counters())
