import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from mac_b200 import synth
from mac_b200.solvers import MAC
fixed, cand, n = synth.chain_plus_random(3000, 15000, seed=5, weighted=True)
mac = MAC(fixed, cand, n, fiedler_max_steps=40)
import warnings
warnings.simplefilter("ignore")
try:
    lam, v = mac.fiedler_pair(synth.first_k_init(15000, 3000))
    print("pipe", mac._h.lanczos_kernel_name(), lam, mac.last_info)
except Exception as e:
    print("solve raised", e)
mac.close()
