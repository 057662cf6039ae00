"""Debug target for k_lanczos_vec: `make debugvec` builds mac_b200/libmacb200_debug.so with bounded spin loops that print which
CTA / column / inbox slot they gave up on and every CTA's progress (MACB_LIB=mac_b200/libmacb200_debug.so python tools/dbg_vec.py)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mac_b200 import synth
from mac_b200.solvers import MAC
fixed, cand, n = synth.chain_plus_random(6000, 60000, seed=3, weighted=True)
x = synth.first_k_init(60000, 12000)
mac = MAC(fixed, cand, n)
try:
    lam, v = mac.fiedler_pair(x)
    print("lambda2", lam, mac.last_info)
except Exception as e:
    print("ERR", e)
