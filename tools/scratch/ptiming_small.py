import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mac_b200 import _lib
from mac_b200.solvers import MAC, NaiveGreedy
from mac_b200.g2o import split_edges
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for name, k in (("intel", 157),):
    z = np.load(os.path.join(G, f"g2o_{name}.npz")); fixed, cand = split_edges(z["i"], z["j"], z["kappa"]); n = int(z["n"])
    mac = MAC(fixed, cand, n); x0 = NaiveGreedy(cand[2]).subset(k)
    mac.fiedler_pair(x0); lam, v = mac.fiedler_pair(x0)
    L = _lib.lib(); ncta = C.c_int(); buf = np.zeros((64, 1, 4), dtype=np.int64)
    L.macb_debug_ptiming(mac._h._h, buf.ctypes.data_as(C.c_void_p), C.byref(ncta))
    t = buf.ravel()
    print(name, mac.last_info, "kernel cycles", t[0], "phases", t[1], "cycles/phase %.0f" % (t[0] / max(t[1], 1)), "pass1 %.0f pass2 %.0f reduce+scalars %.0f" % (t[2] / max(t[1], 1), t[3] / max(t[1], 1), t[4] / max(t[1], 1)))
