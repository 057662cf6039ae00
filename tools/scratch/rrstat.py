import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from mac_b200 import synth
from mac_b200.solvers import MAC
fixed, cand, n, k, x0 = synth.headline()
mac = MAC(fixed, cand, n)
xs = [x0, np.full(len(x0), 0.2)]
w, u, info = mac.frank_wolfe(k, x0, 12)
xs.append(w)
for x in xs:
    lam, v = mac.fiedler_pair(x)
    st = mac._h.device_rr_stats()
    print({kk: st[kk] for kk in st if kk not in ("cycles_by_stage",)}, mac.last_info)
    for div in (8, 12, 48):
        os.environ["MACB_CHECK_DIV"] = str(div)
        m2 = MAC(fixed, cand, n)
        lam2, _ = m2.fiedler_pair(x)
        s2 = m2._h.device_rr_stats()
        print("   check_div", div, "k", s2["last_k"], "checks", s2["last_checks"], "steps", m2.last_info["steps"], "lag", s2.get("lag"))
        m2.close()
    del os.environ["MACB_CHECK_DIV"]
