import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mac_b200 import synth
from mac_b200.solvers import MAC
from oracle import mac_oracle as orc
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
gold = json.load(open(os.path.join(G, "er2000.json")))
fixed, cand, n = synth.chain_plus_random(2000, 20000, seed=0, weighted=True)
mac = MAC(fixed, cand, n); o = orc.OracleMAC(fixed, cand, n)
k = 4000
x = synth.first_k_init(20000, k)
u = np.inf
for i in range(6):
    f, g = mac.problem(x)
    fo, go = o.problem(x)
    s = mac.solve_lp(k)
    so = orc.solve_subset_box_lp(go, k)
    s2 = orc.solve_subset_box_lp(g, k)
    print(i, "f", f, fo, gold["hist"][i]["f"], "dg", np.abs(g - go).max() / go.max(), "sel diff dev-vs-oracle(g_dev)", int(np.abs(s - s2).sum()),
          "dev-vs-oracle", int(np.abs(s - so).sum()), "nsel", s.sum(), mac.last_info)
    x = x + 2.0 / (i + 2.0) * (s - x)
# now the fused loop, one iteration at a time
for iters in (1, 2, 3):
    w, uu, info = mac.frank_wolfe(k, synth.first_k_init(20000, k), iters, 0.0, 0.0)
    ho = []
    wo, uo = orc.frank_wolfe(synth.first_k_init(20000, k), o.problem, lambda g: orc.solve_subset_box_lp(g, k), maxiter=iters, relative_duality_gap_tol=0.0, grad_norm_tol=0.0, history=ho)
    print("fused", iters, info["f_hist"], [h["f"] for h in ho], "dw", np.abs(w - wo).max(), "u", uu, uo)
