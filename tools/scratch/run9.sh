timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python - <<'PY'
import json, os, time, numpy as np
from mac_b200.g2o import split_edges
from mac_b200.solvers import MAC, NaiveGreedy
G = "tests/golden"
for name in ("intel", "sphere2500", "city10000"):
    z = np.load(f"{G}/g2o_{name}.npz"); fixed, cand = split_edges(z["i"], z["j"], z["kappa"]); n = int(z["n"]); m = len(cand[0])
    mac = MAC(fixed, cand, n); naive = NaiveGreedy(cand[2])
    k = int(0.2 * m); x0 = naive.subset(k)
    mac.fiedler_pair(x0)
    t0 = time.perf_counter(); lam, v = mac.fiedler_pair(x0); dt = time.perf_counter() - t0
    st = mac._h.device_rr_stats()
    print(name, "cold solve ms %.2f" % (dt * 1e3), mac.last_info, mac._h.lanczos_kernel_name(), {k_: st[k_] for k_ in ("enabled", "fallbacks", "last_status", "last_k", "last_checks", "lag_steps_at_decision")})
    t0 = time.perf_counter()
    for p in (0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9):
        kk = int(p * m); mac.solve(kk, naive.subset(kk), max_iters=20)
    print("   9-budget sweep s %.3f" % (time.perf_counter() - t0), mac._h.device_rr_stats()["fallbacks"], mac._h.counters()["lanczos_steps"])
    mac.close()
PY
