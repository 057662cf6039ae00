import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
from mac_b200 import farm
from mac_b200.g2o import split_edges
from mac_b200.solvers import NaiveGreedy
name = sys.argv[1] if len(sys.argv) > 1 else "sphere2500"
z = np.load(f"tests/golden/g2o_{name}.npz")
fixed, cand = split_edges(z["i"], z["j"], z["kappa"]); n = int(z["n"]); m = len(cand[0])
budgets = [int(p * m) for p in (0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9)]
naive = NaiveGreedy(cand[2])
with farm.SweepPool(fixed, cand, n, streams="auto") as pool:
    pool.sweep(budgets, naive.subset, max_iters=1, comm=None)
    ref = None
    for rep in range(12):
        for mac in pool.macs:
            if mac is not None: mac._h.reset_counters()
        t = time.perf_counter()
        res = pool.sweep(budgets, naive.subset, max_iters=20, comm=None)
        dt = time.perf_counter() - t
        steps = [mac._h.counters()["lanczos_steps"] for mac in pool.macs if mac is not None]
        solves = [mac._h.counters()["fiedler_solves"] for mac in pool.macs if mac is not None]
        fb = [mac._h.device_rr_stats()["fallbacks"] for mac in pool.macs if mac is not None]
        sig = [(k, float(u), float(lam)) for (k, r, w, u, lam) in res]
        if ref is None: ref = sig
        print(f"rep {rep}: {dt*1e3:.1f} ms  total steps {sum(steps)} solves {sum(solves)} fallbacks(cum) {sum(fb)} same_results {sig == ref}", flush=True)
