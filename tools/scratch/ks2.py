import os, sys, time, json, numpy as np
sys.path.insert(0, os.getcwd())
from mac_b200 import farm
from mac_b200.g2o import split_edges
from mac_b200.solvers import NaiveGreedy
rank, local_rank, world = farm.dist_env()
z = np.load("tests/golden/g2o_city10000.npz")
fixed, cand = split_edges(z["i"], z["j"], z["kappa"]); n = int(z["n"]); m = len(cand[0])
budgets = [int(p * m) for p in (0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9)]
naive = NaiveGreedy(cand[2])
st = os.environ.get("ST", "auto"); st = st if st == "auto" else int(st)
farm.sweep_budgets(fixed, cand, n, budgets, naive.subset, device=local_rank, max_iters=1, streams=st)
os.environ["MACB_FARM_TRACE"] = "1"
t = time.perf_counter()
res = farm.sweep_budgets(fixed, cand, n, budgets, naive.subset, device=local_rank, max_iters=20, streams=st)
print(f"rank {rank} total {(time.perf_counter() - t) * 1e3:.1f} ms", flush=True)
