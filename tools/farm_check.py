"""One rank of a small budget sweep over the C-ABI farm (macb_sweep + ncclAllGather).  Launched once per GPU with
RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT set (by the 2-GPU test, by `__graft_entry__.smoke()` on a
multi-GPU box, or by torchrun); prints one JSON line with every budget's result as seen by this rank."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mac_b200 import farm, synth  # noqa: E402

fixed, cand, n = synth.chain_plus_random(1500, 9000, seed=4, weighted=True)
budgets = [900, 1800, 2700, 3600]
res = farm.sweep_budgets(fixed, cand, n, budgets, lambda k: synth.first_k_init(9000, k), max_iters=5,
                         streams=int(os.environ.get("MACB_FARM_STREAMS", "1")))
rank, _, world = farm.dist_env()
print(json.dumps({"rank": rank, "world": world,
                  "results": [(int(k), int(r.sum()), float(u), float(lam), float(w.sum())) for (k, r, w, u, lam) in res]}), flush=True)
