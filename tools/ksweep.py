"""BASELINE configs 3 and 4: budget sweep on pose graphs (g2o protocol, examples/g2o_experiment.py:284-336):
for K in {10..90 %} of the candidates: x_init = NaiveGreedy.subset(K), MAC.solve(K, x_init, max_iters=20, nearest).
One process per GPU under torchrun (budgets farmed longest-first, results gathered); also runs on one GPU.

    python tools/ksweep.py intel sphere2500 city10000 [--cpu]     (--cpu also times the oracle on the host, budget 20 % only)
    python -m torch.distributed.run --nproc-per-node 8 tools/ksweep.py sphere2500 city10000
"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mac_b200 import farm
from mac_b200.g2o import split_edges
from mac_b200.solvers import NaiveGreedy

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
PCTS = [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9]
STREAMS = int(os.environ.get("KSWEEP_STREAMS", 1))


def main():
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["intel", "sphere2500", "city10000"]
    rank, local_rank, world = farm.dist_env()
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    gold = json.load(open(os.path.join(G, "g2o_fw.json")))
    for name in names:
        z = np.load(os.path.join(G, f"g2o_{name}.npz"))
        fixed, cand = split_edges(z["i"], z["j"], z["kappa"]); n = int(z["n"]); m = len(cand[0])
        budgets = [int(p * m) for p in PCTS]
        naive = NaiveGreedy(cand[2])
        farm.sweep_budgets(fixed, cand, n, budgets, naive.subset, max_iters=1, streams=STREAMS)   # warm-up: contexts, kernels
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        res = farm.sweep_budgets(fixed, cand, n, budgets, naive.subset, max_iters=20, rounding="nearest", streams=STREAMS)
        if dist is not None:
            dist.barrier()
        dt = time.perf_counter() - t0
        if rank == 0:
            rows = []
            for (k, rounded, w, u, lam) in res:
                g = gold[name]["runs"].get(str(k))
                rows.append({"K": k, "lambda2_unrounded": lam, "dual_bound": u, "selected": int(rounded.sum()),
                             "ref_lambda2_unrounded": g["unrounded_l2"] if g else None})
            out = {"dataset": name, "n": n, "candidates": m, "budgets": len(budgets), "gpus": world, "streams_per_gpu": STREAMS, "sweep_seconds": dt, "results": rows}
            if "--cpu" in sys.argv and world == 1:
                from oracle import mac_oracle as orc
                o = orc.OracleMAC(fixed, cand, n)
                k = budgets[1]
                t1 = time.perf_counter(); o.solve(k, orc.naive_greedy_subset(cand[2], k), max_iters=20); out["cpu_oracle_seconds_one_budget_20pct"] = time.perf_counter() - t1
            print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
