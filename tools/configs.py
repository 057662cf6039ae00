"""BASELINE.json configs 1-3 on one GPU: one JSON line per config (throughput + parity against the oracle / goldens).
Config 4 (8-GPU farm) is tools/ksweep.py under torchrun; config 5 is bench.py.

    python tools/configs.py            # all
    python tools/configs.py er10k      # BASELINE configs[1] only
"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mac_b200 import synth
from mac_b200.g2o import split_edges
from mac_b200.solvers import MAC, NaiveGreedy
from oracle import mac_oracle as orc

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
which = set(sys.argv[1:]) or {"petersen", "er10k", "intel"}


def petersen():
    fixed, cand, n = synth.petersen_split()
    gold = json.load(open(os.path.join(G, "petersen.json")))["runs"]["3"]
    mac = MAC(fixed, cand, n)
    t0 = time.perf_counter()
    r, w, u = mac.solve(3, synth.first_k_init(6, 3), max_iters=100)
    dt = time.perf_counter() - t0
    print(json.dumps({"config": "Petersen K=3 (configs[0])", "seconds": dt, "iters": mac.last_info["iters"],
                      "max_abs_dw_vs_reference": float(np.abs(w - np.array(gold["w"])).max()), "du": u - gold["u"],
                      "rounded_equal": bool((r == np.array(gold["rounded"])).all())}), flush=True)
    mac.close()


def er10k():
    fixed, cand, n = synth.erdos_renyi_chain(10_000, 0.01, seed=0, weighted=True)
    m = len(cand[0]); k = int(0.2 * m)
    x0 = synth.first_k_init(m, k)
    mac = MAC(fixed, cand, n)
    mac.frank_wolfe(k, x0, 2, 0.0, 0.0)   # warm-up: engine set-up, kernels
    t0 = time.perf_counter()
    w, u, info = mac.frank_wolfe(k, x0, 20, 0.0, 0.0)
    dt = time.perf_counter() - t0
    c = mac._h.counters()
    # parity: first Fiedler value and gradient against the oracle's ARPACK path (sparse LU takes 27 s per solve here)
    o = orc.OracleMAC(fixed, cand, n, fw_fiedler_method="arpack")
    t1 = time.perf_counter(); f_o, g_o = o.problem(x0); t_cpu = time.perf_counter() - t1
    f, g = mac.problem(x0)
    print(json.dumps({"config": "ER n=10000 p=0.01 + chain, K=0.2m, weighted (configs[1])", "candidates": m, "fw_iters": 20, "seconds": dt,
                      "fw_iters_per_sec": 20 / dt, "lanczos_kernel": mac._h.lanczos_kernel_name(),
                      "lambda2_first": f, "rel_err_vs_oracle": abs(f - f_o) / f_o, "grad_max_rel_diff": float(np.abs(g - g_o).max() / g_o.max()),
                      "cpu_oracle_seconds_one_problem_eval": t_cpu, "final_lambda2": float(info["f_hist"][-1]), "dual_bound": u}), flush=True)
    mac.close()


def intel():
    z = np.load(os.path.join(G, "g2o_intel.npz"))
    gold = json.load(open(os.path.join(G, "g2o_fw.json")))["intel"]["runs"]
    fixed, cand = split_edges(z["i"], z["j"], z["kappa"]); n = int(z["n"]); m = len(cand[0])
    naive = NaiveGreedy(cand[2])
    mac = MAC(fixed, cand, n)
    mac.solve(int(0.5 * m), naive.subset(int(0.5 * m)), max_iters=1)
    rows = []
    t0 = time.perf_counter()
    for p in (0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9):
        k = int(p * m)
        r, w, u = mac.solve(k, naive.subset(k), max_iters=20, rounding="nearest")
        g = gold.get(str(k))
        rows.append({"K": k, "iters": mac.last_info["iters"], "lambda2_unrounded": float(mac.last_info["f_hist"][-1]) if len(mac.last_info["f_hist"]) else None,
                     "dual_bound": u, "ref_dual_bound": g["u"] if g else None})
    dt = time.perf_counter() - t0
    print(json.dumps({"config": "intel.g2o K sweep 10-90 % (configs[2])", "seconds_9_budgets": dt, "fw_iters_total": sum(r["iters"] for r in rows),
                      "lanczos_kernel": mac._h.lanczos_kernel_name(), "results": rows}), flush=True)
    mac.close()


for name, fn in (("petersen", petersen), ("er10k", er10k), ("intel", intel)):
    if name in which:
        fn()
