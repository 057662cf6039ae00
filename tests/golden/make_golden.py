"""Generate the golden fixtures in this directory by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference, networkx, scipy):

    python tests/golden/make_golden.py

The reference (MarineRoboticsGroup/mac @ 60c8b6eb) is imported from /root/reference;
`examples/pose_graph_utils.py` imports matplotlib and evo at module top, neither of
which is installed, so empty stand-ins are placed in sys.modules first (its g2o
reader, split_edges and rpm_to_mac do not touch them).

Outputs (all committed; the GPU box has no /root/reference):
  k5.json            lambda2(K5) through find_fiedler_pair      (tests/utils/test_fiedler.py:26-33)
  petersen.json      MAC.solve on the Petersen split, K = 0..5  (tests/solvers/test_mac.py:35-61)
  g2o_<name>.npz     edge arrays (i, j, kappa), n from read_g2o_file + rpm_to_mac
  g2o_fw.json        g2o protocol (g2o_experiment.py:306-321): naive init, max_iters=20, nearest
  g2o_fw_w.npz       the unrounded w / rounded masks / first Fiedler vector for those runs
  er2000.json/.npz   chain + random graph n=2000 (weighted), reference as-is (tracemin_lu), 10 iterations
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "examples"))
sys.path.insert(0, REPO)

for name in ["matplotlib", "matplotlib.pyplot", "evo", "evo.core", "evo.core.trajectory", "evo.core.sync",
             "evo.core.metrics"]:
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["evo.core.trajectory"].PoseTrajectory3D = object
sys.modules["evo.core.metrics"].PoseRelation = object
sys.modules["evo.core.metrics"].Unit = object
sys.modules["evo.core"].sync = sys.modules["evo.core.sync"]
sys.modules["evo.core"].metrics = sys.modules["evo.core.metrics"]

import networkx as nx  # noqa: E402
from mac.solvers.mac import MAC  # noqa: E402
from mac.solvers.baseline import NaiveGreedy  # noqa: E402
from mac.utils.conversions import nx_to_mac  # noqa: E402
from mac.utils.fiedler import find_fiedler_pair  # noqa: E402
from mac.utils.graphs import Edge, weight_graph_lap_from_edge_list  # noqa: E402
import pose_graph_utils as pgu  # noqa: E402

from mac_b200 import synth  # noqa: E402  (generators only; no CUDA involved)


def run_solve(mac, k, x_init, max_iters, **kw):
    """MAC.solve with the per-iteration (f, g) pairs recorded by wrapping `problem`."""
    calls = []
    orig = mac.problem

    def recording(x, cache=None):
        f, g = orig(x, cache=cache)
        calls.append((float(f), g.copy(), x.copy()))
        return f, g

    mac.problem = recording
    try:
        rounded, w, u = mac.solve(k, x_init, max_iters=max_iters, **kw)
    finally:
        mac.problem = orig
    hist = []
    ub = float("inf")
    from mac.optimization.constraints import solve_subset_box_lp
    for f, g, x in calls:
        s = solve_subset_box_lp(g, k)
        ub = min(ub, f + g @ (s - x))
        kth = np.sort(g)[-k] if k > 0 else float("inf")
        nxt = np.sort(g)[-k - 1] if 0 < k < len(g) else float("-inf")
        hist.append({"f": f, "u": float(ub), "gnorm": float(np.linalg.norm(g)),
                     "gdotx": float(g @ x), "kth": float(kth), "next": float(nxt)})
    return rounded, w, float(u), hist


def edges_to_arrays(edges):
    return (np.array([e.i for e in edges], dtype=np.int32), np.array([e.j for e in edges], dtype=np.int32),
            np.array([e.weight for e in edges], dtype=float))


def main():
    # ---- K5 known answer
    edges = nx_to_mac(nx.complete_graph(5))
    lam, v, _ = find_fiedler_pair(weight_graph_lap_from_edge_list(edges, 5))
    json.dump({"lambda2": float(lam)}, open(os.path.join(HERE, "k5.json"), "w"), indent=1)

    # ---- Petersen (config 1)
    G = nx.petersen_graph()
    T = nx.minimum_spanning_tree(G)
    fixed = nx_to_mac(T)
    cand = nx_to_mac(nx.difference(G, T))
    (sfi, sfj, _), (sci, scj, _), _ = synth.petersen_split()
    assert [(e.i, e.j) for e in fixed] == list(zip(sfi.tolist(), sfj.tolist()))
    assert [(e.i, e.j) for e in cand] == list(zip(sci.tolist(), scj.tolist()))
    out = {"fixed": [(e.i, e.j) for e in fixed], "cand": [(e.i, e.j) for e in cand], "runs": {}}
    for k in range(0, 6):
        x_init = np.zeros(len(cand))
        x_init[:k] = 1.0
        mac = MAC(fixed, cand, 10)
        rounded, w, u, hist = run_solve(mac, k, x_init, 100)
        out["runs"][str(k)] = {
            "init_l2": float(mac.evaluate_objective(x_init)),
            "unrounded_l2": float(mac.evaluate_objective(w)),
            "rounded_l2": float(mac.evaluate_objective(rounded)),
            "u": u, "w": w.tolist(), "rounded": rounded.tolist(), "hist": hist,
        }
        # one-iteration run (after 1 iteration w = s_0)
        _, w1, u1, _ = run_solve(MAC(fixed, cand, 10), k, x_init, 1)
        out["runs"][str(k)]["w_after_1"] = w1.tolist()
        out["runs"][str(k)]["u_after_1"] = u1
    json.dump(out, open(os.path.join(HERE, "petersen.json"), "w"), indent=1)
    print("petersen done")

    # ---- g2o datasets (configs 3, 4)
    fw = {}
    arrays = {}
    for name, ks in [("intel", [78, 157, 392, 706]), ("sphere2500", [245, 1225, 2205]),
                     ("city10000", [1068, 5344, 9619])]:
        meas, n = pgu.read_g2o_file(os.path.join(REF, "data", name + ".g2o"))
        odom, lc = pgu.split_edges(meas)
        odom_e, lc_e = pgu.rpm_to_mac(odom), pgu.rpm_to_mac(lc)
        ai, aj, ak = edges_to_arrays(pgu.rpm_to_mac(meas))
        np.savez_compressed(os.path.join(HERE, f"g2o_{name}.npz"), i=ai, j=aj, kappa=ak, n=n)
        mac = MAC(odom_e, lc_e, n)
        naive = NaiveGreedy.__new__(NaiveGreedy)
        naive.weights = np.array([e.weight for e in lc_e])
        fw[name] = {"n": n, "num_fixed": len(odom_e), "num_cand": len(lc_e), "runs": {}}
        for k in ks:
            idx = np.argpartition(naive.weights, -k)[-k:]  # baseline.py:12 without the prints
            x_init = np.zeros(len(lc_e))
            x_init[idx] = 1.0
            L0 = mac.laplacian(x_init)
            lam0, v0, _ = find_fiedler_pair(L0)
            rounded, w, u, hist = run_solve(mac, k, x_init, 20, rounding="nearest", use_cache=True)
            fw[name]["runs"][str(k)] = {
                "naive_l2": float(lam0),
                "unrounded_l2": float(mac.evaluate_objective(w)),
                "rounded_l2": float(mac.evaluate_objective(rounded)),
                "u": u, "iters": len(hist), "hist": hist,
            }
            arrays[f"{name}_{k}_xinit"] = x_init
            arrays[f"{name}_{k}_w"] = w
            arrays[f"{name}_{k}_rounded"] = rounded
            arrays[f"{name}_{k}_v0"] = v0
            print(name, k, fw[name]["runs"][str(k)]["unrounded_l2"], len(hist))
    json.dump(fw, open(os.path.join(HERE, "g2o_fw.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(HERE, "g2o_fw_w.npz"), **arrays)

    # ---- chain + random, n = 2000, weighted, reference as-is
    (fi, fj, fwt), (ci, cj, ck), n = synth.chain_plus_random(2000, 20000, seed=0, weighted=True)
    fixed = [Edge(int(a), int(b), float(c)) for a, b, c in zip(fi, fj, fwt)]
    cand = [Edge(int(a), int(b), float(c)) for a, b, c in zip(ci, cj, ck)]
    k = 4000
    x_init = synth.first_k_init(len(cand), k)
    mac = MAC(fixed, cand, n)
    lam0, v0, _ = find_fiedler_pair(mac.laplacian(x_init))
    f0, g0 = mac.problem(x_init)
    rounded, w, u, hist = run_solve(mac, k, x_init, 10, relative_duality_gap_tol=0.0, grad_norm_tol=0.0)
    json.dump({"n": n, "m": len(cand), "k": k, "seed": 0, "weighted": True, "lambda2_init": float(lam0),
               "u": u, "hist": hist, "unrounded_l2": float(mac.evaluate_objective(w)),
               "rounded_l2": float(mac.evaluate_objective(rounded))},
              open(os.path.join(HERE, "er2000.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(HERE, "er2000.npz"), w=w, rounded=rounded, v0=v0, g0=g0)
    print("er2000 done", lam0, u)


if __name__ == "__main__":
    main()
