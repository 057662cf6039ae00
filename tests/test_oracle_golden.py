"""Pins the CPU oracle (oracle/mac_oracle.py) to the outputs of the unmodified reference that
tests/golden/make_golden.py recorded.  CPU only."""
import json
import os

import numpy as np
import pytest

from mac_b200 import synth
from mac_b200.g2o import split_edges
from oracle import mac_oracle as orc

pytestmark = pytest.mark.filterwarnings("ignore")


def _load(golden_dir, name):
    return json.load(open(os.path.join(golden_dir, name)))


def test_k5_known_answer(golden_dir):
    # reference tests/utils/test_fiedler.py:26-33: lambda2(K5) == 5 (np.isclose)
    fi, fj, fw = synth.complete_graph(5)
    L = orc.laplacian_from_edges(5, fi, fj, fw)
    lam, v, X = orc.find_fiedler_pair(L)
    assert np.isclose(lam, 5.0)
    assert abs(lam - _load(golden_dir, "k5.json")["lambda2"]) < 1e-12
    assert abs(np.linalg.norm(v) - 1.0) < 1e-12 and abs(v.sum()) < 1e-12


def test_laplacian_matches_dense_definition():
    # reference tests/utils/test_graphs.py:27-50 (against nx.laplacian_matrix): same property, dense.
    rng = np.random.default_rng(7)
    (fi, fj, _), (ci, cj, _), n = synth.petersen_split()
    ei, ej = np.r_[fi, ci], np.r_[fj, cj]
    w = rng.random(len(ei))
    L = orc.laplacian_from_edges(n, ei, ej, w).toarray()
    D = np.zeros((n, n))
    for a, b, ww in zip(ei, ej, w):
        D[a, a] += ww
        D[b, b] += ww
        D[a, b] -= ww
        D[b, a] -= ww
    assert np.allclose(L, D)


@pytest.mark.parametrize("k", [0, 1, 2, 3, 4, 5])
def test_petersen_solve_matches_reference(golden_dir, k):
    gold = _load(golden_dir, "petersen.json")["runs"][str(k)]
    fixed, cand, n = synth.petersen_split()
    mac = orc.OracleMAC(fixed, cand, n)
    x_init = synth.first_k_init(6, k)
    hist = []
    rounded, w, u = mac.solve(k, x_init, max_iters=100, history=hist)
    assert len(hist) == len(gold["hist"])
    assert np.allclose([h["f"] for h in hist], [h["f"] for h in gold["hist"]], rtol=1e-10, atol=0)
    assert np.allclose(w, gold["w"], rtol=0, atol=1e-13)
    assert abs(u - gold["u"]) < 1e-10
    assert (rounded == np.array(gold["rounded"])).all()
    assert abs(mac.evaluate_objective(x_init) - gold["init_l2"]) < 1e-10
    assert abs(mac.evaluate_objective(w) - gold["unrounded_l2"]) < 1e-10


def test_petersen_k3_survey_values(golden_dir):
    # SURVEY section 8c lists these by hand; keep them as a second, human-readable pin.
    gold = _load(golden_dir, "petersen.json")["runs"]["3"]
    assert abs(gold["init_l2"] - 0.527166091005) < 1e-11
    assert abs(gold["unrounded_l2"] - 1.286638585976) < 1e-11
    assert abs(gold["u"] - 1.483359171353) < 1e-11
    assert np.flatnonzero(gold["rounded"]).tolist() == [0, 3, 4]
    assert gold["w_after_1"] == [0, 0, 0, 1, 1, 1]


@pytest.mark.parametrize("name,k", [("intel", 157), ("intel", 706), ("sphere2500", 1225), ("city10000", 9619)])
def test_g2o_protocol_matches_reference(golden_dir, name, k):
    # g2o_experiment.py:306-321: naive init, max_iters=20, nearest rounding
    z = np.load(os.path.join(golden_dir, f"g2o_{name}.npz"))
    W = np.load(os.path.join(golden_dir, "g2o_fw_w.npz"))
    gold = _load(golden_dir, "g2o_fw.json")[name]["runs"][str(k)]
    fixed, cand = split_edges(z["i"], z["j"], z["kappa"])
    n = int(z["n"])
    mac = orc.OracleMAC(fixed, cand, n)
    x_init = orc.naive_greedy_subset(cand[2], k)
    assert (x_init == W[f"{name}_{k}_xinit"]).all()
    hist = []
    rounded, w, u = mac.solve(k, x_init, max_iters=20, history=hist)
    assert len(hist) == gold["iters"]
    assert np.allclose([h["f"] for h in hist], [h["f"] for h in gold["hist"]], rtol=1e-9)
    assert np.allclose(w, W[f"{name}_{k}_w"], atol=1e-12)
    assert (rounded == W[f"{name}_{k}_rounded"]).all()
    assert abs(u - gold["u"]) < 1e-9 * max(1.0, abs(gold["u"]))


def test_er2000_matches_reference(golden_dir):
    gold = _load(golden_dir, "er2000.json")
    Z = np.load(os.path.join(golden_dir, "er2000.npz"))
    fixed, cand, n = synth.chain_plus_random(2000, 20000, seed=0, weighted=True)
    mac = orc.OracleMAC(fixed, cand, n)
    x_init = synth.first_k_init(20000, 4000)
    f, g = mac.problem(x_init)
    assert abs(f - gold["lambda2_init"]) < 1e-11
    assert np.allclose(g, Z["g0"], rtol=0, atol=1e-13)
    # the ARPACK variant (the CPU baseline at headline size) agrees with TraceMIN-LU
    lam_a, v_a, _ = orc.find_fiedler_pair(mac.laplacian(x_init), method="arpack")
    assert abs(lam_a - f) < 1e-9
    assert min(np.abs(v_a - Z["v0"]).max(), np.abs(v_a + Z["v0"]).max()) < 1e-6


def test_rounding_and_lp_semantics():
    w = np.array([0.3, 0.9, 0.1, 0.9, 0.5])
    assert orc.round_nearest(w, 2).tolist() == [0, 1, 0, 1, 0]
    assert orc.round_nearest(w, 0).tolist() == [0, 0, 0, 0, 0]
    # tie on w broken by the larger weight (rounding.py:30-42)
    w = np.array([0.5, 0.5, 0.5, 0.1])
    kappa = np.array([1.0, 3.0, 2.0, 9.0])
    assert orc.round_nearest(w, 2, weights=kappa, break_ties_decimal_tol=10).tolist() == [0, 1, 1, 0]
    x = orc.round_madow_base(np.full(10, 0.3), 3, seed=np.random.RandomState(42))
    assert x.sum() == 3


def test_g2o_reader_matches_fixture_when_reference_present(golden_dir):
    path = "/root/reference/data/intel.g2o"
    if not os.path.exists(path):
        pytest.skip("reference data not on this box")
    i, j, kappa, n = orc.read_g2o_edges(path)
    z = np.load(os.path.join(golden_dir, "g2o_intel.npz"))
    assert n == int(z["n"]) and (i == z["i"]).all() and (j == z["j"]).all() and np.array_equal(kappa, z["kappa"])


def test_greedy_eig_restatement_is_greedy():
    """greedy_eig.py:86-155 restated (oracle.greedy_eig_subset): every step picks an edge at least as good as any single-edge
    alternative; lambda2 never decreases.  (The reference's own GreedyEig needs sksparse and cannot run here: unpinned.)"""
    from mac_b200 import synth
    fixed, cand, n = synth.petersen_split()
    o = orc.OracleMAC(fixed, cand, n)
    sol, evals = orc.greedy_eig_subset(o, 2)
    assert sol.sum() == 2 and evals >= 2
    first = max(range(6), key=lambda j: o.evaluate_objective(np.eye(6)[j]))
    assert o.evaluate_objective(np.eye(6)[first]) <= max(o.evaluate_objective(np.eye(6)[j]) for j in np.flatnonzero(sol)) + 1e-8
    assert o.evaluate_objective(sol) >= o.evaluate_objective(np.zeros(6)) - 1e-12
