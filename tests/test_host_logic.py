"""CPU-side tests: the C-ABI library loads and exports what include/macb200.h declares, the host
helpers inside it (tridiagonal Rayleigh-Ritz, pattern builder), and the Python mirror's host logic.
No CUDA compute calls here."""
import os
import re

import numpy as np
import pytest

from mac_b200 import _lib, g2o, synth
from mac_b200.optimization.constraints import solve_box_lp
from mac_b200.optimization.frankwolfe import frank_wolfe
from mac_b200.utils import graphs, rounding
from oracle import mac_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "macb200.h")).read()
    names = sorted(set(re.findall(r"\b(macb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    L = _lib.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert b"sm_100a" in L.macb_version()


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libmacb200.so")
    with pytest.raises(ImportError):
        _lib.lib()


@pytest.mark.parametrize("k", [1, 2, 3, 17, 200, 1500])
def test_tridiag_smallest_against_lapack(k):
    rng = np.random.default_rng(k)
    a = rng.uniform(0.5, 30.0, k)
    b = np.r_[0.0, rng.uniform(0.01, 6.0, max(k - 1, 0))]
    th, s = _lib.tridiag_smallest(a, b)
    T = np.diag(a) + np.diag(b[1:], 1) + np.diag(b[1:], -1)
    w, V = np.linalg.eigh(T)
    assert abs(th - w[0]) <= 1e-13 * max(1.0, np.abs(w).max())
    assert abs(abs(s @ V[:, 0]) - 1.0) < 1e-9
    assert np.linalg.norm(T @ s - th * s) < 1e-12 * np.abs(w).max()


def test_tridiag_with_tiny_coupling_and_cluster():
    # two weakly coupled blocks with (nearly) equal smallest eigenvalues: any vector of the cluster is fine
    a = np.array([2.0, 3.0, 2.0, 3.0])
    b = np.array([0.0, 1.0, 1e-14, 1.0])
    th, s = _lib.tridiag_smallest(a, b)
    T = np.diag(a) + np.diag(b[1:], 1) + np.diag(b[1:], -1)
    assert abs(th - np.linalg.eigvalsh(T)[0]) < 1e-13
    assert np.linalg.norm(T @ s - th * s) < 1e-12


def test_pattern_builder_matches_scipy_and_maps_edges():
    (fi, fj, _), (ci, cj, _), n = synth.chain_plus_random(300, 2000, seed=3)
    # add a self loop and a duplicate of a fixed edge among the candidates (both legal inputs)
    ci = np.r_[ci, 5, 0].astype(np.int32)
    cj = np.r_[cj, 5, 1].astype(np.int32)
    rp, col, eid = _lib.host_build_pattern(n, fi, fj, ci, cj)
    ei, ej = np.r_[fi, ci], np.r_[fj, cj]
    keep = ei != ej
    # duplicate edges are kept as separate slots, so compare per-row counts rather than a summed CSR
    counts = np.bincount(np.r_[ei[keep], ej[keep]], minlength=n)
    assert (np.diff(rp) == counts).all()
    rows = np.repeat(np.arange(n), np.diff(rp))
    assert ((ei[eid] == rows) & (ej[eid] == col) | (ej[eid] == rows) & (ei[eid] == col)).all()
    for r in range(n):  # rows sorted by (column, edge id)
        seg = list(zip(col[rp[r]:rp[r + 1]], eid[rp[r]:rp[r + 1]]))
        assert seg == sorted(seg)
    assert len(col) == 2 * keep.sum()


def test_pattern_builder_rejects_out_of_range():
    with pytest.raises(_lib.MacbError):
        _lib.host_build_pattern(4, [0, 1], [1, 7], [], [])


def test_laplacian_builders_match_oracle():
    (fi, fj, fw), (ci, cj, ck), n = synth.chain_plus_random(50, 120, seed=1, weighted=True)
    edges = [graphs.Edge(int(a), int(b), float(c)) for a, b, c in zip(ci, cj, ck)]
    L1 = graphs.weight_graph_lap_from_edge_list(edges, n)
    L2 = graphs.weight_graph_lap_from_edges(np.stack([ci, cj], 1), ck, n)
    L0 = orc.laplacian_from_edges(n, ci, cj, ck)
    assert abs(L1 - L0).max() == 0 and abs(L2 - L0).max() == 0
    assert graphs.select_edges(edges, np.r_[1.0, np.zeros(len(edges) - 1)]) == [edges[0]]
    mask = graphs.get_edge_selection_as_binary_mask(edges, edges[:3])
    assert mask.sum() == 3


def test_conversions_roundtrip():
    nx = pytest.importorskip("networkx")
    from mac_b200.utils.conversions import mac_to_nx, nx_to_mac
    G = nx.petersen_graph()
    edges = nx_to_mac(G)
    assert len(edges) == 15 and all(e.i < e.j and e.weight == 1.0 for e in edges)
    assert nx.is_isomorphic(mac_to_nx(edges), G)
    (fi, fj, _), (ci, cj, _), _ = synth.petersen_split()
    T = nx.minimum_spanning_tree(G)
    assert [(e.i, e.j) for e in nx_to_mac(T)] == list(zip(fi.tolist(), fj.tolist()))
    assert [(e.i, e.j) for e in nx_to_mac(nx.difference(G, T))] == list(zip(ci.tolist(), cj.tolist()))


def test_round_madow_matches_reference_loop():
    rng = np.random.default_rng(5)
    for trial in range(20):
        m, k = 200, int(rng.integers(1, 60))
        w = rng.random(m)
        w *= k / w.sum()
        w = np.minimum(w, 1.0)
        w *= k / w.sum()
        if w.max() > 1.0:
            continue
        a = rounding.round_madow_base(w, k, seed=np.random.RandomState(trial))
        b = orc.round_madow_base(w, k, seed=np.random.RandomState(trial))
        assert (a == b).all()
    best = rounding.round_madow(w, k, seed=np.random.RandomState(1), value_fn=lambda x: float(x[:10].sum()), max_iters=5)
    assert best.sum() == k


def test_frank_wolfe_generic_box_lp():
    # reference tests/optimization/test_frankwolfe.py:22-34 and :53-72
    problem = lambda x: (-np.inner(x, x), -2 * x)  # noqa: E731
    x, u = frank_wolfe(0.5 * np.ones(10), problem, solve_box_lp)
    assert np.allclose(x, np.zeros(10))
    problem = lambda x: (-np.inner(x, x) + 0.25, -2 * x)  # noqa: E731
    init = np.zeros(10)
    init[0] = 0.5
    x, u = frank_wolfe(init, problem, solve_box_lp)
    assert np.allclose(x, np.zeros(10))


def test_mac_constructor_feasibility_asserts():
    from mac_b200.solvers import MAC
    fi, fj, fw = np.array([0]), np.array([1]), np.array([1.0])
    with pytest.raises(AssertionError):  # mac.py:47: fewer than n - 1 edges
        MAC((fi, fj, fw), (np.zeros(0, int), np.zeros(0, int), np.zeros(0)), 4)
    with pytest.raises(AssertionError):  # mac.py:52: more than n(n-1)/2 edges
        MAC((np.array([0, 0]), np.array([1, 1]), np.ones(2)), (np.zeros(0, int), np.zeros(0, int), np.zeros(0)), 2)
    with pytest.raises(ValueError):      # nx:226: unknown method string
        MAC((fi, fj, fw), (np.zeros(0, int), np.zeros(0, int), np.zeros(0)), 2, fiedler_method="bogus")


def test_g2o_reader_on_synthetic_file(tmp_path):
    lines = [
        "VERTEX_SE2 0 0 0 0",
        "EDGE_SE2 0 1 1.0 0.0 0.1 10 1 0 20 0 7.5",
        "EDGE_SE2 1 2 1.0 0.0 0.1 11 2 0 21 0 8.5",
        "EDGE_SE2 0 2 2.0 0.0 0.2 12 3 0 22 0 9.5",
        "EDGE_SE3:QUAT 2 3 1 0 0 0 0 0 1  10 0 0 0 0 0 10 0 0 0 0 10 0 0 0 400 1 2 300 3 200",
        "EDGE_SE3:QUAT 0 3 1 0 0 0 0 0 1  10 0 0 0 0 0 10 0 0 0 0 10 0 0 0 100 0 0 100 0 100",
        "",
    ]
    p = tmp_path / "toy.g2o"
    p.write_text("\n".join(lines))
    i, j, kappa, tau, n = g2o.read_g2o(str(p))
    oi, oj, ok, on = orc.read_g2o_edges(str(p))
    assert n == on == 4 and (i == oi).all() and (j == oj).all()
    assert np.allclose(kappa, ok, rtol=1e-14)
    assert kappa[0] == 7.5 and abs(kappa[4] - 50.0) < 1e-12
    assert abs(tau[0] - 2.0 / np.trace(np.linalg.inv(np.array([[10.0, 1.0], [1.0, 20.0]])))) < 1e-14
    fixed, cand = g2o.split_edges(i, j, kappa)
    assert fixed[0].tolist() == [0, 1, 2] and cand[0].tolist() == [0, 0]
    (ofi, _, _), (oci, _, _) = orc.split_edges(oi, oj, ok)
    assert ofi.tolist() == fixed[0].tolist() and oci.tolist() == cand[0].tolist()


def test_g2o_reader_matches_reference_fixtures(golden_dir):
    for name in ("intel", "sphere2500", "city10000"):
        path = f"/root/reference/data/{name}.g2o"
        if not os.path.exists(path):
            pytest.skip("reference data not on this box")
        i, j, kappa, _, n = g2o.read_g2o(path)
        z = np.load(os.path.join(golden_dir, f"g2o_{name}.npz"))
        assert n == int(z["n"]) and (i == z["i"]).all() and (j == z["j"]).all()
        assert np.allclose(kappa, z["kappa"], rtol=1e-13)


def test_synthetic_generators_are_deterministic():
    a = synth.chain_plus_random(1000, 5000, seed=0)
    b = synth.chain_plus_random(1000, 5000, seed=0)
    assert all((x == y).all() for x, y in zip(a[1], b[1]))
    ci, cj, _ = a[1]
    assert (np.abs(ci - cj) > 1).all() and len(set(zip(ci.tolist(), cj.tolist()))) == 5000
    fixed, cand, n, k, x0 = synth.headline(n=2000, m=20000)
    assert k == 4000 and x0.sum() == 4000 and len(fixed[0]) == n - 1


def _check_slices(n, rp, col, eid, row_start, bankfit):
    """Invariants of the layout k_lanczos_pipe reads (csrc/api.cu build_slice_layout).  Returns (bank conflicts, groups)."""
    from mac_b200 import _lib
    S = _lib.SLICE_STRIDE
    lens = np.diff(rp)
    jrow, jlen, jcol, jeid, jw, positions = _lib.host_build_slices(n, rp, col, eid, row_start, bankfit)
    assert sorted(jrow.tolist()) == list(range(n))
    conflicts = groups = 0
    for b in range(len(row_start) - 1):
        ra, rb = int(row_start[b]), int(row_start[b + 1])
        R = rb - ra
        sa, ns = int(rp[ra]), int(rp[rb] - rp[ra])
        assert sorted(jrow[ra:rb].tolist()) == list(range(ra, rb))           # permutation inside the CTA
        assert np.array_equal(jlen[ra:rb], lens[jrow[ra:rb]])
        assert np.all(np.diff(jlen[ra:rb]) <= 0)                              # decreasing length
        # slice table: warp w owns engine rows 32 w .. 32 w + 31, padded to its first (longest) row
        nw = (R + 31) // 32
        Lw = np.array([jlen[ra + 32 * w] for w in range(nw)], dtype=np.int64)
        base = np.concatenate([[0], np.cumsum(Lw * S)])[:nw].astype(np.int64)
        assert np.array_equal(jw[b][0:2 * nw:2], base) and np.array_equal(jw[b][1:2 * nw:2], Lw)
        assert positions[b] == int((Lw * S).sum())
        if ns == 0:
            continue
        words = jcol[sa:sa + ns].astype(np.int64) & 0xffffffff
        cols, pos = words & 0x1ffff, words >> 17
        assert np.all(np.diff(cols) >= 0)                                     # column order
        assert np.all(cols != 0x1ffff)                                        # all ones is reserved for "inactive"
        assert len(set(pos.tolist())) == ns and pos.max() < positions[b]      # every slot its own position
        for g0 in range(0, ns, 16):
            grp = pos[g0:g0 + 16] & 15
            groups += 1
            conflicts += len(grp) - len(set(grp.tolist()))
        # position -> (slice w, entry d, lane l) -> engine row t.  Slices without entries share their base with the next one.
        ends = base + Lw * S
        w_of = np.searchsorted(ends, pos, side="right")
        off = pos - base[w_of]
        d_of, l_of = off // S, off % S
        assert np.all(l_of < 32)
        t_of = 32 * w_of + l_of
        assert np.all(t_of < R) and np.all(d_of < jlen[ra + t_of])            # inside the row, never in the padding
        rows = jrow[ra + t_of]                                                # caller row of every slot
        triples = sorted(zip(rows.tolist(), jrow[cols].tolist(), jeid[sa:sa + ns].tolist()))
        ref = sorted((r, int(col[s]), int(eid[s])) for r in range(ra, rb) for s in range(rp[r], rp[r + 1]))
        assert triples == ref
        # every row uses each of its entries exactly once
        per_row = {}
        for t, d in zip(t_of.tolist(), d_of.tolist()):
            per_row.setdefault(t, []).append(d)
        assert all(sorted(ds) == list(range(jlen[ra + t])) for t, ds in per_row.items())
    return conflicts, groups


@pytest.mark.parametrize("bankfit", [False, True])
def test_slice_layout_invariants(bankfit):
    """Engine numbering is a within-CTA permutation by decreasing length, every row owns exactly one position per entry,
    positions never fall into the padding, and every slot still carries its (row, column, edge) triple."""
    from mac_b200 import _lib
    fixed, cand, n = synth.chain_plus_random(700, 5000, seed=11, weighted=True)
    rp, col, eid = _lib.host_build_pattern(n, fixed[0], fixed[1], cand[0], cand[1])
    row_start = np.array([0, 150, 151, 400, 700], dtype=np.int32)      # four CTAs, one with a single row
    conflicts, groups = _check_slices(n, rp, col, eid, row_start, bankfit)
    if bankfit:
        # first fit removes most of the bank collisions a first-free-entry placement has
        base, _ = _check_slices(n, rp, col, eid, row_start, False)
        assert conflicts < 0.4 * base


def test_slice_layout_on_ragged_random_graphs():
    """Same invariants on many small ragged inputs: isolated nodes (rows without slots), CTAs without rows, CTAs with a
    single row, hubs.  Exercises the corner cases the GPU tests cannot enumerate."""
    from mac_b200 import _lib
    rng = np.random.default_rng(2024)
    for trial in range(60):
        n = int(rng.integers(2, 70))
        m = int(rng.integers(0, 4 * n))
        # a partial chain (some nodes stay isolated), random candidates, one hub
        keep = rng.random(n - 1) < 0.7
        fi = np.arange(n - 1, dtype=np.int32)[keep]
        fj = fi + 1
        ci = rng.integers(0, n, size=m).astype(np.int32)
        cj = rng.integers(0, n, size=m).astype(np.int32)
        if n > 8:
            hub = np.full(n // 2, 3, dtype=np.int32)
            ci = np.concatenate([ci, hub]); cj = np.concatenate([cj, rng.integers(0, n, size=len(hub)).astype(np.int32)])
        rp, col, eid = _lib.host_build_pattern(n, fi, fj, ci, cj)
        cuts = np.sort(rng.integers(0, n + 1, size=int(rng.integers(0, 5))))
        row_start = np.concatenate([[0], cuts, [n]]).astype(np.int32)          # may contain empty CTAs
        _check_slices(n, rp, col, eid, row_start, True)


def test_sweep_owner_matches_python_assignment():
    """macb_sweep_owner (the C-ABI's longest-first assignment of budgets to ranks) == farm.assign on the same costs."""
    from mac_b200 import farm
    m = 10688
    budgets = [int(p * m) for p in (0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9)]
    for world in (1, 2, 3, 4, 8, 16):
        owner = _lib.sweep_owner(budgets, m, world)
        parts = farm.assign([1.0 + (m - k) / m for k in budgets], world)
        for r, idxs in enumerate(parts):
            assert all(owner[i] == r for i in idxs)
        assert sorted(i for p in parts for i in p) == list(range(len(budgets)))


def test_unique_id_exchange_over_tcp():
    """The out-of-band channel of the farm (rank 0 -> everybody, 128 bytes over a socket; no torch, no GPU)."""
    import socket
    import threading
    from mac_b200 import farm
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    uid = bytes(range(128))
    got = {}

    def run(rank):
        got[rank] = farm.exchange_unique_id(rank, 3, addr="127.0.0.1", port=port, make_id=lambda: uid, timeout=30)

    ts = [threading.Thread(target=run, args=(r,)) for r in (1, 2, 0)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=60)
    assert got == {0: uid, 1: uid, 2: uid}


def test_sweep_records_round_trip_through_a_fake_allgather():
    """farm.SweepPool with several ranks: every rank packs the budgets `macb_sweep_owner` gave it, one allgather, every rank unpacks
    the same list in budget order (the GPU test of this path needs two GPUs)."""
    from mac_b200 import _lib, farm
    rng = np.random.default_rng(5)
    m = 37
    budgets = [3, 30, 11, 36, 7, 18, 25]
    for nranks in (1, 2, 3, 8):
        owner = _lib.sweep_owner(budgets, m, nranks)
        truth = {i: (k, (rng.random(m) < 0.5).astype("u1"), rng.random(m), float(rng.random()), float(rng.random())) for i, k in enumerate(budgets)}
        bufs = []
        for r in range(nranks):
            local = {i: truth[i] for i in range(len(budgets)) if owner[i] == r}
            bufs.append(farm.pack_sweep_records(local, owner, r, nranks, m))
        assert len({b.shape for b in bufs}) == 1                      # equal counts on every rank, as ncclAllGather needs
        out = farm.unpack_sweep_records(np.stack(bufs), owner, budgets, m)
        for i, (k, r_, w, u, lam) in enumerate(out):
            assert k == budgets[i] and np.array_equal(r_, truth[i][1]) and np.array_equal(w, truth[i][2])
            assert u == truth[i][3] and lam == truth[i][4]
