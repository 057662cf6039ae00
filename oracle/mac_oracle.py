"""CPU oracle for MAC's Frank-Wolfe hot path.  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (numpy / scipy) of the reference algorithm on the
path `MAC.solve -> frank_wolfe -> MAC.problem -> find_fiedler_pair -> LP step`.
It is the *checker* for the CUDA path in `mac_b200/`; nothing under `mac_b200/`
imports it.  Only `tests/`, `__graft_entry__.smoke()` and the CPU-baseline /
`--impl reference` legs of `bench.py` may import this module.

Parity status: PINNED.  `tests/golden/make_golden.py` runs the unmodified
reference (`/root/reference`, networkx 3.6.1 / scipy 1.18.1) in the build
container and commits its outputs under `tests/golden/`;
`tests/test_oracle_golden.py` checks this restatement against those files.

Every function cites the reference site it restates (paths relative to
`/root/reference`; `nx:` = networkx/linalg/algebraicconnectivity.py 3.6.1,
the third-party module where the eigen-solver arithmetic lives).

Differences from the reference that are deliberate (and harmless to parity):
  * the per-edge Python loops (graphs.py:77-96, mac.py:117-124, rounding.py:34-37)
    are written as vectorised numpy -- same arithmetic, same order of
    floating-point operations per element;
  * edges are carried as arrays (i, j, w) instead of lists of `Edge` tuples.
"""
from __future__ import annotations

import numpy as np
import scipy as sp
import scipy.linalg
import scipy.sparse
import scipy.sparse.linalg


# --------------------------------------------------------------------------- graphs
def laplacian_from_edges(n, ei, ej, w):
    """Weighted graph Laplacian as CSR (graphs.py:13-48 and graphs.py:58-98).

    Four COO triplets per edge -- (i,i,+w) (j,j,+w) (i,j,-w) (j,i,-w) -- in the
    same interleaved order as the reference so that duplicate summation in
    COO->CSR sees the same sequence of addends.
    """
    ei = np.asarray(ei, dtype=np.int64)
    ej = np.asarray(ej, dtype=np.int64)
    w = np.asarray(w, dtype=np.float64)
    rows = np.stack([ei, ej, ei, ej], axis=1).ravel()
    cols = np.stack([ei, ej, ej, ei], axis=1).ravel()
    data = np.stack([w, w, -w, -w], axis=1).ravel()
    return sp.sparse.csr_matrix(sp.sparse.coo_matrix((data, (rows, cols)), shape=(n, n)))


# --------------------------------------------------------------------------- eigen-solvers
def tracemin_fiedler_lu(L, X, tol):
    """TraceMIN-Fiedler with sparse-LU inverse iteration (nx:149-253, un-normalised
    branch, method='tracemin_lu'; `_LUSolver` nx:77-105).

    Returns (sigma[q], X[n,q]).
    """
    n = X.shape[0]
    X = np.array(X, dtype=float, copy=True)

    def project(X):  # nx:206-210
        for j in range(X.shape[1]):
            X[:, j] -= X[:, j].sum() / n

    # nx:215-224: make L nonsingular by an infinite diagonal at the max-degree row.
    A = sp.sparse.csc_array(L, dtype=float, copy=True)
    i = (A.indptr[1:] - A.indptr[:-1]).argmax()
    A[i, i] = np.inf
    LU = sp.sparse.linalg.splu(  # nx:91-96
        A,
        permc_spec="MMD_AT_PLUS_A",
        diag_pivot_thresh=0.0,
        options={"Equil": True, "SymmetricMode": True},
    )

    Lnorm = abs(L).sum(axis=1).flatten().max()  # nx:229
    project(X)
    W = np.ndarray(X.shape, order="F")
    while True:  # nx:233-251
        X = np.linalg.qr(X)[0]
        W[:, :] = L @ X
        H = X.T @ W
        sigma, Y = sp.linalg.eigh(H, overwrite_a=True)
        X = X @ Y
        res = sp.linalg.blas.dasum(W @ Y[:, 0] - sigma[0] * X[:, 0]) / Lnorm
        if res < tol:
            break
        for j in range(X.shape[1]):  # nx:98-105
            W[:, j] = LU.solve(X[:, j])
        X = (sp.linalg.inv(W.T @ X) @ W.T).T
        project(X)
    return sigma, np.asarray(X)


def fiedler_arpack(L, tol):
    """The 'lanczos' branch of networkx `_get_fiedler_func` (nx:273-289):
    `eigsh(L, 2, which='SM', tol=tol)` and take the second pair.

    This is the "scipy/ARPACK CPU path" BASELINE.json names; it is the only CPU
    path that finishes at the 100k-node headline size.
    The vector is rescaled to unit 2-norm and zero mean so that it is
    interchangeable with the TraceMIN Ritz vector.
    """
    A = sp.sparse.csc_array(L, dtype=float)
    sigma, X = sp.sparse.linalg.eigsh(A, 2, which="SM", tol=tol, return_eigenvectors=True)
    order = np.argsort(sigma)
    v = X[:, order[1]].copy()
    v -= v.mean()
    v /= np.linalg.norm(v)
    return float(sigma[order[1]]), v


def seeded_start_block(n, q=None):
    """fiedler.py:27-32: fresh RandomState(7) per call, X0 = normal((q,n)).T."""
    if q is None:
        q = min(4, n - 1)
    seed = np.random.RandomState(7)
    return np.asarray(seed.normal(size=(q, n))).T


def find_fiedler_pair(L, X=None, method="tracemin_lu", tol=1e-8):
    """fiedler.py:9-44.  `method='arpack'` selects the nx 'lanczos' variant."""
    n = L.shape[0]
    q = min(4, n - 1)
    if X is None:
        X = seeded_start_block(n, q)
    assert X.shape[0] == n
    assert X.shape[1] == q
    if method == "tracemin_lu":
        sigma, X = tracemin_fiedler_lu(L, X, tol)
        return sigma[0], X[:, 0], X
    if method == "arpack":
        lam, v = fiedler_arpack(L, tol)
        return lam, v, None
    raise ValueError(f"Unknown linear system solver: {method}")  # nx:226


def residual_l1(L, lam, v):
    """The reference's convergence measure, nx:229,243: ||L v - lam v||_1 / ||L||_inf."""
    Lnorm = abs(L).sum(axis=1).flatten().max()
    return float(np.abs(L @ v - lam * v).sum() / Lnorm)


# --------------------------------------------------------------------------- rounding / LP
def round_nearest(w, k, weights=None, break_ties_decimal_tol=None):
    """rounding.py:7-42.  Structured (w, weight) argpartition built without the
    per-element tuple loop; same dtype, same `order=` call."""
    w = np.asarray(w, dtype=float)
    if weights is None or break_ties_decimal_tol is None:
        idx = np.argpartition(w, -k)[-k:]  # rounding.py:24
        rounded = np.zeros(len(w))
        if k > 0:
            rounded[idx] = 1.0
        return rounded
    truncated_w = w.round(decimals=break_ties_decimal_tol)  # rounding.py:33
    zipped = np.empty(len(w), dtype=[("w", "float"), ("weight", "float")])
    zipped["w"] = truncated_w
    zipped["weight"] = np.asarray(weights, dtype=float)
    idx = np.argpartition(zipped, -k, order=["w", "weight"])[-k:]  # rounding.py:38
    rounded = np.zeros(len(w))
    if k > 0:
        rounded[idx] = 1.0
    return rounded


def solve_subset_box_lp(g, k):
    """constraints.py:12-22."""
    return round_nearest(g, k)


def round_madow_base(w, k, seed=None):
    """rounding.py:78-95 (systematic sampling)."""
    u = np.random.rand() if seed is None else seed.rand()
    x = np.zeros(len(w))
    pi = np.zeros(len(w))
    sumw = np.cumsum(w)
    pi[1:] = sumw[:-1]
    for i in range(k):
        total = u + i
        x[np.where((pi <= total) & (total < sumw))] = 1.0
    assert np.sum(x) == k, f"Error: {np.sum(x)} != {k}"
    return x


# --------------------------------------------------------------------------- Frank-Wolfe
def naive_stepsize(k):
    """frankwolfe.py:7-8."""
    return 2.0 / (k + 2.0)


def frank_wolfe(initial, problem, solve_lp, maxiter=50, relative_duality_gap_tol=1e-5,
                grad_norm_tol=1e-10, history=None):
    """frankwolfe.py:10-79 (default step size).  `history`, if a list, receives one
    dict per iteration {f, u, gnorm, s_idx} -- an addition for the parity tests."""
    x = initial
    u = float("inf")
    for i in range(maxiter):
        f, gradf = problem(x)
        s = solve_lp(gradf)
        u = min(u, f + gradf @ (s - x))
        if history is not None:
            history.append({"f": float(f), "u": float(u), "gnorm": float(np.linalg.norm(gradf)),
                            "s_idx": np.flatnonzero(s).tolist()})
        if np.linalg.norm(gradf) < grad_norm_tol:
            return x, u
        if (u - f) < relative_duality_gap_tol * abs(f):
            return x, u
        x = x + naive_stepsize(i) * (s - x)
    return x, u


# --------------------------------------------------------------------------- MAC
class OracleMAC:
    """mac/solvers/mac.py:15-225 restated over edge arrays.

    fixed = (fi, fj, fw), cand = (ci, cj, ckappa).  `fw_fiedler_method` selects the
    eigen-solver used inside `problem` -- the reference hard-codes 'tracemin_lu'
    there (mac.py:115); 'arpack' is the swap BASELINE.md section 3 describes for
    sizes where sparse LU does not finish.
    """

    def __init__(self, fixed, cand, num_nodes, fiedler_method="tracemin_lu", fiedler_tol=1e-8,
                 min_selection_weight_tol=1e-10, fw_fiedler_method="tracemin_lu"):
        fi, fj, fw = fixed
        ci, cj, ck = cand
        num_edges = len(fi) + len(ci)
        assert (num_nodes - 1) <= num_edges  # mac.py:47
        assert num_edges <= 0.5 * num_nodes * (num_nodes - 1)  # mac.py:52
        self.num_nodes = num_nodes
        self.L_fixed = laplacian_from_edges(num_nodes, fi, fj, fw)  # mac.py:55
        self.weights = np.asarray(ck, dtype=float)  # mac.py:58-65
        self.edge_list = np.stack([np.asarray(ci, dtype=np.int64), np.asarray(cj, dtype=np.int64)], axis=1)
        self.fiedler_method = fiedler_method
        self.fiedler_tol = fiedler_tol
        self.min_selection_weight_tol = min_selection_weight_tol
        self.fw_fiedler_method = fw_fiedler_method

    def laplacian(self, x):
        """mac.py:74-89."""
        idx = np.where(x > self.min_selection_weight_tol)
        prod = x[idx] * self.weights[idx]
        e = self.edge_list[idx]
        return self.L_fixed + laplacian_from_edges(self.num_nodes, e[:, 0], e[:, 1], prod)

    def evaluate_objective(self, x):
        """mac.py:91-102."""
        return find_fiedler_pair(self.laplacian(x), method=self.fiedler_method, tol=self.fiedler_tol)[0]

    def gradient(self, v):
        """mac.py:117-124: g_k = (kappa_k (v_i - v_j)) (v_i - v_j)."""
        d = v[self.edge_list[:, 0]] - v[self.edge_list[:, 1]]
        return (self.weights * d) * d

    def problem(self, x):
        """mac.py:104-128 (cache is a no-op in the reference)."""
        f, v, _ = find_fiedler_pair(self.laplacian(x), method=self.fw_fiedler_method)
        return f, self.gradient(v)

    def solve(self, k, x_init, rounding="nearest", max_iters=5, relative_duality_gap_tol=1e-4,
              grad_norm_tol=1e-8, history=None):
        """mac.py:130-225 (nearest rounding; fallback omitted -- it raises NameError
        in the reference)."""
        m = len(self.weights)
        if k >= m:  # mac.py:173-180
            result = np.ones(m)
            return result, result, self.evaluate_objective(np.ones(m))
        assert len(x_init) == m
        w, u = frank_wolfe(x_init, self.problem, lambda g: solve_subset_box_lp(g, k),
                           maxiter=max_iters, relative_duality_gap_tol=relative_duality_gap_tol,
                           grad_norm_tol=grad_norm_tol, history=history)
        rounded = round_nearest(w, k, weights=self.weights, break_ties_decimal_tol=10)  # mac.py:207
        return rounded, w, u


def greedy_eig_subset(omac, k):
    """mac/solvers/greedy_eig.py:86-155 restated with exact eigen-solves (`omac.problem`) in place of the CHOLMOD
    factor up/down-dates (sksparse is not installable here, so the reference's own GreedyEig cannot run: this row's
    parity is UNPINNED -- the restatement is checked for the greedy property only).  Returns (solution, evaluations)."""
    m = len(omac.weights)
    solution = np.zeros(m)
    solution_l2, solution_grad = omac.problem(solution)
    evaluations = 0
    for _ in range(k):
        best_idx, best_l2, best_grad = -1, 0.0, None
        for j in range(m):
            if solution[j] > 0:
                continue
            if solution_l2 + solution_grad[j] < best_l2:   # greedy_eig.py:118-120
                continue
            w = solution.copy()
            w[j] = 1.0
            l2, grad = omac.problem(w)
            evaluations += 1
            if l2 > best_l2 + 1e-8:                         # greedy_eig.py:137-141
                best_idx, best_l2, best_grad = j, l2, grad
        assert best_idx != -1
        solution[best_idx] = 1.0
        solution_l2, solution_grad = best_l2, best_grad
    return solution, evaluations


def naive_greedy_subset(weights, k):
    """solvers/baseline.py:9-16 (without the stray prints)."""
    weights = np.asarray(weights, dtype=float)
    idx = np.argpartition(weights, -k)[-k:]
    solution = np.zeros(len(weights))
    if k > 0:
        solution[idx] = 1.0
    return solution


# --------------------------------------------------------------------------- g2o (input format)
def read_g2o_edges(filename):
    """examples/pose_graph_utils.py:228-351 restricted to what MAC consumes:
    (i, j, kappa) per EDGE line and num_poses = max id + 1.
    SE2: kappa = I33 (:336).  SE3: kappa = 3 / (2 tr(inv(I[3:6,3:6]))) (:297)."""
    ii, jj, kk = [], [], []
    num_poses = 0
    with open(filename, "r") as f:
        for line in f:
            p = [t for t in line.split(" ") if t not in ("", "\n")]
            if not p:
                continue
            if p[0] == "EDGE_SE3:QUAT":
                vals = list(map(float, p[1:]))
                i, j = int(vals[0]), int(vals[1])
                info = vals[9:]
                # upper-triangular 6x6, row-major: rotational block = I44 I45 I46 / I55 I56 / I66
                I44, I45, I46, I55, I56, I66 = info[15], info[16], info[17], info[18], info[19], info[20]
                R = np.array([[I44, I45, I46], [I45, I55, I56], [I46, I56, I66]])
                kappa = 3.0 / (2.0 * np.trace(np.linalg.inv(R)))
            elif p[0] == "EDGE_SE2":
                vals = list(map(float, p[1:]))
                i, j = int(vals[0]), int(vals[1])
                kappa = vals[10]
            else:
                continue
            ii.append(i)
            jj.append(j)
            kk.append(kappa)
            num_poses = max(num_poses, i, j)
    return np.array(ii), np.array(jj), np.array(kk, dtype=float), num_poses + 1


def split_edges(ei, ej, w):
    """examples/pose_graph_utils.py:18-45: candidate iff |i - j| > 1."""
    loop = np.abs(ej - ei) > 1
    return (ei[~loop], ej[~loop], w[~loop]), (ei[loop], ej[loop], w[loop])
