"""GreedyEig baseline on the device primitives (reference: mac/solvers/greedy_eig.py:9-155).

The reference grows the selection one edge at a time: every still-unselected candidate j whose linear upper bound
`lambda2(current) + grad[j]` can still beat the best value seen is evaluated exactly (`lambda2` of the graph with j
added), and the first candidate whose value exceeds all earlier ones by more than 1e-8 wins (greedy_eig.py:98-155).
There the exact evaluations are CHOLMOD factor up/down-dates (sksparse, not installable here); here each one is an
eigen-solve on the B200 through the same handle `MAC` uses (`MAC.problem`: assemble L(w), Fiedler pair, gradient).
Same loop, same pruning rule, same tie rule; O(K m) eigen-solves in the worst case -- a comparison baseline, not a hot path.
"""
from __future__ import annotations

import numpy as np

from ..utils.fiedler import find_fiedler_pair
from ..utils.graphs import Edge
from .mac import MAC


class GreedyEig:
    def __init__(self, odom_measurements, lc_measurements, num_poses, device=-1):
        """greedy_eig.py:10-25 (edges as lists of `Edge` or (i, j, w) array triples)."""
        self._mac = MAC(odom_measurements, lc_measurements, num_poses, device=device)
        self.num_poses = int(num_poses)
        self.weights = self._mac.weights
        self.edge_list = self._mac.edge_list

    @property
    def L_odom(self):
        return self._mac.L_fixed

    def find_fiedler_pair(self, L, method="tracemin_lu", tol=1e-8):
        """greedy_eig.py:27-47: (lambda2(L), v2(L)) of a caller-supplied Laplacian."""
        assert method != "lobpcg"   # greedy_eig.py:43
        lam, vec, _ = find_fiedler_pair(L, method="tracemin_lu", tol=tol)
        return lam, vec

    def combined_laplacian(self, w, tol=1e-10):
        """greedy_eig.py:49-64 (scipy CSR, host; the device path never needs it)."""
        w = np.asarray(w, dtype=float)
        keep = np.where(w > tol, w, 0.0)
        old = self._mac.min_selection_weight_tol
        self._mac.min_selection_weight_tol = tol
        try:
            return self._mac.laplacian(keep)
        finally:
            self._mac.min_selection_weight_tol = old

    def grad_from_fiedler(self, fiedler_vec):
        """greedy_eig.py:66-84, vectorised: grad[k] = (kappa_k (v_i - v_j)) (v_i - v_j)."""
        v = np.asarray(fiedler_vec, dtype=float)
        d = v[self.edge_list[:, 0]] - v[self.edge_list[:, 1]]
        return (self.weights * d) * d

    def subset(self, k, save_intermediate=False):
        """greedy_eig.py:86-155.  Returns (solution 0/1 vector, list of selected `Edge`)."""
        m = len(self.weights)
        solution = np.zeros(m)
        solution_l2, solution_grad = self._mac.problem(solution)   # greedy_eig.py:92-96 (fixed sub-graph)
        selected_edges = []
        self.evaluations = 0
        for _ in range(int(k)):
            best_idx, best_l2, best_grad = -1, 0.0, None
            for j in range(m):
                if solution[j] > 0:
                    continue
                # linear upper bound at w[j] = 1 (concavity of lambda2): no need to evaluate if it cannot win
                if solution_l2 + solution_grad[j] < best_l2:   # greedy_eig.py:118-120
                    continue
                w = solution.copy()
                w[j] = 1.0
                l2, grad = self._mac.problem(w)                 # greedy_eig.py:127-131: exact value + gradient
                self.evaluations += 1
                if l2 > best_l2 + 1e-8:                         # greedy_eig.py:137-141: first edge with the max wins
                    best_idx, best_l2, best_grad = j, l2, grad
            assert best_idx != -1                               # greedy_eig.py:144
            solution[best_idx] = 1.0
            solution_l2, solution_grad = best_l2, best_grad
            i, jn = self.edge_list[best_idx]
            selected_edges.append(Edge(int(i), int(jn), float(self.weights[best_idx])))
        return solution, selected_edges

    def close(self):
        self._mac.close()
