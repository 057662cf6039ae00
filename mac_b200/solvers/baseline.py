"""Naive top-k-by-weight baseline (reference: mac/solvers/baseline.py:5-16).  The g2o protocol uses it to build the
Frank-Wolfe starting point (g2o_experiment.py:312-315), which is why it lives on this path at all.  Host only.

Accepts either the reference's list of `Edge` tuples or a plain weight array (the array-based `MAC(...)` construction
never materialises Edge objects)."""
import numpy as np


class NaiveGreedy:
    def __init__(self, edges):
        self.weights = (np.asarray(edges, dtype=float) if isinstance(edges, np.ndarray)
                        else np.fromiter((e.weight for e in edges), dtype=float))

    def subset(self, k):
        """0/1 indicator of the k heaviest candidates.  Selection (and tie behaviour) is numpy's introselect at
        position m - k, i.e. exactly what the reference's `argpartition(w, -k)[-k:]` yields."""
        m = self.weights.shape[0]
        indicator = np.zeros(m)
        k = int(k)
        if k <= 0 or m == 0:
            return indicator
        if k >= m:
            indicator[:] = 1.0
            return indicator
        pivot = m - k
        indicator[np.argpartition(self.weights, pivot)[pivot:]] = 1.0
        return indicator
