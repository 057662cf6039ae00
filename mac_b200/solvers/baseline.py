"""Naive top-k-by-weight baseline (mac/solvers/baseline.py:5-16); the g2o protocol uses it to
build x_init (g2o_experiment.py:312-315).  Host only."""
import numpy as np


class NaiveGreedy:
    def __init__(self, edges):
        if isinstance(edges, np.ndarray):
            self.weights = np.asarray(edges, dtype=float)
        else:
            self.weights = np.array([e.weight for e in edges])

    def subset(self, k):
        idx = np.argpartition(self.weights, -k)[-k:]
        solution = np.zeros(len(self.weights))
        if k > 0:
            solution[idx] = 1.0
        return solution
