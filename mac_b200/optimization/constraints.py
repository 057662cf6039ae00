"""LP oracles (mac/optimization/constraints.py)."""
import numpy as np

from ..utils.rounding import round_nearest


def solve_subset_box_lp(g, k):
    """constraints.py:12-22: indicator of the k largest entries of g (device radix select)."""
    return round_nearest(g, k)


def solve_box_lp(g):
    """constraints.py:24-37 (used only by the reference's own unit tests; host)."""
    g = np.asarray(g)
    solution = np.zeros_like(g)
    solution[g > 0.0] = 1.0
    return solution
