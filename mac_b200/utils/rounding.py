"""Rounding onto {0,1}^m, |w| = k (mac/utils/rounding.py).

`round_nearest(w, k)` without tie-break arguments is the LP oracle of the Frank-Wolfe loop
(constraints.py:22) and runs on the device (`macb_topk_dense`: radix select, ties at the k-th
value to the lowest index -- numpy's introselect leaves that order unspecified).
The tie-broken variant (rounding.py:30-42) is a two-key radix select on the device
(`macb_round_nearest`); Madow / random rounding run once per solve on the host as array code.
"""
from __future__ import annotations

import numpy as np

from .. import _lib


def round_nearest(w, k, weights=None, break_ties_decimal_tol=None, device=-1):
    w = np.asarray(w, dtype=np.float64)
    k = int(k)
    if weights is None or break_ties_decimal_tol is None:
        if k <= 0:  # rounding.py:26
            return np.zeros(len(w))
        if k >= len(w):
            return np.ones(len(w))
        return _lib.topk_dense(w, k, device=device)
    # rounding.py:33-42: lexicographic (w rounded to `tol` decimals, weight) top-k, two-key radix select on the device
    if k <= 0:
        return np.zeros(len(w))
    if k >= len(w):
        return np.ones(len(w))
    return _lib.round_nearest_dense(w, weights, k, int(break_ties_decimal_tol), device=device)


def round_random(w, k):
    """rounding.py:44-61 (one uniform draw per edge, in index order)."""
    w = np.asarray(w, dtype=np.float64)
    r = np.random.rand(len(w))
    return (w > r).astype(np.float64)


def round_madow_base(w, k, seed=None):
    """rounding.py:78-95: systematic (Madow) sampling."""
    w = np.asarray(w, dtype=np.float64)
    u = np.random.rand() if seed is None else seed.rand()
    sumw = np.cumsum(w)
    pi = np.zeros(len(w))
    pi[1:] = sumw[:-1]
    x = np.zeros(len(w))
    totals = u + np.arange(k)
    # element e is hit by `total` iff pi[e] <= total < sumw[e]
    lo = np.searchsorted(totals, pi, side="left")
    hi = np.searchsorted(totals, sumw, side="left")
    x[hi > lo] = 1.0
    assert np.sum(x) == k, f"Error: {np.sum(x)} != {k}"
    return x


def round_madow(w, k, seed=None, value_fn=None, max_iters=1):
    """rounding.py:63-75."""
    if value_fn is None or max_iters == 1:
        return round_madow_base(w, k, seed)
    batch = getattr(getattr(value_fn, "__self__", None), "evaluate_objectives", None)
    if batch is not None and getattr(value_fn, "__name__", "") == "evaluate_objective":
        # value_fn is MAC.evaluate_objective: draw all candidates first (same random stream: the evaluations draw nothing),
        # evaluate them in one batched device call, keep the first best -- the same result as the loop below
        xs = [round_madow_base(w, k, seed) for _ in range(max_iters)]
        vals = batch(np.stack(xs))
        return xs[int(np.argmax(vals))]   # argmax returns the FIRST maximum, as `val > best_val` does
    best_x, best_val = None, -np.inf
    for _ in range(max_iters):
        x = round_madow_base(w, k, seed)
        val = value_fn(x)
        if val > best_val:
            best_val, best_x = val, x
    return best_x
