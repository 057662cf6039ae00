"""Fiedler pair of an arbitrary graph Laplacian on the device (mac/utils/fiedler.py:9-44)."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .. import _lib

_VALID = ("tracemin_lu", "tracemin_pcg", "tracemin_cholesky", "cuda", "cuda_lanczos")


def seeded_start(n, seed=None):
    """fiedler.py:27-32: a fresh RandomState(7) block (q, n) per call; the device iteration is
    single-vector and uses the first column."""
    if seed is None:
        seed = np.random.RandomState(7)
    q = min(4, n - 1)
    return np.asarray(seed.normal(size=(q, n))).T


def find_fiedler_pair(L, X=None, method="tracemin_lu", tol=1e-8, seed=None, device=-1, max_steps=0):
    """Same signature and return as the reference: (lambda2, v2[n], X[n, q]).

    Every accepted `method` string runs the device Lanczos solver (the strings select CPU
    factorisation back-ends in the reference; the stopping test is the same, nx:243).  The
    returned X carries v2 in column 0 (the remaining columns are the start block, as a
    warm-start seed has no further meaning for a single-vector iteration).
    """
    if method not in _VALID:
        raise ValueError(f"Unknown linear system solver: {method}")  # nx:226 raises NetworkXError
    n = L.shape[0]
    q = min(4, n - 1)
    if X is None:
        X = seeded_start(n, seed)
    assert X.shape[0] == L.shape[0]
    assert X.shape[1] == q
    U = sp.triu(sp.csr_matrix(L), k=1).tocoo()
    h = _lib.Handle(n, U.row, U.col, -U.data, [], [], [], device=device)
    try:
        h.set_start(np.ascontiguousarray(X[:, 0]))
        h.set_x(np.zeros(0))
        lam, v, _ = h.fiedler(tol=tol, max_steps=max_steps)
    finally:
        h.close()
    Xout = np.array(X, dtype=float, copy=True)
    Xout[:, 0] = v
    return lam, v, Xout
