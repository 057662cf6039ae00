"""Synthetic inputs for the BASELINE.json configs (host numpy, seeded).

The constructions follow the reference's examples:
  * Petersen split            -- tests/solvers/test_mac.py:24-30 (MST fixed, rest candidate)
  * ER + forced chain         -- examples/random_graph_sparsification.py:8-18, with the chain
                                 i -> i+1 taken as the fixed set (BASELINE.json wording, and what
                                 `split_edges` yields for .g2o odometry, pose_graph_utils.py:40-43)
All graphs are returned as plain arrays: fixed = (fi, fj, fw), cand = (ci, cj, ckappa), n.
"""
from __future__ import annotations

import numpy as np


def petersen_split():
    """Petersen graph split as `nx.minimum_spanning_tree` does on unit weights
    (tests/solvers/test_mac.py:24-30).  Candidate order is that of
    `nx_to_mac(nx.difference(G, T))` (conversions.py:9-31)."""
    fixed = [(0, 1), (0, 4), (0, 5), (1, 2), (1, 6), (2, 3), (2, 7), (3, 8), (4, 9)]
    cand = [(3, 4), (5, 7), (5, 8), (6, 8), (6, 9), (7, 9)]
    fi = np.array([e[0] for e in fixed], dtype=np.int32)
    fj = np.array([e[1] for e in fixed], dtype=np.int32)
    ci = np.array([e[0] for e in cand], dtype=np.int32)
    cj = np.array([e[1] for e in cand], dtype=np.int32)
    return (fi, fj, np.ones(len(fixed))), (ci, cj, np.ones(len(cand))), 10


def complete_graph(n):
    """K_n as fixed edges only (tests/utils/test_fiedler.py:23-33: lambda2(K5) = 5)."""
    iu = np.triu_indices(n, 1)
    return iu[0].astype(np.int32), iu[1].astype(np.int32), np.ones(len(iu[0]))


def _distinct_pairs(n, count, rng):
    """`count` distinct unordered pairs (i < j) with |i - j| > 1, in random order."""
    keys = np.empty(0, dtype=np.int64)
    while len(keys) < count:
        need = count - len(keys)
        a = rng.integers(0, n, size=int(need * 1.1) + 16, dtype=np.int64)
        b = rng.integers(0, n, size=len(a), dtype=np.int64)
        lo = np.minimum(a, b)
        hi = np.maximum(a, b)
        ok = (hi - lo) > 1
        keys = np.unique(np.concatenate([keys, lo[ok] * n + hi[ok]]))
    keys = rng.permutation(keys)[:count]
    return (keys // n).astype(np.int32), (keys % n).astype(np.int32)


def chain_plus_random(n, m, seed=0, weighted=False):
    """Fixed chain (i, i+1), weight 1.0, plus `m` distinct random candidate pairs with
    |i - j| > 1.  kappa = 1.0, or U(0.5, 1.5) (seed + 1) when `weighted` -- the weighted
    variant breaks the exact ties an all-ones graph produces."""
    rng = np.random.default_rng(seed)
    fi = np.arange(n - 1, dtype=np.int32)
    fj = fi + 1
    ci, cj = _distinct_pairs(n, m, rng)
    if weighted:
        ck = np.random.default_rng(seed + 1).uniform(0.5, 1.5, size=m)
    else:
        ck = np.ones(m)
    return (fi, fj, np.ones(n - 1)), (ci, cj, ck), n


def erdos_renyi_chain(n, p, seed=0, weighted=False):
    """BASELINE config 2: G(n, p)-sized random pair set + chain; candidates are the
    sampled pairs with |i - j| > 1."""
    m = int(round(p * n * (n - 1) / 2))
    return chain_plus_random(n, m, seed=seed, weighted=weighted)


def headline(seed=0, weighted=False, n=100_000, m=1_000_000):
    """BASELINE config 5 (H): n = 100k, 1M candidates, K = 0.2 m, x_init = first-K ones."""
    fixed, cand, n = chain_plus_random(n, m, seed=seed, weighted=weighted)
    k = int(0.2 * m)
    x_init = np.zeros(m)
    x_init[:k] = 1.0
    return fixed, cand, n, k, x_init


def first_k_init(m, k):
    """tests/solvers/test_mac.py:44-45."""
    x = np.zeros(m)
    x[:k] = 1.0
    return x
