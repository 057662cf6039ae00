""".g2o pose-graph reader for MAC inputs (host; input format).

Restates what MAC consumes of examples/pose_graph_utils.py:228-351 (`read_g2o_file`), :18-45
(`split_edges`) and :381-396 (`rpm_to_mac`): per EDGE line the pair (i, j) and the rotational
concentration kappa; tau is also returned.  Batched numpy instead of a per-line 6x6 inverse.
"""
from __future__ import annotations

import numpy as np


def read_g2o(filename):
    """Returns (i[E], j[E], kappa[E], tau[E], num_poses) in file order.
    EDGE_SE2:      kappa = I33,                         tau = 2 / tr(inv([[I11,I12],[I12,I22]]))  (:332-336)
    EDGE_SE3:QUAT: kappa = 3 / (2 tr(inv(I[3:6,3:6]))), tau = 3 / tr(inv(I[0:3,0:3]))             (:296-297)
    """
    order, se2, se3 = [], [], []
    with open(filename, "r") as f:
        for line in f:
            p = line.split()
            if not p:
                continue
            if p[0] == "EDGE_SE2":
                order.append((0, len(se2)))
                se2.append(p[1:12])
            elif p[0] == "EDGE_SE3:QUAT":
                order.append((1, len(se3)))
                se3.append(p[1:31])
    E = len(order)
    ii = np.zeros(E, dtype=np.int32)
    jj = np.zeros(E, dtype=np.int32)
    kappa = np.zeros(E)
    tau = np.zeros(E)
    order = np.asarray(order, dtype=np.int64).reshape(-1, 2)
    if se2:
        a = np.asarray(se2, dtype=np.float64)
        sel = np.flatnonzero(order[:, 0] == 0)
        ii[sel] = a[:, 0].astype(np.int32)
        jj[sel] = a[:, 1].astype(np.int32)
        I11, I12, I22, I33 = a[:, 5], a[:, 6], a[:, 8], a[:, 10]
        cov = np.stack([np.stack([I11, I12], -1), np.stack([I12, I22], -1)], -2)
        tau[sel] = 2.0 / np.trace(np.linalg.inv(cov), axis1=-2, axis2=-1)
        kappa[sel] = I33
    if se3:
        a = np.asarray(se3, dtype=np.float64)
        sel = np.flatnonzero(order[:, 0] == 1)
        ii[sel] = a[:, 0].astype(np.int32)
        jj[sel] = a[:, 1].astype(np.int32)
        info = a[:, 9:30]
        full = np.zeros((len(a), 6, 6))
        iu = np.triu_indices(6)
        full[:, iu[0], iu[1]] = info
        full[:, iu[1], iu[0]] = info
        tau[sel] = 3.0 / np.trace(np.linalg.inv(full[:, 0:3, 0:3]), axis1=-2, axis2=-1)
        kappa[sel] = 3.0 / (2.0 * np.trace(np.linalg.inv(full[:, 3:6, 3:6]), axis1=-2, axis2=-1))
    num_poses = int(max(ii.max(initial=-1), jj.max(initial=-1))) + 1
    return ii, jj, kappa, tau, num_poses


def split_edges(ii, jj, w):
    """pose_graph_utils.py:18-45: fixed = |i - j| <= 1 (odometry chain), candidates = the rest."""
    ii, jj, w = np.asarray(ii), np.asarray(jj), np.asarray(w)
    loop = np.abs(jj.astype(np.int64) - ii.astype(np.int64)) > 1
    return (ii[~loop], jj[~loop], w[~loop]), (ii[loop], jj[loop], w[loop])


def load_mac_problem(filename):
    """(fixed, cand, n) ready for `MAC(fixed, cand, n)`: edges weighted by kappa (rpm_to_mac :381-396)."""
    ii, jj, kappa, _, n = read_g2o(filename)
    fixed, cand = split_edges(ii, jj, kappa)
    return fixed, cand, n
