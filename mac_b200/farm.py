"""K-sweep / multi-graph farm: one process per GPU, independent solves, no data-path collective.

The reference runs its budget sweep serially (examples/g2o_experiment.py:284,306-336); the
iterations share nothing but the read-only graph, so rank r simply takes its share of the
(graph, K) work items.  The only communication is ONE gather of the per-item results at the end:
`sweep_budgets` does it with ncclAllGather behind the C-ABI (`macb_sweep` / `macb_comm_*`, no torch
on that path); `run_sweep` is the generic host-side variant over `torch.distributed` (gloo in the
CPU tests).  The single-graph eigen-solve is never split across devices ("replicas only").
"""
from __future__ import annotations

import os
from typing import Callable, Sequence


def assign(costs: Sequence[float], world: int):
    """Longest-processing-time-first assignment of work items to ranks.
    Returns a list (per rank) of item indices.  Low budgets run all FW iterations while high
    budgets exit early, so items carry a cost estimate."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda q: (load[q], q))
        out[r].append(i)
        load[r] += costs[i]
    return out


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def run_sweep(items: Sequence, solve_item: Callable, costs: Sequence[float] | None = None, group=None):
    """Every rank calls this with the same `items`.  `solve_item(item)` -> picklable result.
    Returns, on every rank, the list of results in item order."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    costs = [1.0] * len(items) if costs is None else list(costs)
    mine = assign(costs, world)[rank]
    local = {i: solve_item(items[i]) for i in mine}
    if world == 1:
        return [local[i] for i in range(len(items))]
    gathered = [None] * world
    dist.all_gather_object(gathered, local, group=group)
    merged = {}
    for part in gathered:
        merged.update(part)
    return [merged[i] for i in range(len(items))]


def exchange_unique_id(rank, world, addr=None, port=None, make_id=None, timeout=120.0):
    """The 128-byte NCCL unique id from rank 0 to every rank over a plain TCP socket (no torch): rank 0 listens on
    (MASTER_ADDR, MASTER_PORT + 29), the others connect (with retries) and read it."""
    import socket
    import time
    addr = addr or os.environ.get("MASTER_ADDR", "127.0.0.1")
    port = int(port if port is not None else int(os.environ.get("MASTER_PORT", 29500)) + 29)
    if rank == 0:
        uid = make_id()
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind((addr, port))
        srv.listen(world)
        srv.settimeout(timeout)
        try:
            for _ in range(world - 1):
                conn, _ = srv.accept()
                conn.sendall(uid)
                conn.close()
        finally:
            srv.close()
        return uid
    deadline = time.time() + timeout
    while True:
        try:
            c = socket.create_connection((addr, port), timeout=5.0)
            buf = b""
            while len(buf) < 128:
                part = c.recv(128 - len(buf))
                if not part:
                    break
                buf += part
            c.close()
            if len(buf) == 128:
                return buf
        except OSError:
            pass
        if time.time() > deadline:
            raise TimeoutError("no NCCL unique id from rank 0")
        time.sleep(0.05)


_COMM = None


def farm_comm(device=None):
    """The process-wide NCCL communicator of the farm (created on first use from RANK / WORLD_SIZE / MASTER_*), or None in a
    single-process run.  Pure C-ABI + sockets: no torch on this path."""
    global _COMM
    rank, local_rank, world = dist_env()
    if world <= 1:
        return None
    if _COMM is None:
        from . import _lib
        uid = exchange_unique_id(rank, world, make_id=_lib.Comm.unique_id)
        _COMM = _lib.Comm(world, rank, uid, device=local_rank if device is None else device)
    return _COMM


def sweep_budgets(fixed, cand, n, budgets, x_init_fn, device=None, max_iters=20, streams=1, comm="env", **solve_kw):
    """The g2o protocol (g2o_experiment.py:306-321) farmed over ranks: for each budget K,
    x_init = x_init_fn(K), MAC.solve(K, x_init, max_iters=20, rounding='nearest').
    Returns [(K, rounded, w, u, lambda2_unrounded)] in budget order on every rank.

    One process per GPU (torchrun or any launcher that sets RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT); the budgets are
    assigned longest-first (`macb_sweep_owner`), every rank solves its share through `macb_sweep`, and ONE ncclAllGather of
    fixed-size records (C-ABI, `macb_comm_allgather`) leaves all results on every rank -- no torch, no pickling.

    `streams` > 1 (single process only) runs that many budgets concurrently on one GPU, each on its own handle / CUDA stream /
    host thread: a pose graph (n <= 1e4) occupies 1-20 of the 148 SMs, so independent budgets overlap almost perfectly.
    Results do not depend on `streams` or on the number of ranks (every solve is a pure function of its input)."""
    import numpy as np
    from .solvers.mac import MAC
    rank, local_rank, world = dist_env()
    dev = local_rank if device is None else device
    m = len(cand[0])
    budgets = [int(k) for k in budgets]
    if comm == "env":
        comm = farm_comm(dev)
    if comm is not None or streams <= 1:
        mac = MAC(fixed, cand, n, device=dev)
        try:
            x_inits = np.stack([np.asarray(x_init_fn(k), dtype=float) for k in budgets]) if budgets else np.zeros((0, m))
            rounded, w, u, lam, iters = mac._h.sweep(comm, budgets, x_inits, max_iters=max_iters,
                                                     rel_gap_tol=solve_kw.get("relative_duality_gap_tol", 1e-4),
                                                     grad_norm_tol=solve_kw.get("grad_norm_tol", 1e-8),
                                                     min_sel_tol=mac.min_selection_weight_tol, fiedler_max_steps=mac.fiedler_max_steps)
            return [(k, rounded[i], w[i], float(u[i]), float(lam[i])) for i, k in enumerate(budgets)]
        finally:
            mac.close()

    import queue
    import threading
    costs = [1.0 + (m - k) / max(m, 1) for k in budgets]

    def solve_with(mac, k):
        rounded, w, u = mac.solve(k, x_init_fn(k), max_iters=max_iters, **solve_kw)
        return (k, rounded.astype("u1"), w, u, mac.evaluate_objective(w))

    todo = queue.Queue()
    for i in sorted(range(len(budgets)), key=lambda i: -costs[i]):
        todo.put(i)
    local, errors = {}, []

    def worker():
        mac = None
        try:
            mac = MAC(fixed, cand, n, device=dev)
            while True:
                try:
                    i = todo.get_nowait()
                except queue.Empty:
                    return
                local[i] = solve_with(mac, budgets[i])
        except Exception as e:  # surfaced below
            errors.append(e)
        finally:
            if mac is not None:
                mac.close()

    threads = [threading.Thread(target=worker) for _ in range(min(streams, max(len(budgets), 1)))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return [local[i] for i in range(len(budgets))]
