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


def pack_sweep_records(local, owner, rank, nranks, m):
    """This rank's results as one fixed-size array for `macb_comm_allgather`: a row per budget the rank owns (in budget order),
    [rounded mask (m), relaxed solution (m), dual bound, lambda2]; every rank sends max-over-ranks rows (zero padded)."""
    import numpy as np
    per_rank = [[i for i in range(len(owner)) if owner[i] == r] for r in range(nranks)]
    slots = max(1, max(len(p) for p in per_rank))
    buf = np.zeros((slots, 2 * m + 2))
    for j, i in enumerate(per_rank[rank]):
        k, rounded, w, u, lam = local[i]
        buf[j, :m] = rounded
        buf[j, m:2 * m] = w
        buf[j, 2 * m] = u
        buf[j, 2 * m + 1] = lam
    return buf


def unpack_sweep_records(allb, owner, budgets, m):
    """Inverse of `pack_sweep_records` over the gathered array [nranks, rows, 2 m + 2]: results in budget order."""
    nranks = allb.shape[0]
    out = [None] * len(budgets)
    for r in range(nranks):
        for j, i in enumerate([i for i in range(len(budgets)) if owner[i] == r]):
            rec = allb[r, j]
            out[i] = (budgets[i], rec[:m].astype("u1"), rec[m:2 * m].copy(), float(rec[2 * m]), float(rec[2 * m + 1]))
    return out


class SweepPool:
    """The graph of one budget sweep resident on this rank's GPU, `streams` times: one MAC handle (own CUDA stream, own host
    thread while a sweep runs) per budget solved concurrently.  A pose graph (n <= 1e4) occupies 6-21 of the 148 SMs, so
    independent budgets overlap almost perfectly; `streams="auto"` asks the engine how many of its launches fit side by side
    (`macb_lanczos_footprint`).  The counterpart of the one `MAC(...)` the reference builds per dataset before its budget loop
    (g2o_experiment.py:284); reuse it for several sweeps of the same graph, `close()` it when done."""

    def __init__(self, fixed, cand, n, device=None, streams="auto"):
        from .solvers.mac import MAC
        _, local_rank, _ = dist_env()
        self.dev = local_rank if device is None else device
        self.fixed, self.cand, self.n, self.m = fixed, cand, n, len(cand[0])
        self.macs = [MAC(fixed, cand, n, device=self.dev)]
        self._built = set()
        if streams == "auto":
            ctas, sms = self.macs[0]._h.lanczos_footprint()
            streams = max(1, min(sms // max(ctas, 1), 8))
        self.streams = max(1, int(streams))

    def _mac(self, slot):
        """Handle of worker `slot` (created by the worker itself on first use: the host-side layout builds run in parallel)."""
        from .solvers.mac import MAC
        while len(self.macs) <= slot:
            self.macs.append(None)
        if self.macs[slot] is None:
            self.macs[slot] = MAC(self.fixed, self.cand, self.n, device=self.dev)
        if slot not in self._built:
            # build the engine now (layout, basis: a multi-GB cudaMalloc synchronises the device), not inside some later sweep in
            # which this worker happens to get its first budget
            self.macs[slot]._h.lanczos_footprint()
            self._built.add(slot)
        return self.macs[slot]

    def sweep(self, budgets, x_init_fn, max_iters=20, comm="env", **solve_kw):
        """[(K, rounded, w, u, lambda2_unrounded)] in budget order, identical on every rank."""
        import queue
        import threading
        import numpy as np
        from . import _lib
        m = self.m
        budgets = [int(k) for k in budgets]
        if comm == "env":
            comm = farm_comm(self.dev)
        if self.streams <= 1:
            mac = self.macs[0]
            x_inits = np.stack([np.asarray(x_init_fn(k), dtype=float) for k in budgets]) if budgets else np.zeros((0, m))
            rounded, w, u, lam, iters = mac._h.sweep(comm, budgets, x_inits, max_iters=max_iters,
                                                     rel_gap_tol=solve_kw.get("relative_duality_gap_tol", 1e-4),
                                                     grad_norm_tol=solve_kw.get("grad_norm_tol", 1e-8),
                                                     min_sel_tol=mac.min_selection_weight_tol, fiedler_max_steps=mac.fiedler_max_steps)
            return [(k, rounded[i], w[i], float(u[i]), float(lam[i])) for i, k in enumerate(budgets)]

        nranks = comm.nranks if comm is not None else 1
        myrank = comm.rank if comm is not None else 0
        owner = _lib.sweep_owner(budgets, m, nranks) if budgets else np.zeros(0, dtype=np.int32)
        mine = [i for i in range(len(budgets)) if owner[i] == myrank]
        costs = [1.0 + (m - k) / max(m, 1) for k in budgets]
        todo = queue.Queue()
        for i in sorted(mine, key=lambda i: -costs[i]):
            todo.put(i)
        local, errors = {}, []

        def worker(slot):
            try:
                mac = self._mac(slot)
                while True:
                    try:
                        i = todo.get_nowait()
                    except queue.Empty:
                        return
                    k = budgets[i]
                    rounded, w, u = mac.solve(k, x_init_fn(k), max_iters=max_iters, **solve_kw)
                    local[i] = (k, rounded.astype("u1"), w, u, mac.evaluate_objective(w))
            except Exception as e:  # surfaced below
                errors.append(e)

        nthreads = min(self.streams, max(len(mine), 1))
        while len(self.macs) < nthreads:
            self.macs.append(None)
        if any(slot not in self._built for slot in range(nthreads)):   # first use: all handles and engines exist before any budget is solved
            builders = [threading.Thread(target=self._mac, args=(slot,)) for slot in range(nthreads)]
            for t in builders:
                t.start()
            for t in builders:
                t.join()
        threads = [threading.Thread(target=worker, args=(slot,)) for slot in range(nthreads)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        if comm is None:
            return [local[i] for i in range(len(budgets))]
        return unpack_sweep_records(comm.allgather(pack_sweep_records(local, owner, myrank, nranks, m)), owner, budgets, m)

    def close(self):
        for mac in self.macs:
            if mac is not None:
                mac.close()
        self.macs = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def sweep_budgets(fixed, cand, n, budgets, x_init_fn, device=None, max_iters=20, streams=1, comm="env", pool=None, **solve_kw):
    """The g2o protocol (g2o_experiment.py:306-321) farmed over ranks: for each budget K,
    x_init = x_init_fn(K), MAC.solve(K, x_init, max_iters=20, rounding='nearest').
    Returns [(K, rounded, w, u, lambda2_unrounded)] in budget order on every rank.

    One process per GPU (torchrun or any launcher that sets RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT); the budgets are
    assigned longest-first (`macb_sweep_owner`), every rank solves its share, and ONE ncclAllGather of fixed-size records
    (C-ABI, `macb_comm_allgather`) leaves all results on every rank -- no torch, no pickling.  With `streams` <= 1 all of it
    happens inside `macb_sweep`.

    `streams` > 1 (or "auto": as many as fit) runs that many of a rank's budgets concurrently on its GPU (`SweepPool`); pass a
    `pool` to reuse its handles across calls.  Results do not depend on `streams` or on the number of ranks (every solve is a
    pure function of its input)."""
    if pool is not None:
        return pool.sweep(budgets, x_init_fn, max_iters=max_iters, comm=comm, **solve_kw)
    with SweepPool(fixed, cand, n, device=device, streams=streams) as p:
        return p.sweep(budgets, x_init_fn, max_iters=max_iters, comm=comm, **solve_kw)


