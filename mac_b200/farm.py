"""K-sweep / multi-graph farm: one process per GPU, independent solves, no data-path collective.

The reference runs its budget sweep serially (examples/g2o_experiment.py:284,306-336); the
iterations share nothing but the read-only graph, so rank r simply takes its share of the
(graph, K) work items.  The only communication is a gather of the per-item results at the end
(`torch.distributed`, NCCL on GPUs / gloo in the CPU tests).  The single-graph eigen-solve is
never split across devices ("replicas only").
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


def sweep_budgets(fixed, cand, n, budgets, x_init_fn, device=None, max_iters=20, streams=1, **solve_kw):
    """The g2o protocol (g2o_experiment.py:306-321) farmed over ranks: for each budget K,
    x_init = x_init_fn(K), MAC.solve(K, x_init, max_iters=20, rounding='nearest').
    Returns [(K, rounded, w, u, lambda2_unrounded)] in budget order on every rank.

    `streams` > 1 additionally runs that many budgets of this rank's share concurrently on its GPU, each on its
    own handle / CUDA stream / host thread: a pose graph (n <= 1e4) occupies 1-20 of the 148 SMs, so independent
    budgets overlap almost perfectly.  Results do not depend on `streams` (every solve is a pure function of its
    input)."""
    from .solvers.mac import MAC
    _, local_rank, _ = dist_env()
    dev = local_rank if device is None else device
    m = len(cand[0])
    costs = [1.0 + (m - k) / max(m, 1) for k in budgets]

    def solve_with(mac, k):
        rounded, w, u = mac.solve(k, x_init_fn(k), max_iters=max_iters, **solve_kw)
        return (k, rounded.astype("u1"), w, u, mac.evaluate_objective(w))

    if streams <= 1:
        mac = MAC(fixed, cand, n, device=dev)
        try:
            return run_sweep(list(budgets), lambda k: solve_with(mac, k), costs=costs)
        finally:
            mac.close()

    import queue
    import threading
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(), dist.get_world_size()
    else:
        rank, world = 0, 1
    mine = assign(costs, world)[rank]
    todo = queue.Queue()
    for i in sorted(mine, key=lambda i: -costs[i]):
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

    threads = [threading.Thread(target=worker) for _ in range(min(streams, max(len(mine), 1)))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    if world == 1:
        return [local[i] for i in range(len(budgets))]
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    merged = {}
    for part in gathered:
        merged.update(part)
    return [merged[i] for i in range(len(budgets))]
