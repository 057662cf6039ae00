"""world_size-2 gloo test of the K-sweep farm's host logic (CPU; the solver is a stand-in)."""
import os
import socket
import sys

import numpy as np
import pytest

from mac_b200 import farm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_assign_longest_first_balances():
    parts = farm.assign([9, 8, 7, 6, 5, 4, 3, 2, 1], 4)
    assert sorted(i for p in parts for i in p) == list(range(9))
    loads = [sum([9, 8, 7, 6, 5, 4, 3, 2, 1][i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= 3
    assert farm.assign([1.0, 1.0], 4)[2:] == [[], []]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        items = [10, 20, 30, 40, 50]
        seen = []

        def solve(k):
            seen.append(k)
            return (k, np.full(3, k, dtype="u1"), float(k) * 0.5, rank)

        out = farm.run_sweep(items, solve, costs=[5, 4, 3, 2, 1])
        q.put((rank, seen, [(o[0], o[1].tolist(), o[2], o[3]) for o in out]))
    finally:
        dist.destroy_process_group()


def test_run_sweep_two_ranks_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    (r0, seen0, out0), (r1, seen1, out1) = res
    assert sorted(seen0 + seen1) == [10, 20, 30, 40, 50] and seen0 and seen1
    assert out0 == out1                                  # every rank holds the full, ordered result list
    assert [o[0] for o in out0] == [10, 20, 30, 40, 50]
    assert {o[3] for o in out0} == {0, 1}                # both ranks contributed
    for k, arr, half, _ in out0:
        assert arr == [k] * 3 and half == k * 0.5


def test_run_sweep_single_process_needs_no_group():
    out = farm.run_sweep([3, 1, 2], lambda k: k * k)
    assert out == [9, 1, 4]
