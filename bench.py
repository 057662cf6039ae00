"""Headline benchmark: Frank-Wolfe iterations/sec (= Fiedler solves/sec) on BASELINE.json configs[4]
(chain + random graph, n = 100 000, 1 000 000 candidate edges, K = 200 000), one graph per GPU.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU algorithm (oracle port) on the host cores

A "step" is one Frank-Wolfe iteration: assemble L(x) -> Fiedler pair -> gradient -> top-K LP -> dual bound,
stop tests -> x update.  Early exit is disabled (both tolerances 0) so exactly K steps run.

Timing: `value` = N*K / (max over ranks of the summed per-iteration CUDA-event times); every iteration is
bracketed by its own event pair on the library's stream and a 512 MB buffer (> 126 MB L2) is rewritten between
iterations outside the brackets.  `e2e` = the same K iterations through the public `MAC.frank_wolfe` call with
host numpy buffers (x_init uploaded, w / u / histories downloaded inside the timed region), wall clock, no
bench hooks.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fw_iters_per_sec"
UNIT = "it/s"
N_NODES, N_CAND, BUDGET_FRAC = 100_000, 1_000_000, 0.2
PARITY_TOL = 1e-6   # north-star: Fiedler value within 1e-6 relative of the reference CPU path
WORKLOAD = "chain+random graph n=100000, 1000000 candidate edges, K=200000, x_init=first-K (BASELINE configs[4])"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def mark(self):
        """Samples taken before this call (warm-up) are dropped."""
        self.f.flush()
        try:
            self.skip = len(open(self.f.name).read().strip().splitlines())
        except OSError:
            self.skip = 0

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        rows = rows[max(getattr(self, "skip", 0) - 1, 0):]
        os.unlink(self.f.name)
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); smax.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for name, val in zip(names, r[5:9]):
                if val.strip().lower() == "active":
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power), "samples": len(sm),
                "reasons": sorted(reasons)}


def make_problem(seed):
    from mac_b200 import synth
    return synth.headline(seed=seed, n=N_NODES, m=N_CAND)


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_fw(fixed, cand, n, k, x0, budget_s, max_steps):
    """The reference's algorithm on the host (oracle port; eigen-solve = the scipy/ARPACK path BASELINE.md
    section 3 names, since sparse LU does not finish at this size).  Runs whole FW iterations until
    `budget_s` seconds or `max_steps`; returns (iterations, seconds)."""
    from oracle import mac_oracle as orc
    mac = orc.OracleMAC(fixed, cand, n, fw_fiedler_method="arpack")
    x, u, done = x0, float("inf"), 0
    fs = []
    t0 = time.perf_counter()
    while done < max_steps:
        f, g = mac.problem(x)
        fs.append((float(f), x))
        s = orc.solve_subset_box_lp(g, k)
        u = min(u, f + g @ (s - x))
        x = x + orc.naive_stepsize(done) * (s - x)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    return done, time.perf_counter() - t0, fs


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU legs run on rank 0 alone and may use the whole host."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=len(os.sched_getaffinity(0)))
    except Exception:
        pass


def threads_used():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return 1


def run_reference(args, rank, world):
    if rank != 0:
        return
    use_all_host_threads()
    fixed, cand, n, k, x0 = make_problem(0)
    if args.warmup > 0:
        cpu_fw(fixed, cand, n, k, x0, 0.0, 1)  # one untimed iteration: imports, page-in
    iters, secs, _ = cpu_fw(fixed, cand, n, k, x0, args.cpu_budget, args.steps)
    value = iters / secs
    sample = f"{iters} whole FW iterations of the same workload (time-boxed to {args.cpu_budget:.0f} s of the requested {args.steps})"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "cpu_path": "oracle port of MAC.solve's loop, eigen-solve = scipy ARPACK eigsh(which='SM') "
                   "(networkx 'lanczos'); the reference's default sparse-LU TraceMIN does not finish one solve at this size "
                   "(BASELINE.md: > 3000 s)", "steps_measured": iters},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads_used(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ K-sweep (BASELINE configs[3])
def ksweep(world, rank, local_rank):
    """sphere2500 + city10000, nine budgets each (10..90 % of the candidates), the protocol of g2o_experiment.py:306-321,
    farmed one-budget-per-GPU through the C-ABI (`macb_sweep`: longest-first assignment, ONE ncclAllGather of the results).
    Every rank returns the same dict: seconds per dataset (max over ranks) and the deviation from the reference's own
    results (tests/golden/g2o_ksweep.json, generated by running the unmodified reference)."""
    from mac_b200 import farm
    from mac_b200.g2o import split_edges
    from mac_b200.solvers import NaiveGreedy
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "g2o_ksweep.json")))
    comm = farm.farm_comm(local_rank)

    def sync_max(x):
        return float(comm.allgather(np.array([x])).max()) if comm is not None else float(x)

    out = {}
    for name in ("sphere2500", "city10000"):
        z = np.load(os.path.join(ROOT, "tests", "golden", f"g2o_{name}.npz"))
        fixed, cand = split_edges(z["i"], z["j"], z["kappa"])
        n, m = int(z["n"]), len(cand[0])
        budgets = [int(p * m) for p in (0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9)]
        naive = NaiveGreedy(cand[2])
        streams = os.environ.get("MACB_KSWEEP_STREAMS", "auto")   # budgets of a rank solved concurrently on its GPU
        streams = streams if streams == "auto" else int(streams)
        # the graph resident on the GPU before the loop over budgets, as in g2o_experiment.py:284 (one MAC per dataset)
        with farm.SweepPool(fixed, cand, n, device=local_rank, streams=streams) as pool:
            pool.sweep(budgets, naive.subset, max_iters=1)   # warm-up: contexts, engines
            sync_max(0.0)
            t0 = time.perf_counter()
            res = pool.sweep(budgets, naive.subset, max_iters=20)
            dt = sync_max(time.perf_counter() - t0)
            streams = pool.streams
        runs = gold[name]["runs"]
        dlam = [abs(lam - runs[str(k)]["unrounded_l2"]) / runs[str(k)]["unrounded_l2"] for (k, r, w, u, lam) in res]
        du = [abs(u - runs[str(k)]["u"]) / runs[str(k)]["u"] for (k, r, w, u, lam) in res]
        out[name] = {"seconds": dt, "budgets": len(budgets), "max_rel_dlambda2_vs_reference": max(dlam),
                     "median_rel_dlambda2_vs_reference": float(np.median(dlam)), "max_rel_du_vs_reference": max(du),
                     "selected_ok": all(int(r.sum()) == k for (k, r, w, u, lam) in res), "streams_per_gpu": streams}
    out["note"] = ("deviations above ~1e-6 come from budgets whose Frank-Wolfe trajectory crosses an LP tie (city10000: all kappa = 100; "
                   "tests/test_gpu_parity.py asserts the fork happens inside the tie window)")
    return out


def hbm_spmv_point(local_rank, peak):
    """The HBM-bound SpMV point the north-star asks for: a matrix ten times the L2 (n = 4M, 88M off-diagonals, 1.18 GB
    algorithmic, band |i - j| <= 2000 like a pose graph), best of the two SpMV kernels, against the measured copy bandwidth."""
    from mac_b200 import _lib
    n, m, band = 4_000_000, 40_000_000, 2000
    rng = np.random.default_rng(0)
    fi = np.arange(n - 1, dtype=np.int32)
    a = rng.integers(0, n, size=m, dtype=np.int64)
    b = a + rng.integers(2, band, size=m, dtype=np.int64)
    b = np.where(b >= n, a - (b - a), b)
    ok = np.abs(a - b) > 1
    ci, cj = a[ok].astype(np.int32), b[ok].astype(np.int32)
    h = _lib.Handle(n, fi, fi + 1, np.ones(n - 1), ci, cj, np.ones(len(ci)), device=local_rank)
    h.set_x(np.ones(len(ci)))
    best = None
    for engine, name in ((0, "k_spmv"), (1, "k_spmv_jds")):
        h.spmv_engine(engine)
        ms, by = h.spmv_bench(20, False)
        gbs = by / ms / 1e6
        if best is None or gbs > best["GBs"]:
            best = {"kernel": name, "GBs": gbs, "frac": gbs / peak, "ms": ms, "algorithmic_GB": by / 1e9, "n": n,
                    "nnz_offdiag": h.sizes()["nnz_union"], "matrix": f"chain + random candidates with |i-j| <= {band}"}
    h.close()
    return best


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args, rank, local_rank, world):
    from mac_b200.solvers import MAC

    dist = None
    if world > 1:
        # rank 0 prints ONE JSON line on stdout: keep NCCL's version banner (NCCL_DEBUG=VERSION on some boxes) off it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(vals):
        if dist is None:
            return list(vals)
        import torch
        t = torch.tensor(list(vals), dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    parity_failed, rel = False, []
    # One replica of THE benchmark graph per GPU (weak scaling: per-GPU work is exactly the N = 1 work).  Round 1 gave rank r the
    # graph of seed r; those graphs need 204 - 249 Lanczos steps per solve (measured), so the max over ranks then measured the
    # hardest seed, not the scaling.
    fixed, cand, n, k, x0 = make_problem(0)
    mac = MAC(fixed, cand, n, device=local_rank)
    h = mac._h
    K, W = args.steps, max(args.warmup, 0)

    sampler = ClockSampler(local_rank)
    sampler.start()
    if W > 0:
        mac.frank_wolfe(k, x0, W, 0.0, 0.0)
    time.sleep(0.3)  # let nvidia-smi take its first samples before the timed region
    sampler.mark()

    # ---- device-timed region: per-iteration CUDA events, L2 flushed between iterations
    h.set_bench(True, True)
    h.reset_counters()
    h.device_sync(); barrier()
    t0 = time.perf_counter()
    w, u, info = mac.frank_wolfe(k, x0, K, 0.0, 0.0)
    h.device_sync(); barrier()
    wall_timed = time.perf_counter() - t0
    iter_ms = h.iter_ms()
    dev_s = float(iter_ms.sum()) / 1e3
    counters = h.counters()
    lz = h.lanczos_kernel_time()
    assert info["iters"] == K and len(iter_ms) == K

    # ---- end-to-end region: the public call, host buffers, no bench hooks
    h.set_bench(False, False)
    h.device_sync(); barrier()
    t0 = time.perf_counter()
    w2, u2, info2 = mac.frank_wolfe(k, x0, K, 0.0, 0.0)
    h.device_sync(); barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    assert np.array_equal(w, w2)

    dev_max, e2e_max = max_over_ranks([dev_s, e2e_s])
    rr_stats = h.device_rr_stats()
    sweep = None
    if not args.no_ksweep:
        sweep = ksweep(world, rank, local_rank)   # every rank takes part (one budget per GPU, NCCL gather behind the C-ABI)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: the persistent Lanczos kernel, timed live inside the timed region (CUDA events
    # around its launches on the library stream).  Bytes per step follow the ACTIVE slots of every launch (sum over
    # launches of steps x bytes(nnz_active), macb_lanczos_kernel_time).
    from mac_b200 import _lib
    peak, peak_kind = peaks()
    l2_peak = _lib.measure_l2_bandwidth(local_rank, 24 << 20, 20)
    h.set_x(w)
    spmv_ms, algo_bytes = h.spmv_bench(300, False)
    spmv_cold_ms, _ = h.spmv_bench(30, True)
    sizes = h.sizes()
    spmv_achieved = algo_bytes / (spmv_ms * 1e-3) / 1e9
    lz_us = lz["ms"] * 1e3 / max(lz["phases"], 1)
    achieved = lz["algo_bytes_per_phase"] / (lz_us * 1e-6) / 1e9 if lz["phases"] else spmv_achieved
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "lanczos_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_step")
    hbm_spmv = None
    if world == 1 and not args.no_hbm_spmv:
        hbm_spmv = hbm_spmv_point(local_rank, peak)

    line = {
        "metric": METRIC, "value": world * K / dev_max, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_max * 1e3 / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": WORKLOAD, "parallelism": f"one replica of the benchmark graph per GPU x{world}, no data-path collective",
            "l2": "512 MB buffer rewritten between timed iterations (outside the event brackets); each iteration also "
                  "writes its own Lanczos basis (> L2)",
            "timing": "sum of per-iteration CUDA-event brackets on the library stream, max over ranks",
            "lanczos_steps_per_solve": counters["lanczos_steps"] / max(counters["fiedler_solves"], 1),
            "lanczos_us_per_step": lz_us, "lanczos_ms_per_iter": lz["ms"] / K,
            "other_kernels_ms_per_iter": dev_s * 1e3 / K - lz["ms"] / K,
            "nnz_union": sizes["nnz_union"], "nnz_active_final": sizes["nnz_active"],
            "wall_s_timed_region_incl_flush": wall_timed, "final_lambda2": float(info["f_hist"][-1]), "dual_bound": u,
            "spmv_us_l2_resident": spmv_ms * 1e3, "spmv_us_l2_flushed": spmv_cold_ms * 1e3,
            "roofline_peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})",
            "stop_decision": ("on the device (Rayleigh-Ritz CTA of the Lanczos launch); one host synchronisation per FW iteration"
                              if rr_stats["enabled"] else "host Rayleigh-Ritz"),
            "device_rr_fallbacks": rr_stats["fallbacks"],
            "ksweep": sweep,
        },
        "e2e": {"value": world * K / e2e_max, "unit": UNIT,
                "h2d_bytes_per_step": 8 * len(x0) / K, "d2h_bytes_per_step": (8 * len(x0) + 16 * K + 16) / K,
                "api": "MAC.frank_wolfe(k, x_init, max_iters=K) -> macb_fw_run, host numpy buffers"},
        "gpu_launches": counters["kernel_launches"],
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": f"{h.lanczos_kernel_name()}: one launch per eigen-solve (solver CTAs + one Rayleigh-Ritz "
                     "CTA); per Lanczos step one SpMV (8-byte gathers of the published vector, slots in column order), row sums along "
                     "32-row slices and a pipelined three-term update whose reduction overlaps the next SpMV",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "us_per_lanczos_step": lz_us, "lanczos_steps_timed": lz["phases"],
                     "share_of_timed_region": lz["ms"] / (dev_max * 1e3),
                     "algorithmic_bytes_per_step": lz["algo_bytes_per_phase"],
                     "l2": {"achieved": achieved, "peak": l2_peak, "unit": "GB/s", "frac": achieved / l2_peak,
                            "what": "the same algorithmic bytes against the L2 -> SM read bandwidth measured in this run "
                                    "(macb_measure_l2_bandwidth: all SMs streaming a 24 MB L2-resident buffer); the matrix is "
                                    "L2-resident, so this is the bound that binds"},
                     "standalone_spmv": {"kernel": "k_spmv", "achieved": spmv_achieved, "frac": spmv_achieved / peak},
                     "hbm_spmv": hbm_spmv,
                     "note": "achieved = algorithmic bytes (active slots) x steps / CUDA-event time of the kernel launches inside the "
                             "timed region. The 29.6 MB matrix is L2-resident (traffic = DRAM bytes per step of the ncu --set full "
                             "capture under profiles/), every gather moves a 32-byte L2 sector for 8 useful bytes, and pass 1 of a step "
                             "runs at ~70 % of the measured L2 bandwidth in sector traffic: see DESIGN.md section 5"},
    }
    if world == 1 and not args.no_cpu_baseline:
        fixed0, cand0, n0, k0, x00 = (fixed, cand, n, k, x0)
        iters, secs, f_cpu = cpu_fw(fixed0, cand0, n0, k0, x00, args.cpu_baseline_budget, 3)
        line["cpu_baseline"] = {"value": iters / secs, "unit": UNIT, "cores": threads_used(), "kind": "port",
                                "sample": f"{iters} whole FW iterations of the same workload (oracle, scipy ARPACK eigen-solve)"}
        # self-check (untimed): lambda2 of the device on the oracle's own iterates x_t against the oracle's f_t -- same
        # matrix in, same Fiedler value out.  The free-running histories are reported beside it: after iteration 0 they
        # differ at the few-1e-6 level because two solvers that both stop at the reference's residual 1e-8 (nx:243) pick
        # different LP entries inside the eigenvector-noise window (tests/test_gpu_parity.py asserts the window).
        ncmp = min(len(f_cpu), K)
        rel = [abs(mac.evaluate_objective(xt) - ft) / abs(ft) for ft, xt in f_cpu[:ncmp]]
        free = [abs(float(info["f_hist"][i]) - f_cpu[i][0]) / abs(f_cpu[i][0]) for i in range(ncmp)]
        line["parity_check"] = {"max_rel_err": max(rel) if rel else None, "iters_compared": ncmp, "tolerance": PARITY_TOL,
                                "what": "lambda2(L(x_t)) on the device vs the oracle (scipy ARPACK), x_t = the oracle's FW iterates "
                                        "from the same x_init (same matrix in both)",
                                "free_running_max_rel_dev": max(free) if free else None,
                                "free_running_iter0_rel_err": free[0] if free else None}
        parity_failed = bool(rel) and (max(rel) > PARITY_TOL or free[0] > PARITY_TOL)
    print(json.dumps(line), flush=True)
    mac.close()
    if dist is not None:
        dist.destroy_process_group()
    if world == 1 and not args.no_cpu_baseline and parity_failed:
        print(f"bench: PARITY FAILURE: max relative error {max(rel):.3e} > {PARITY_TOL:g}", file=sys.stderr)
        sys.exit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)   # max_iters of the reference's g2o protocol (g2o_experiment.py:319)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=120.0, help="seconds of CPU work for --impl reference")
    ap.add_argument("--cpu-baseline-budget", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ksweep", action="store_true", help="skip the sphere2500 + city10000 budget sweep (config.ksweep)")
    ap.add_argument("--no-hbm-spmv", action="store_true", help="skip the 1.18 GB HBM-bound SpMV point (roofline.hbm_spmv)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
