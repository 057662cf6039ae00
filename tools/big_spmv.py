"""HBM-bound SpMV points (SURVEY hard part 2): matrices far larger than the 126 MB L2.
  local : chain + candidates with |i - j| <= band (pose-graph-like locality; gathers mostly hit cache)
  random: chain + uniformly random candidates (expander; every gather is a random 32-byte sector)
Prints algorithmic GB/s (SURVEY 8d bytes / CUDA-event time) and the fraction of the measured HBM peak."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mac_b200 import _lib
n = int(os.environ.get("BIG_N", 4_000_000)); m = int(os.environ.get("BIG_M", 40_000_000)); band = int(os.environ.get("BIG_BAND", 2000))
peak = 6445.6
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p): peak = float(json.load(open(p))["hbm_gbs"])
rng = np.random.default_rng(0)
fi = np.arange(n - 1, dtype=np.int32); fj = fi + 1; fw = np.ones(n - 1)
out = {}
for kind in ("local", "random"):
    t0 = time.time()
    a = rng.integers(0, n, size=m, dtype=np.int64)
    if kind == "local":
        b = a + rng.integers(2, band, size=m, dtype=np.int64); b = np.where(b >= n, a - (b - a), b)
    else:
        b = rng.integers(0, n, size=m, dtype=np.int64)
    ok = np.abs(a - b) > 1
    ci = a[ok].astype(np.int32); cj = b[ok].astype(np.int32)
    h = _lib.Handle(n, fi, fj, fw, ci, cj, np.ones(len(ci)))
    h.set_x(np.ones(len(ci)))
    t1 = time.time()
    s = h.sizes()
    for engine, name in ((0, "k_spmv (CSR, 8 lanes per row)"), (1, "k_spmv_jds (chunked jagged-diagonal, column-sorted slots)")):
        t2 = time.time(); h.spmv_engine(engine); t3 = time.time()
        ms, by = h.spmv_bench(20, False); ms_f, _ = h.spmv_bench(10, True)
        out[kind, engine] = {"kernel": name, "n": n, "nnz_offdiag": s["nnz_union"], "algorithmic_GB": by / 1e9, "ms": ms, "GBs": by / ms / 1e6,
                             "frac_of_hbm_peak": by / ms / 1e6 / peak, "ms_l2_flushed": ms_f, "GBs_l2_flushed": by / ms_f / 1e6,
                             "build_s": t1 - t0, "engine_build_s": t3 - t2}
        print(kind, json.dumps(out[kind, engine]), flush=True)
    if os.environ.get("BIG_HOLD"):   # keep the last engine selected and spin a few launches for ncu
        h.spmv_bench(3, False)
    h.close()
