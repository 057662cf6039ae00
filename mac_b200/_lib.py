"""ctypes binding of libmacb200.so (include/macb200.h).  Thin by design: numpy arrays in, numpy out.

There is NO fallback: if the shared library is missing, or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import warnings

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MACB_LIB", os.path.join(_HERE, "libmacb200.so"))

MACB_OK = 0
MACB_NOT_CONVERGED = 1
T_NAMES = ["assemble", "fiedler", "gradient", "topk", "update", "copy"]


class MacbError(RuntimeError):
    """A libmacb200 call returned a negative status."""


class MacbNotConverged(RuntimeWarning):
    pass


_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


def _sig(fn, argtypes, restype=C.c_int):
    fn.argtypes = argtypes
    fn.restype = restype


def lib():
    """Load the library once.  Raises ImportError with the build hint when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "mac_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    H = C.c_void_p
    _sig(L.macb_create, [C.c_int32, C.c_int64, _ip, _ip, _dp, C.c_int64, _ip, _ip, _dp, C.c_int, C.POINTER(H)])
    _sig(L.macb_destroy, [H])
    _sig(L.macb_last_error, [H], C.c_char_p)
    _sig(L.macb_set_x, [H, _dp, C.c_double])
    _sig(L.macb_get_x, [H, _dp])
    _sig(L.macb_spmv, [H, _dp, _dp])
    _sig(L.macb_lnorm, [H, _dp])
    _sig(L.macb_set_start, [H, _dp])
    _sig(L.macb_fiedler, [H, C.c_double, C.c_int, C.c_int, _dp, _dp, C.POINTER(C.c_int), _dp])
    _sig(L.macb_gradient, [H, _dp])
    _sig(L.macb_evaluate_batch, [H, _dp, C.c_int, C.c_double, C.c_double, C.c_int, _dp, _dp])
    _sig(L.macb_topk, [H, C.c_int64, _dp])
    _sig(L.macb_topk_dense, [C.c_int, _dp, C.c_int64, C.c_int64, _dp])
    L.macb_dense_cache_clear.argtypes = []
    L.macb_dense_cache_clear.restype = None
    _sig(L.macb_round_nearest, [H, _dp, C.c_int64, C.c_int, _dp])
    _sig(L.macb_round_nearest_dense, [C.c_int, _dp, _dp, C.c_int64, C.c_int64, C.c_int, _dp])
    _sig(L.macb_fw_run, [H, C.c_int64, _dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                         _dp, _dp, C.POINTER(C.c_int), _dp, _dp])
    _sig(L.macb_counters, [H, _lp, _lp, _lp, _lp, _dp])
    _sig(L.macb_reset_counters, [H])
    _sig(L.macb_set_profile, [H, C.c_int])
    _sig(L.macb_spmv_bench, [H, C.c_int, C.c_int, _dp, _dp])
    _sig(L.macb_sizes, [H, _lp, _lp, _lp, _lp])
    _sig(L.macb_l2_flush, [H])
    _sig(L.macb_set_bench, [H, C.c_int, C.c_int])
    _sig(L.macb_iter_ms, [H, _dp, C.c_int, C.POINTER(C.c_int)])
    _sig(L.macb_device_sync, [H])
    _sig(L.macb_lanczos_kernel_time, [H, _dp, _lp, _dp])
    _sig(L.macb_lanczos_kernel_name, [H], C.c_char_p)
    _sig(L.macb_lanczos_footprint, [H, C.POINTER(C.c_int32), C.POINTER(C.c_int32)])
    _sig(L.macb_spmv_engine, [H, C.c_int])
    _sig(L.macb_tridiag_smallest, [_dp, _dp, C.c_int, _dp, _dp])
    _sig(L.macb_host_build_pattern, [C.c_int32, C.c_int64, _ip, _ip, C.c_int64, _ip, _ip, _ip, _ip, _ip, _lp])
    _sig(L.macb_host_build_slices, [C.c_int32, _ip, _ip, _ip, C.c_int32, _ip, C.c_int, _ip, _ip, _ip, _ip, _ip, _ip])
    _sig(L.macb_comm_unique_id, [C.c_char_p])
    _sig(L.macb_comm_init, [C.c_int, C.c_int, C.c_char_p, C.c_int, C.POINTER(H)])
    _sig(L.macb_comm_allgather, [H, C.c_void_p, C.c_void_p, C.c_int64])
    _sig(L.macb_comm_destroy, [H])
    _sig(L.macb_comm_last_error, [H], C.c_char_p)
    _sig(L.macb_sweep_owner, [_lp, C.c_int, C.c_int64, C.c_int, _ip])
    _sig(L.macb_sweep, [H, H, _lp, C.c_int, _dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                        C.POINTER(C.c_uint8), _dp, _dp, _dp, _ip])
    _sig(L.macb_device_rr_stats, [H, C.POINTER(C.c_int), _lp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), _dp])
    _sig(L.macb_measure_l2_bandwidth, [C.c_int, C.c_int64, C.c_int, _dp])
    _sig(L.macb_version, [], C.c_char_p)
    _lib = L
    return L


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, typ):
    return a.ctypes.data_as(typ) if a is not None and a.size else None


def measure_l2_bandwidth(device=-1, nbytes=24 << 20, reps=20):
    """GB/s of L2 -> SM reads (all SMs streaming an L2-resident buffer); see include/macb200.h."""
    out = C.c_double()
    rc = lib().macb_measure_l2_bandwidth(int(device), int(nbytes), int(reps), C.byref(out))
    if rc != MACB_OK:
        raise MacbError(f"macb_measure_l2_bandwidth failed ({rc}): {lib().macb_last_error(None).decode()}")
    return out.value


def sweep_owner(budgets, m, nranks):
    """Rank that solves each budget (host-only: macb_sweep_owner)."""
    ks = np.ascontiguousarray(budgets, dtype=np.int64)
    owner = np.zeros(len(ks), dtype=np.int32)
    rc = lib().macb_sweep_owner(_p(ks, _lp), len(ks), int(m), int(nranks), _p(owner, _ip))
    if rc != MACB_OK:
        raise MacbError(f"macb_sweep_owner failed ({rc})")
    return owner


class Comm:
    """NCCL communicator of the K-sweep farm (macb_comm_*): one per process, created from a 128-byte unique id."""

    def __init__(self, nranks, rank, unique_id, device=-1):
        self._L = lib()
        self._c = C.c_void_p()
        self.nranks, self.rank = int(nranks), int(rank)
        rc = self._L.macb_comm_init(self.nranks, self.rank, unique_id, int(device), C.byref(self._c))
        if rc != MACB_OK:
            raise MacbError(f"macb_comm_init failed ({rc}): {self._L.macb_last_error(None).decode()}")

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        rc = lib().macb_comm_unique_id(buf)
        if rc != MACB_OK:
            raise MacbError(f"macb_comm_unique_id failed ({rc}): {lib().macb_last_error(None).decode()}")
        return buf.raw

    def allgather(self, arr):
        arr = np.ascontiguousarray(arr)
        out = np.empty((self.nranks,) + arr.shape, dtype=arr.dtype)
        rc = self._L.macb_comm_allgather(self._c, arr.ctypes.data, out.ctypes.data, arr.nbytes)
        if rc != MACB_OK:
            raise MacbError(f"macb_comm_allgather failed ({rc}): {self._L.macb_comm_last_error(self._c).decode()}")
        return out

    def close(self):
        if self._c:
            self._L.macb_comm_destroy(self._c)
            self._c = C.c_void_p()


class Handle:
    """One graph resident on one GPU (macb_handle).  Not thread-safe."""

    def __init__(self, n, fi, fj, fw, ci, cj, ckappa, device=-1):
        L = lib()
        self._L = L
        self._h = C.c_void_p()
        fi, fj, fw = _i32(fi), _i32(fj), _f64(fw)
        ci, cj, ck = _i32(ci), _i32(cj), _f64(ckappa)
        assert len(fi) == len(fj) == len(fw) and len(ci) == len(cj) == len(ck)
        self.n, self.nf, self.m = int(n), len(fi), len(ci)
        rc = L.macb_create(self.n, self.nf, _p(fi, _ip), _p(fj, _ip), _p(fw, _dp), self.m, _p(ci, _ip), _p(cj, _ip),
                           _p(ck, _dp), int(device), C.byref(self._h))
        if rc != MACB_OK:
            raise MacbError(f"macb_create failed ({rc}): {L.macb_last_error(None).decode()}")

    # -- plumbing
    def _check(self, rc, what):
        if rc == MACB_OK:
            return
        msg = self._L.macb_last_error(self._h).decode()
        if rc == MACB_NOT_CONVERGED:
            warnings.warn(f"{what}: {msg}", MacbNotConverged, stacklevel=3)
            return
        raise MacbError(f"{what} failed ({rc}): {msg}")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.macb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- Laplacian
    def set_x(self, x, min_sel_tol=1e-10):
        x = _f64(x)
        assert x.shape == (self.m,)
        self._check(self._L.macb_set_x(self._h, _p(x, _dp), float(min_sel_tol)), "macb_set_x")

    def get_x(self):
        x = np.empty(self.m)
        self._check(self._L.macb_get_x(self._h, _p(x, _dp)), "macb_get_x")
        return x

    def spmv(self, v):
        v = _f64(v)
        assert v.shape == (self.n,)
        y = np.empty(self.n)
        self._check(self._L.macb_spmv(self._h, _p(v, _dp), _p(y, _dp)), "macb_spmv")
        return y

    def lnorm(self):
        out = C.c_double()
        self._check(self._L.macb_lnorm(self._h, C.byref(out)), "macb_lnorm")
        return out.value

    # -- eigen-solve
    def set_start(self, x0):
        if x0 is None:
            self._check(self._L.macb_set_start(self._h, None), "macb_set_start")
            return
        x0 = _f64(x0)
        assert x0.shape == (self.n,)
        self._check(self._L.macb_set_start(self._h, _p(x0, _dp)), "macb_set_start")

    def fiedler(self, tol=1e-8, max_steps=0, warm=False, want_vector=True):
        lam, res, steps = C.c_double(), C.c_double(), C.c_int()
        v = np.empty(self.n) if want_vector else None
        rc = self._L.macb_fiedler(self._h, float(tol), int(max_steps), int(bool(warm)), C.byref(lam),
                                  _p(v, _dp) if want_vector else None, C.byref(steps), C.byref(res))
        self._check(rc, "macb_fiedler")
        return lam.value, v, {"steps": steps.value, "resid": res.value, "converged": rc == MACB_OK}

    # -- gradient / LP
    def gradient(self, want=True):
        g = np.empty(self.m) if want else None
        self._check(self._L.macb_gradient(self._h, _p(g, _dp) if want else None), "macb_gradient")
        return g

    def topk(self, k, want=True):
        s = np.empty(self.m) if want else None
        self._check(self._L.macb_topk(self._h, int(k), _p(s, _dp) if want else None), "macb_topk")
        return s

    def round_nearest(self, w, k, decimals=10):
        w = _f64(w)
        assert w.shape == (self.m,)
        out = np.empty(self.m)
        self._check(self._L.macb_round_nearest(self._h, _p(w, _dp), int(k), int(decimals), _p(out, _dp)), "macb_round_nearest")
        return out

    # -- whole loop
    def fw_run(self, k, x_init, max_iters, rel_gap_tol, grad_norm_tol, fiedler_tol=1e-8, min_sel_tol=1e-10,
               fiedler_max_steps=0, warm=False):
        x_init = _f64(x_init)
        assert x_init.shape == (self.m,)
        w = np.empty(self.m)
        u = C.c_double()
        iters = C.c_int()
        fh = np.full(max(1, max_iters), np.nan)
        uh = np.full(max(1, max_iters), np.nan)
        rc = self._L.macb_fw_run(self._h, int(k), _p(x_init, _dp), int(max_iters), float(rel_gap_tol), float(grad_norm_tol),
                                 float(fiedler_tol), float(min_sel_tol), int(fiedler_max_steps), int(bool(warm)),
                                 _p(w, _dp), C.byref(u), C.byref(iters), _p(fh, _dp), _p(uh, _dp))
        self._check(rc, "macb_fw_run")
        it = iters.value
        return w, u.value, {"iters": it, "f_hist": fh[:it].copy(), "u_hist": uh[:it].copy()}

    def evaluate_batch(self, xs, tol=1e-8, min_sel_tol=1e-10, max_steps=0):
        """lambda2(L(x_b)) for every row of xs, one host synchronisation (macb_evaluate_batch)."""
        xs = np.ascontiguousarray(xs, dtype=np.float64).reshape(-1, self.m)
        lam = np.zeros(len(xs))
        res = np.zeros(len(xs))
        rc = self._L.macb_evaluate_batch(self._h, _p(xs, _dp), len(xs), float(tol), float(min_sel_tol), int(max_steps), _p(lam, _dp),
                                         _p(res, _dp))
        self._check(rc, "macb_evaluate_batch")
        return lam, res

    def sweep(self, comm, budgets, x_inits, max_iters=20, rel_gap_tol=1e-4, grad_norm_tol=1e-8, fiedler_tol=1e-8, min_sel_tol=1e-10,
              fiedler_max_steps=0, want_w=True):
        """macb_sweep: the g2o budget sweep, this rank's share solved here, results gathered over `comm` (None: single process)."""
        ks = np.ascontiguousarray(budgets, dtype=np.int64)
        nk = len(ks)
        x_inits = np.ascontiguousarray(x_inits, dtype=np.float64).reshape(nk, self.m)
        rounded = np.zeros((nk, self.m), dtype=np.uint8)
        w = np.zeros((nk, self.m)) if want_w else None
        u, lam = np.zeros(nk), np.zeros(nk)
        iters = np.zeros(nk, dtype=np.int32)
        rc = self._L.macb_sweep(self._h, comm._c if comm is not None else None, _p(ks, _lp), nk, _p(x_inits, _dp), int(max_iters),
                                float(rel_gap_tol), float(grad_norm_tol), float(fiedler_tol), float(min_sel_tol), int(fiedler_max_steps),
                                rounded.ctypes.data_as(C.POINTER(C.c_uint8)), _p(w, _dp) if want_w else None, _p(u, _dp), _p(lam, _dp),
                                _p(iters, _ip))
        self._check(rc, "macb_sweep")
        return rounded, w, u, lam, iters

    # -- measurement
    def counters(self):
        a, b, c_, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        ms = (C.c_double * len(T_NAMES))()
        self._L.macb_counters(self._h, C.byref(a), C.byref(b), C.byref(c_), C.byref(d), ms)
        return {"kernel_launches": a.value, "spmv_launches": b.value, "lanczos_steps": c_.value,
                "fiedler_solves": d.value, "phase_ms": dict(zip(T_NAMES, list(ms)))}

    def device_rr_stats(self):
        en, st, k, ch = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        fb = C.c_int64()
        tet = (C.c_double * 13)()
        self._L.macb_device_rr_stats(self._h, C.byref(en), C.byref(fb), C.byref(st), C.byref(k), C.byref(ch), tet)
        return {"enabled": bool(en.value), "fallbacks": fb.value, "last_status": st.value, "last_k": k.value, "last_checks": ch.value,
                "last_theta": tet[0], "last_est": tet[1], "last_target": tet[2], "cycles_waiting": tet[3], "cycles_computing": tet[4],
                "lag_steps_at_decision": int(tet[5]), "multisection_rounds": int(tet[6]),
                "cycles_by_stage": dict(zip(["fetch", "bounds", "warm_bracket", "multisection", "twisted", "rest"], [int(v) for v in tet[7:13]]))}

    def reset_counters(self):
        self._L.macb_reset_counters(self._h)

    def set_profile(self, on):
        self._L.macb_set_profile(self._h, int(bool(on)))

    def spmv_bench(self, reps=200, flush_l2=False):
        ms, by = C.c_double(), C.c_double()
        self._check(self._L.macb_spmv_bench(self._h, int(reps), int(bool(flush_l2)), C.byref(ms), C.byref(by)),
                    "macb_spmv_bench")
        return ms.value, by.value

    def sizes(self):
        a, b, c_, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        self._L.macb_sizes(self._h, C.byref(a), C.byref(b), C.byref(c_), C.byref(d))
        return {"n": a.value, "m": b.value, "nnz_union": c_.value, "nnz_active": d.value}

    def l2_flush(self):
        self._check(self._L.macb_l2_flush(self._h), "macb_l2_flush")

    def set_bench(self, time_iters=True, flush_l2_between_iters=True):
        self._L.macb_set_bench(self._h, int(bool(time_iters)), int(bool(flush_l2_between_iters)))

    def iter_ms(self):
        cnt = C.c_int()
        self._L.macb_iter_ms(self._h, None, 0, C.byref(cnt))
        ms = np.zeros(max(cnt.value, 1))
        self._L.macb_iter_ms(self._h, _p(ms, _dp), cnt.value, C.byref(cnt))
        return ms[:cnt.value]

    def lanczos_kernel_time(self):
        ms, by, ph = C.c_double(), C.c_double(), C.c_int64()
        self._L.macb_lanczos_kernel_time(self._h, C.byref(ms), C.byref(ph), C.byref(by))
        return {"ms": ms.value, "phases": ph.value, "algo_bytes_per_phase": by.value}

    def spmv_engine(self, engine):
        """0: CSR kernel (default); 1: chunked jagged-diagonal kernel (built on first use)."""
        self._check(self._L.macb_spmv_engine(self._h, int(engine)), "macb_spmv_engine")

    def lanczos_kernel_name(self):
        return self._L.macb_lanczos_kernel_name(self._h).decode()

    def lanczos_footprint(self):
        """(CTAs one eigen-solve launch occupies, SMs of the device).  Builds the engine if no solve has done so yet."""
        a, b = C.c_int32(), C.c_int32()
        self._check(self._L.macb_lanczos_footprint(self._h, C.byref(a), C.byref(b)), "macb_lanczos_footprint")
        return a.value, b.value

    def device_sync(self):
        self._check(self._L.macb_device_sync(self._h), "macb_device_sync")


def dense_cache_clear():
    """Release the handles `topk_dense` / `round_nearest_dense` keep per (device, m)."""
    lib().macb_dense_cache_clear()


def topk_dense(g, k, device=-1):
    """solve_subset_box_lp on the device for an arbitrary host vector."""
    L = lib()
    g = _f64(g)
    s = np.empty_like(g)
    rc = L.macb_topk_dense(int(device), _p(g, _dp), g.size, int(k), _p(s, _dp))
    if rc != MACB_OK:
        raise MacbError(f"macb_topk_dense failed ({rc}): {L.macb_last_error(None).decode()}")
    return s


def round_nearest_dense(w, weights, k, decimals, device=-1):
    """Tie-broken nearest rounding (rounding.py:30-42) on the device for arbitrary host vectors."""
    L = lib()
    w, weights = _f64(w), _f64(weights)
    assert w.shape == weights.shape
    out = np.empty_like(w)
    rc = L.macb_round_nearest_dense(int(device), _p(w, _dp), _p(weights, _dp), w.size, int(k), int(decimals), _p(out, _dp))
    if rc != MACB_OK:
        raise MacbError(f"macb_round_nearest_dense failed ({rc}): {L.macb_last_error(None).decode()}")
    return out


def tridiag_smallest(a, b):
    """Host helper (no GPU): smallest eigenpair of tridiag(a; b[1:])."""
    L = lib()
    a, b = _f64(a), _f64(b)
    k = a.size
    assert b.size >= k
    th = C.c_double()
    s = np.empty(k)
    rc = L.macb_tridiag_smallest(_p(a, _dp), _p(b, _dp), k, C.byref(th), _p(s, _dp))
    if rc != MACB_OK:
        raise MacbError("macb_tridiag_smallest failed")
    return th.value, s


def host_build_pattern(n, fi, fj, ci, cj):
    """Host helper (no GPU): the union CSR pattern macb_create builds."""
    L = lib()
    fi, fj, ci, cj = _i32(fi), _i32(fj), _i32(ci), _i32(cj)
    cap = 2 * (len(fi) + len(ci))
    rp = np.empty(n + 1, dtype=np.int32)
    col = np.empty(max(cap, 1), dtype=np.int32)
    eid = np.empty(max(cap, 1), dtype=np.int32)
    nnz = C.c_int64()
    rc = L.macb_host_build_pattern(n, len(fi), _p(fi, _ip), _p(fj, _ip), len(ci), _p(ci, _ip), _p(cj, _ip),
                                   _p(rp, _ip), _p(col, _ip), _p(eid, _ip), C.byref(nnz))
    if rc != MACB_OK:
        raise MacbError(f"macb_host_build_pattern failed ({rc}): {L.macb_last_error(None).decode()}")
    return rp, col[:nnz.value], eid[:nnz.value]


SLICE_STRIDE = 33   # kLzSlice
SLICE_TAB = 64      # kLzSliceTab


def host_build_slices(n, rp, col, eid, row_start, bankfit=True):
    """Host helper (no GPU): the sliced layout of k_lanczos_pipe for a given CTA partition `row_start` (len ncta + 1).
    Returns (jrow, jlen, jcol, jeid, jw[ncta, 64], positions[ncta])."""
    L = lib()
    rp, col, eid, row_start = _i32(rp), _i32(col), _i32(eid), _i32(row_start)
    ncta = len(row_start) - 1
    nnz = len(col)
    jrow = np.empty(max(n, 1), dtype=np.int32)
    jlen = np.empty(max(n, 1), dtype=np.int32)
    jcol = np.empty(max(nnz, 1), dtype=np.int32)
    jeid = np.empty(max(nnz, 1), dtype=np.int32)
    jw = np.zeros(max(ncta * SLICE_TAB, 1), dtype=np.int32)
    positions = np.zeros(max(ncta, 1), dtype=np.int32)
    rc = L.macb_host_build_slices(n, _p(rp, _ip), _p(col, _ip), _p(eid, _ip), ncta, _p(row_start, _ip), int(bool(bankfit)),
                                  _p(jrow, _ip), _p(jlen, _ip), _p(jcol, _ip), _p(jeid, _ip), _p(jw, _ip), _p(positions, _ip))
    if rc != MACB_OK:
        raise MacbError(f"macb_host_build_slices failed ({rc})")
    return jrow[:n], jlen[:n], jcol[:nnz], jeid[:nnz], jw.reshape(ncta, SLICE_TAB), positions[:ncta]
