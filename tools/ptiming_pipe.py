"""Per-phase clock64 breakdown of k_lanczos_pipe (needs `make timing`; run with MACB_LIB=mac_b200/libmacb200_timing.so).
Stamps per phase and CTA: t_start, t_p1 (after the pass-1 barrier), t_rows (own row sum done, thread 0), poll start, poll end,
end of the update, record push done (last warp)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mac_b200 import synth, _lib
from mac_b200.solvers import MAC
which = sys.argv[1] if len(sys.argv) > 1 else "dense"
if which in ("H", "dense"):
    fixed, cand, n, k, x0 = synth.headline(m=int(os.environ.get("MACB_PT_M", 1000000)))
    if which == "dense":
        x0 = np.full(len(x0), 0.2)
else:
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", f"g2o_{which}.npz"))
    from mac_b200.g2o import split_edges
    fixed, cand = split_edges(z["i"], z["j"], z["kappa"]); n = int(z["n"]); x0 = np.ones(len(cand[0])) * 0.3
mac = MAC(fixed, cand, n)
lam, v = mac.fiedler_pair(x0)
print(which, "lambda2", lam, mac.last_info, mac._h.sizes(), mac._h.lanczos_kernel_name())
L = _lib.lib()
ncta = C.c_int()
L.macb_debug_ptiming(mac._h._h, None, C.byref(ncta))
raw = np.zeros(64 * ncta.value * 9, dtype=np.int64)
L.macb_debug_ptiming(mac._h._h, raw.ctypes.data_as(C.c_void_p), C.byref(ncta))
t = raw[:64 * ncta.value * 8].reshape(64, ncta.value, 8)[4:60].astype(np.float64)
def st(x): return "mean %.0f  min-cta %.0f  max-cta %.0f" % (x.mean(), x.min(axis=1).mean(), x.max(axis=1).mean())
print("ncta", ncta.value)
print("pass 1 (gathers + barrier A)        ", st(t[:, :, 1] - t[:, :, 0]))
print("pass 2 (row sum, thread 0)          ", st(t[:, :, 2] - t[:, :, 1]))
print("update: barrier B -> un, zn computed  ", st(t[:, :, 3] - t[:, :, 6]))
print("update: -> stores issued            ", st(t[:, :, 4] - t[:, :, 3]))
print("update: -> warp sums done           ", st(t[:, :, 5] - t[:, :, 4]))
print("rows done -> barrier B passed (t0)  ", st(t[:, :, 6] - t[:, :, 2]))
print("update, stores, block sums (t0)     ", st(t[:, :, 7] - t[:, :, 6]))
print("end -> next start (barrier C)       ", st(t[1:, :, 0] - t[:-1, :, 7]))
print("whole phase (start -> start)        ", st(np.diff(t[:, :, 0], axis=0)))
