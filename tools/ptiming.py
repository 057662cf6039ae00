"""Per-phase timing breakdown of the persistent Lanczos kernel (needs `make timing`; run with
MACB_LIB=mac_b200/libmacb200_timing.so).  Scratch tool."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mac_b200 import synth, _lib
from mac_b200.solvers import MAC
which = sys.argv[1] if len(sys.argv) > 1 else "H"
print("stream", os.environ.get("MACB_PERSIST_STREAM", "1"))
if which == "H":
    fixed, cand, n, k, x0 = synth.headline()
elif which == "dense":
    fixed, cand, n, k, x0 = synth.headline(); x0 = np.full(len(x0), 0.2)
else:
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", f"g2o_{which}.npz"))
    from mac_b200.g2o import split_edges
    fixed, cand = split_edges(z["i"], z["j"], z["kappa"]); n = int(z["n"]); x0 = np.ones(len(cand[0])) * 0.3
mac = MAC(fixed, cand, n)
lam, v = mac.fiedler_pair(x0)
print(which, "lambda2", lam, mac.last_info, mac._h.sizes())
L = _lib.lib()
ncta = C.c_int()
L.macb_debug_ptiming(mac._h._h, None, C.byref(ncta))
raw = np.zeros(64 * ncta.value * 9, dtype=np.int64)
L.macb_debug_ptiming(mac._h._h, raw.ctypes.data_as(C.c_void_p), C.byref(ncta))
buf = raw[:64 * ncta.value * 4].reshape(64, ncta.value, 4)
p1 = raw[64 * ncta.value * 4:64 * ncta.value * 5].reshape(64, ncta.value)
ex = raw[64 * ncta.value * 5:].reshape(64, ncta.value, 4)[2:33].astype(np.float64)
t = buf[2:33].astype(np.float64)   # phases of the last launch (skip first two)
rows = t[:, :, 1] - t[:, :, 0]; bar = t[:, :, 2] - t[:, :, 1]; red = t[:, :, 3] - t[:, :, 2]
print("ncta", ncta.value)
print("cycles per phase (mean over phases): rows mean %.0f max %.0f min %.0f | barrier wait mean %.0f min %.0f | reduce mean %.0f" % (
    rows.mean(), rows.max(axis=1).mean(), rows.min(axis=1).mean(), bar.mean(), bar.min(axis=1).mean(), red.mean()))
if p1[2:33].any():
    g = (p1[2:33] - buf[2:33, :, 0]).astype(np.float64)
    print("slots kernel: pass 1 (gathers) mean %.0f max %.0f | pass 2 + block reduce mean %.0f" % (g.mean(), g.max(axis=1).mean(), (rows - g).mean()))
if os.environ.get("PT_DETAIL"):
    t0 = t[:, :, 0]; tc = t[:, :, 3]
    print("vec: slots that had to be gathered again per CTA and step (t[3] in the vec kernel): mean %.0f max %.0f" % (tc.mean(), tc.max(axis=1).mean()))
    print("jds: start->coef ready mean %.0f max %.0f min %.0f" % ((tc - t0).mean(), (tc - t0).max(axis=1).mean(), (tc - t0).min(axis=1).mean()))
    rr = rows.mean(axis=0); order = np.argsort(rr)
    print("rows per CTA: slowest", [(int(b), int(rr[b])) for b in order[-6:]], "fastest", [(int(b), int(rr[b])) for b in order[:4]])
    g1 = (p1[2:33] - buf[2:33, :, 0]).mean(axis=0)
    print("pass1 per CTA: slowest", [(int(b), int(g1[b])) for b in np.argsort(g1)[-6:]])
    if ex.any():
        tr = t[:, :, 1]
        def st(x): return "mean %.0f min %.0f max %.0f" % (x.mean(), x.min(axis=1).mean(), x.max(axis=1).mean())
        print("barrier split: t_rows->record stored", st(ex[:, :, 0] - tr), "| red.release + first poll", st(ex[:, :, 1] - ex[:, :, 0]),
              "| polling", st(ex[:, :, 2] - ex[:, :, 1]), "| record fetch + reduce", st(ex[:, :, 3] - ex[:, :, 2]), "| final sync", st(t[:, :, 2] - ex[:, :, 3]))
    # skew of the step start across CTAs
    print("start skew (max-min over CTAs of t_start, mean over phases): %.0f" % (t0.max(axis=1) - t0.min(axis=1)).mean())
print("total per phase (cta 0):", np.diff(t[:, 0, 0]).mean())
