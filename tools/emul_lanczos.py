"""CPU emulation of the device Lanczos driver (api.cu run_fiedler) for algorithm debugging.  Scratch tool."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mac_b200 import _lib
from mac_b200.utils.fiedler import seeded_start

S = 32
def sinv(b): return 1.0 / b if b > 1e-290 else 0.0

def fiedler(L, tol=1e-8, max_steps=20000, x0=None, verbose=False, cap=65536):
    n = L.shape[0]
    lnorm = abs(L).sum(axis=1).max(); sqrtn = np.sqrt(n); brk = max(1e-12 * lnorm, 0.25 * tol * lnorm / np.sqrt(n))
    u = (seeded_start(n)[:, 0] if x0 is None else x0).copy(); u -= u.mean()
    total = 0
    for restart in range(64):
        U = [u]; beta = [np.linalg.norm(u)]; alpha = []; usum = [0.0]
        k_done = 0; k_limit = min(cap, max_steps - total, n - 1); invariant = False; theta_prev = np.inf
        while True:
            batch = S if k_done < 4 * S else min(16 * S, ((k_done // 4) // S) * S)
            batch = min(batch, k_limit - k_done)
            for _ in range(batch):
                j = len(alpha)
                binv = sinv(beta[j]); bpinv = sinv(beta[j - 1]) if j > 0 else 0.0
                y = (L @ U[j]) * binv
                a = (U[j] * binv) @ y; alpha.append(a)
                ca = a * binv; cb = beta[j] * bpinv
                c = (y.sum() - ca * usum[j] - (cb * usum[j - 1] if j > 0 else 0.0)) / n
                w = y - ca * U[j] - (cb * U[j - 1] if j > 0 else 0.0) - c
                U.append(w); beta.append(np.sqrt(w @ w)); usum.append(w.sum())
            k_done += batch; total += batch
            k = k_done
            for j in range(1, k_done + 1):
                if not (beta[j] > brk):
                    k = j; invariant = True; break
            th, s = _lib.tridiag_smallest(np.array(alpha[:k]), np.array(beta[:k]))
            est = abs(beta[k]) * abs(s[k - 1])
            exhausted = invariant or k_done >= k_limit
            if verbose: print("k", k, "theta", th, "est", est * sqrtn / lnorm, "inv", invariant, "beta_k", beta[k])
            if est * sqrtn < tol * lnorm or exhausted:
                yv = sum((s[t] / beta[t]) * U[t] for t in range(k)); yv -= yv.mean(); yv /= np.linalg.norm(yv)
                Ly = L @ yv; lam = (yv @ Ly) / (yv @ yv); res = np.abs(Ly - lam * yv).sum() / (np.linalg.norm(yv) * lnorm)
                if verbose: print("   finalize lam", lam, "res", res)
                if res < tol: return lam, yv, total, True
                if exhausted: break
        u = yv
        if total >= max_steps or invariant: break
    return lam, yv, total, False

if __name__ == "__main__":
    from mac_b200 import synth
    from oracle import mac_oracle as orc
    fixed, cand, n = synth.petersen_split(); o = orc.OracleMAC(fixed, cand, n)
    for k in (1, 2, 4):
        x = synth.first_k_init(6, k)
        for i in range(100):
            L = o.laplacian(x)
            lam, v, steps, ok = fiedler(L)
            if not ok:
                print("FAIL k", k, "iter", i, "x", x.tolist(), lam, steps, np.linalg.eigvalsh(L.toarray())[:4]); fiedler(L, verbose=True); break
            g = o.gradient(v); s = orc.solve_subset_box_lp(g, k)
            x = x + 2.0 / (i + 2.0) * (s - x)
        else:
            print("k", k, "ok")
