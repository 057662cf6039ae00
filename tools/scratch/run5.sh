python - <<'PY'
import numpy as np, time
from mac_b200 import synth
from mac_b200.solvers import MAC
fixed, cand, n, k, x0 = synth.headline()
mac = MAC(fixed, cand, n)
lam, v = mac.fiedler_pair(x0)
print("single solve", lam, mac.last_info, mac._h.device_rr_stats())
t=time.time(); w,u,info = mac.frank_wolfe(k, x0, 10, 0.0, 0.0); print("10 iters", time.time()-t)
print("rr stats", mac._h.device_rr_stats(), mac._h.counters())
PY
