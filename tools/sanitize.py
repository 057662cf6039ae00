"""compute-sanitizer target: one small problem through every kernel family (vec / jds / slots / small engines,
assembly, SpMV engines, gradient, top-k, rounding).  Run: compute-sanitizer --tool memcheck python tools/sanitize.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mac_b200 import synth
from mac_b200.solvers import MAC

def run(tag, n, m, k):
    fixed, cand, n = synth.chain_plus_random(n, m, seed=5, weighted=True)
    x0 = synth.first_k_init(m, k)
    mac = MAC(fixed, cand, n)
    r, w, u = mac.solve(k, x0, max_iters=3)
    h = mac._h
    h.set_x(w)
    v = np.random.default_rng(0).normal(size=n)
    y0 = h.spmv(v); h.spmv_engine(1); y1 = h.spmv(v); h.spmv_engine(0)
    print(tag, h.lanczos_kernel_name(), "u", u, "selected", int(r.sum()), "spmv engines agree", float(np.abs(y0 - y1).max()))
    mac.close()

run("small", 600, 3000, 600)            # k_lanczos_small
run("multi-CTA", 6000, 40000, 8000)     # k_lanczos_pipe, several CTAs
for env in ({"MACB_NO_VEC": "1"}, {"MACB_NO_JDS": "1"}, {"MACB_PERSIST_V": "1"}):
    os.environ.update(env)
    run(str(env), 6000, 40000, 8000)
    for k_ in env: del os.environ[k_]
