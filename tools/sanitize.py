"""compute-sanitizer target: one small problem through every kernel family (pipe / slots / persist / small engines,
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
for env in ({"MACB_NO_PIPE": "1"}, {"MACB_NO_JDS": "1"}, {"MACB_PERSIST_V": "1"}, {"MACB_HOST_RR": "1"}):
    os.environ.update(env)
    run(str(env), 6000, 40000, 8000)
    for k_ in env: del os.environ[k_]

# array-only entry points (cached handles), a top-k large enough for the two-pass select + direct ranking, batched evaluation
from mac_b200.optimization.constraints import solve_subset_box_lp
from mac_b200 import _lib
rng = np.random.default_rng(1)
g = rng.random(300000) ** 3
for k in (1, 60000, 299999):
    s_ = solve_subset_box_lp(g, k)
    assert s_.sum() == k
g[::3] = 0.25   # 100 000 exact ties at the k-th value
assert solve_subset_box_lp(g, 150000).sum() == 150000
_lib.dense_cache_clear()
fixed, cand, n = synth.chain_plus_random(6000, 40000, seed=5, weighted=True)
mac = MAC(fixed, cand, n)
xs = np.stack([synth.first_k_init(40000, k) for k in (4000, 8000, 12000)])
print("batched lambda2", mac.evaluate_objectives(xs))
mac.close()
