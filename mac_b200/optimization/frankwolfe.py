"""Generic Frank-Wolfe driver for user-supplied `problem` / `solve_lp` callables -- the fine seam of the drop-in
(reference: mac/optimization/frankwolfe.py:10-79, same signature and return value).  `MAC.solve` does not come through
here: it runs the whole loop on the device (`macb_fw_run`, csrc/api.cu), which implements exactly these rules:

    s_t  = argmax_{s in C} <g_t, s>                      (solve_lp)
    u    = min(u, f_t + <g_t, s_t - x_t>)                (dual upper bound, frankwolfe.py:62)
    stop   if ||g_t||_2 < grad_norm_tol                  (frankwolfe.py:65)
    stop   if u - f_t < relative_duality_gap_tol |f_t|   (frankwolfe.py:71)
    x_{t+1} = x_t + gamma_t (s_t - x_t),  gamma_t = 2 / (t + 2) unless `stepsize` is given   (frankwolfe.py:7-8,76)
"""
import numpy as np


def naive_stepsize(k):
    """gamma_k = 2 / (k + 2)  (frankwolfe.py:7-8)."""
    return 2.0 / (k + 2.0)


def frank_wolfe(initial, problem, solve_lp, stepsize=None, maxiter=50, relative_duality_gap_tol=1e-5,
                grad_norm_tol=1e-10, verbose=False):
    """Maximise a concave function over a compact convex set.  Returns (x, u): the last iterate and the best dual
    upper bound seen.  `problem(x) -> (f, grad)`, `solve_lp(grad) -> vertex`, `stepsize(x, grad, s, k) -> gamma`."""
    step = stepsize if stepsize is not None else (lambda _x, _g, _s, k: naive_stepsize(k))
    say = print if verbose else (lambda *_: None)
    x, upper = initial, np.inf
    for t in range(maxiter):
        value, grad = problem(x)
        vertex = solve_lp(grad)
        direction = vertex - x
        upper = min(upper, value + grad @ direction)
        if np.linalg.norm(grad) < grad_norm_tol:
            say(f"iteration {t}: gradient norm below {grad_norm_tol:g}, stationary point")
            break
        if upper - value < relative_duality_gap_tol * abs(value):
            say(f"iteration {t}: duality gap {upper - value:.3e} within tolerance")
            break
        x = x + step(x, grad, vertex, t) * direction
    else:
        say(f"stopped after maxiter = {maxiter} iterations")
    return x, float(upper) if np.isfinite(upper) else upper
