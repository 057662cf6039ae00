"""Generic Frank-Wolfe driver for user-supplied `problem` / `solve_lp` callables
(mac/optimization/frankwolfe.py:10-79) -- the fine seam of the drop-in.  `MAC.solve` does not
come through here: it runs the whole loop on the device (`macb_fw_run`)."""
import numpy as np


def naive_stepsize(k):
    return 2.0 / (k + 2.0)


def frank_wolfe(initial, problem, solve_lp, stepsize=None, maxiter=50, relative_duality_gap_tol=1e-5,
                grad_norm_tol=1e-10, verbose=False):
    if stepsize is None:
        stepsize = lambda x, g, s, k: naive_stepsize(k)  # noqa: E731
    x = initial
    u = float("inf")
    for i in range(maxiter):
        f, gradf = problem(x)
        s = solve_lp(gradf)
        u = min(u, f + gradf @ (s - x))
        if np.linalg.norm(gradf) < grad_norm_tol:
            if verbose:
                print("Gradient norm is approximately 0. Found optimal solution")
            return x, u
        if (u - f) < relative_duality_gap_tol * abs(f):
            if verbose:
                print("Duality gap tolerance reached, found optimal solution")
            return x, u
        x = x + stepsize(x, gradf, s, i) * (s - x)
    if verbose:
        print("Reached maximum number of iterations, returning best solution")
    return x, u
