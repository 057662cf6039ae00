// Host-side Rayleigh-Ritz for the Lanczos driver: smallest eigenpair of a symmetric tridiagonal.
#pragma once
#include <vector>

namespace macb {

// T_k: diagonal a[0..k), off-diagonal b[1..k) (b[i] couples i-1 and i; b[0] is ignored).
// Returns the smallest eigenvalue; `hint_hi`, if finite, is a known upper bound on it (Cauchy
// interlacing: the smallest Ritz value never increases as k grows), which shortens bisection.
// `hint_delta` (> 0, or <= 0 for none): how far below hint_hi the value is expected to lie (e.g. twice
// the previous decrease); it seeds the downward bracket search so that a nearly converged Ritz value
// costs ~15 Sturm passes instead of ~60.
double tridiag_smallest_value(const double* a, const double* b, int k, double hint_hi, double hint_delta = -1.0);

// Eigenvector s[0..k) (unit 2-norm) of T_k for the eigenvalue `theta` by twisted factorisation
// followed by one step of Rayleigh-quotient-free inverse refinement.  Returns ||T s - theta s||_2.
double tridiag_vector(const double* a, const double* b, int k, double theta, double* s);

}  // namespace macb
