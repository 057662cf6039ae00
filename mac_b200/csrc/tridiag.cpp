// Smallest eigenpair of a symmetric tridiagonal matrix (host, double precision).
//
// Role on the hot path: the Lanczos iteration (lanczos.cuh) reduces L(w) restricted to the
// complement of the all-ones vector to a tridiagonal T_k on the device; the Rayleigh-Ritz step --
// the counterpart of the reference's 4x4 `eigh` at networkx algebraicconnectivity.py:239 -- needs
// only the smallest eigenpair of T_k, which is O(k) work per bisection step and is done here on
// the host while the device runs the next batch of Lanczos steps.
#include "tridiag.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <limits>

namespace macb {

namespace {

// True iff T_k has at least one eigenvalue < x (Sturm sequence; early exit on the first negative pivot).
inline bool has_eig_below(const double* a, const double* b2, int k, double x, double pivmin) {
    double q = a[0] - x;
    if (std::fabs(q) < pivmin) q = -pivmin;
    if (q < 0.0) return true;
    for (int i = 1; i < k; ++i) {
        q = (a[i] - x) - b2[i] / q;
        if (std::fabs(q) < pivmin) q = -pivmin;
        if (q < 0.0) return true;
    }
    return false;
}

}  // namespace

double tridiag_smallest_value(const double* a, const double* b, int k, double hint_hi, double hint_delta) {
    if (k <= 0) return 0.0;
    if (k == 1) return a[0];
    std::vector<double> b2(k);
    double bmax = 0.0;
    b2[0] = 0.0;
    for (int i = 1; i < k; ++i) {
        b2[i] = b[i] * b[i];
        bmax = std::max(bmax, b2[i]);
    }
    const double pivmin = std::max(DBL_MIN, DBL_MIN * bmax) * 4.0;
    // Gershgorin bounds
    double gl = std::numeric_limits<double>::infinity(), gu = -gl;
    for (int i = 0; i < k; ++i) {
        double r = (i > 0 ? std::fabs(b[i]) : 0.0) + (i + 1 < k ? std::fabs(b[i + 1]) : 0.0);
        gl = std::min(gl, a[i] - r);
        gu = std::max(gu, a[i] + r);
    }
    const double tnorm = std::max(std::fabs(gl), std::fabs(gu));
    double lo = gl - 2.0 * DBL_EPSILON * tnorm * k - 2.0 * pivmin;
    double hi = gu + 2.0 * DBL_EPSILON * tnorm * k + 2.0 * pivmin;
    // smallest diagonal entry is an upper bound on the smallest eigenvalue
    double amin = a[0];
    for (int i = 1; i < k; ++i) amin = std::min(amin, a[i]);
    hi = std::min(hi, amin + 4.0 * DBL_EPSILON * tnorm);
    bool warm = false;
    if (std::isfinite(hint_hi)) {
        double h = hint_hi + 8.0 * DBL_EPSILON * tnorm;
        if (h < hi && has_eig_below(a, b2.data(), k, h, pivmin)) {
            hi = h;
            warm = true;
        }
    }
    if (warm) {
        // walk down from hi in growing strides until the interval below holds no eigenvalue
        double delta = (hint_delta > 0.0) ? hint_delta : 1e-6 * std::max(std::fabs(hi), 1e-300);
        delta = std::max(delta, 16.0 * DBL_EPSILON * tnorm);
        for (int it = 0; it < 64; ++it) {
            double x = hi - delta;
            if (x <= lo) break;
            if (has_eig_below(a, b2.data(), k, x, pivmin)) {
                hi = x;
                delta *= 8.0;
            } else {
                lo = x;
                break;
            }
        }
    }
    for (int it = 0; it < 200; ++it) {
        double mid = 0.5 * (lo + hi);
        if (!(hi - lo > 2.0 * DBL_EPSILON * std::max(std::fabs(lo), std::fabs(hi)) + 2.0 * pivmin)) break;
        if (mid <= lo || mid >= hi) break;
        if (has_eig_below(a, b2.data(), k, mid, pivmin))
            hi = mid;
        else
            lo = mid;
    }
    return 0.5 * (lo + hi);
}

double tridiag_vector(const double* a, const double* b, int k, double theta, double* s) {
    if (k == 1) {
        s[0] = 1.0;
        return std::fabs(a[0] - theta);
    }
    double tnorm = 0.0;
    for (int i = 0; i < k; ++i)
        tnorm = std::max(tnorm, std::fabs(a[i]) + (i > 0 ? std::fabs(b[i]) : 0.0) + (i + 1 < k ? std::fabs(b[i + 1]) : 0.0));
    const double tiny = std::max(DBL_EPSILON * tnorm, DBL_MIN * 1e8);

    // Twisted factorisation of T - theta I (Parlett & Dhillon): forward pivots dp, backward pivots dm.
    std::vector<double> dp(k), dm(k), lp(k), um(k);
    dp[0] = a[0] - theta;
    for (int i = 0; i + 1 < k; ++i) {
        if (std::fabs(dp[i]) < tiny) dp[i] = (dp[i] < 0 ? -tiny : tiny);
        lp[i] = b[i + 1] / dp[i];
        dp[i + 1] = (a[i + 1] - theta) - lp[i] * b[i + 1];
    }
    dm[k - 1] = a[k - 1] - theta;
    for (int i = k - 2; i >= 0; --i) {
        if (std::fabs(dm[i + 1]) < tiny) dm[i + 1] = (dm[i + 1] < 0 ? -tiny : tiny);
        um[i] = b[i + 1] / dm[i + 1];
        dm[i] = (a[i] - theta) - um[i] * b[i + 1];
    }
    int r = 0;
    double best = std::numeric_limits<double>::infinity();
    for (int i = 0; i < k; ++i) {
        double gamma = dp[i] + dm[i] - (a[i] - theta);
        if (std::fabs(gamma) < best) {
            best = std::fabs(gamma);
            r = i;
        }
    }
    // Solve N z = gamma_r e_r with z_r = 1; rescale on the fly to avoid overflow on long recurrences.
    s[r] = 1.0;
    for (int i = r - 1; i >= 0; --i) {
        s[i] = -lp[i] * s[i + 1];
        if (std::fabs(s[i]) > 1e150) {
            double sc = 1e-150;
            for (int t = i; t <= r; ++t) s[t] *= sc;
        }
    }
    for (int i = r; i + 1 < k; ++i) {
        s[i + 1] = -um[i] * s[i];
        if (std::fabs(s[i + 1]) > 1e150) {
            double sc = 1e-150;
            for (int t = 0; t <= i + 1; ++t) s[t] *= sc;
        }
    }
    // normalise (scaled 2-norm)
    double mx = 0.0;
    for (int i = 0; i < k; ++i) mx = std::max(mx, std::fabs(s[i]));
    if (!(mx > 0.0) || !std::isfinite(mx)) {
        for (int i = 0; i < k; ++i) s[i] = 0.0;
        s[r] = 1.0;
        mx = 1.0;
    }
    double ss = 0.0;
    for (int i = 0; i < k; ++i) {
        s[i] /= mx;
        ss += s[i] * s[i];
    }
    double inv = 1.0 / std::sqrt(ss);
    for (int i = 0; i < k; ++i) s[i] *= inv;
    // residual
    double rr = 0.0;
    for (int i = 0; i < k; ++i) {
        double t = (a[i] - theta) * s[i];
        if (i > 0) t += b[i] * s[i - 1];
        if (i + 1 < k) t += b[i + 1] * s[i + 1];
        rr += t * t;
    }
    return std::sqrt(rr);
}

}  // namespace macb
