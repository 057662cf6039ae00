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
#if defined(__x86_64__)
#include <immintrin.h>
#endif

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

#if defined(__x86_64__)
// Eight Sturm recurrences at once (two AVX2 vectors): neg[j] != 0 <=> T - x[j] I has a negative pivot.
__attribute__((target("avx2"))) void sturm8_avx2(const double* a, const double* b2, int k, const double* x, double pivmin,
                                                 double* neg) {
    const __m256d pm = _mm256_set1_pd(pivmin), npm = _mm256_set1_pd(-pivmin), zero = _mm256_setzero_pd();
    const __m256d absmask = _mm256_castsi256_pd(_mm256_set1_epi64x(0x7fffffffffffffffLL));
    const __m256d x0 = _mm256_loadu_pd(x), x1 = _mm256_loadu_pd(x + 4);
    __m256d q0 = _mm256_sub_pd(_mm256_set1_pd(a[0]), x0), q1 = _mm256_sub_pd(_mm256_set1_pd(a[0]), x1);
    q0 = _mm256_blendv_pd(q0, npm, _mm256_cmp_pd(_mm256_and_pd(q0, absmask), pm, _CMP_LT_OQ));
    q1 = _mm256_blendv_pd(q1, npm, _mm256_cmp_pd(_mm256_and_pd(q1, absmask), pm, _CMP_LT_OQ));
    __m256d n0 = _mm256_cmp_pd(q0, zero, _CMP_LT_OQ), n1 = _mm256_cmp_pd(q1, zero, _CMP_LT_OQ);
    for (int i = 1; i < k; ++i) {
        const __m256d ai = _mm256_set1_pd(a[i]), bi = _mm256_set1_pd(b2[i]);
        __m256d t0 = _mm256_sub_pd(_mm256_sub_pd(ai, x0), _mm256_div_pd(bi, q0));
        __m256d t1 = _mm256_sub_pd(_mm256_sub_pd(ai, x1), _mm256_div_pd(bi, q1));
        t0 = _mm256_blendv_pd(t0, npm, _mm256_cmp_pd(_mm256_and_pd(t0, absmask), pm, _CMP_LT_OQ));
        t1 = _mm256_blendv_pd(t1, npm, _mm256_cmp_pd(_mm256_and_pd(t1, absmask), pm, _CMP_LT_OQ));
        n0 = _mm256_or_pd(n0, _mm256_cmp_pd(t0, zero, _CMP_LT_OQ));
        n1 = _mm256_or_pd(n1, _mm256_cmp_pd(t1, zero, _CMP_LT_OQ));
        q0 = t0;
        q1 = t1;
    }
    alignas(32) double m[8];
    _mm256_store_pd(m, n0);
    _mm256_store_pd(m + 4, n1);
    for (int j = 0; j < 8; ++j) {
        unsigned long long bits;
        __builtin_memcpy(&bits, &m[j], 8);
        neg[j] = bits ? 1.0 : 0.0;
    }
}
#endif

void sturm8(const double* a, const double* b2, int k, const double* x, double pivmin, double* neg) {
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) {
        sturm8_avx2(a, b2, k, x, pivmin, neg);
        return;
    }
#endif
    for (int j = 0; j < 8; ++j) neg[j] = has_eig_below(a, b2, k, x[j], pivmin) ? 1.0 : 0.0;
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
    // Multisection: kSec interior shifts per sweep, evaluated together (the division chain of one Sturm
    // recurrence is latency-bound; eight independent chains fill the divider and vectorise), so one sweep
    // shrinks the bracket 9x instead of 2x.
    constexpr int kSec = 8;   // sturm8 evaluates exactly eight shifts
    for (int it = 0; it < 80; ++it) {
        const double width = hi - lo;
        if (!(width > 2.0 * DBL_EPSILON * std::max(std::fabs(lo), std::fabs(hi)) + 2.0 * pivmin)) break;
        alignas(64) double x[kSec], neg[kSec];   // neg[j] > 0 <=> chain j has seen a negative pivot
        for (int j = 0; j < kSec; ++j) {
            x[j] = lo + width * (double)(j + 1) / (double)(kSec + 1);
            neg[j] = 0.0;
        }
        if (!(x[0] > lo) || !(x[kSec - 1] < hi)) {  // bracket at rounding resolution: finish with plain bisection
            double mid = 0.5 * (lo + hi);
            if (mid <= lo || mid >= hi) break;
            if (has_eig_below(a, b2.data(), k, mid, pivmin)) hi = mid; else lo = mid;
            continue;
        }
        sturm8(a, b2.data(), k, x, pivmin, neg);
        // eigenvalue lies between the last shift with no negative pivot and the first with one
        int first_neg = kSec;
        for (int j = kSec - 1; j >= 0; --j)
            if (neg[j] > 0.0) first_neg = j;
        const double new_lo = (first_neg == 0) ? lo : x[first_neg - 1];
        const double new_hi = (first_neg == kSec) ? hi : x[first_neg];
        lo = new_lo;
        hi = new_hi;
    }
    return 0.5 * (lo + hi);
}

namespace {

double tridiag_residual(const double* a, const double* b, int k, double theta, const double* s) {
    double rr = 0.0;
    for (int i = 0; i < k; ++i) {
        double t = (a[i] - theta) * s[i];
        if (i > 0) t += b[i] * s[i - 1];
        if (i + 1 < k) t += b[i + 1] * s[i + 1];
        rr += t * t;
    }
    return std::sqrt(rr);
}

// Eigenvector for the SMALLEST eigenvalue by the forward pivot recurrence alone: theta <= lambda_min(T_j) for
// every leading block (interlacing), so T_j - theta I is positive (semi)definite, the pivots q_i are positive
// and s_{i+1} = -(q_i / b_{i+1}) s_i is the stable LDL^T forward substitution.  One pass, one division per row.
bool forward_vector(const double* a, const double* b, int k, double theta, double tnorm, double* s) {
    const double tiny = std::max(DBL_EPSILON * tnorm, DBL_MIN * 1e8);
    double q = a[0] - theta;
    s[0] = 1.0;
    double ss = 1.0;
    for (int i = 0; i + 1 < k; ++i) {
        const double bn = b[i + 1];
        if (!(std::fabs(bn) > 0.0)) return false;
        double si = -(q / bn) * s[i];
        if (!(std::fabs(si) < 1e140)) return false;   // growth: leave it to the twisted factorisation
        s[i + 1] = si;
        ss += si * si;
        if (std::fabs(q) < tiny) q = (q < 0 ? -tiny : tiny);
        q = (a[i + 1] - theta) - bn * bn / q;
    }
    if (!(ss > 0.0) || !std::isfinite(ss)) return false;
    const double inv = 1.0 / std::sqrt(ss);
    for (int i = 0; i < k; ++i) s[i] *= inv;
    return true;
}

}  // namespace

double tridiag_vector(const double* a, const double* b, int k, double theta, double* s) {
    if (k == 1) {
        s[0] = 1.0;
        return std::fabs(a[0] - theta);
    }
    {
        double tn = 0.0;
        for (int i = 0; i < k; ++i)
            tn = std::max(tn, std::fabs(a[i]) + (i > 0 ? std::fabs(b[i]) : 0.0) + (i + 1 < k ? std::fabs(b[i + 1]) : 0.0));
        if (forward_vector(a, b, k, theta, tn, s)) {
            const double r = tridiag_residual(a, b, k, theta, s);
            if (r <= 1e-11 * tn) return r;
        }
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
    return tridiag_residual(a, b, k, theta, s);
}

}  // namespace macb
