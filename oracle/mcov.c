/* Oracle (TEST INFRASTRUCTURE): plain-C restatement of the reference's Cython
 * nucleoatac/multinomial_cov.pyx:19-31 `calculateCov(p, v, r)`.
 *
 *   value = sum_i p_i (1-p_i) v_i^2  +  sum_{i<j} p_i p_j (-2) v_i v_j ;  return value * r
 *
 * mcov_pairwise follows the reference loop nest literally (same i<=j order, same
 * expression shapes, accumulator starting at 0 -- the reference leaves it
 * uninitialised, multinomial_cov.pyx:23, which happens to be 0).  `r` is a C int
 * in the reference (float arguments are truncated by the Cython call wrapper);
 * the Python wrapper oracle/mcov.py performs that truncation.
 *
 * mcov_closed is the O(n) identity r*(sum p v^2 - (sum p v)^2) that the CUDA
 * path uses; tests pin it against mcov_pairwise and against the reference's own
 * compiled .pyx (oracle/_ref) to <= 1e-12 relative.
 *
 * Build: gcc -O2 -shared -fPIC -o oracle/libmcov.so oracle/mcov.c
 */
#include <stddef.h>

double mcov_pairwise(const double *p, const double *v, size_t n, int r)
{
    double value = 0.0;
    for (size_t i = 0; i < n; i++) {
        for (size_t j = i; j < n; j++) {
            if (i == j)
                value += p[i] * (1 - p[i]) * (v[i] * v[i]);
            else
                value += p[i] * p[j] * -2 * v[i] * v[j];
        }
    }
    return value * r;
}

double mcov_closed(const double *p, const double *v, size_t n, int r)
{
    double s1 = 0.0, s2 = 0.0;
    for (size_t i = 0; i < n; i++) {
        double pv = p[i] * v[i];
        s1 += pv;
        s2 += pv * v[i];
    }
    return (s2 - s1 * s1) * r;
}
