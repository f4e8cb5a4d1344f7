/* libjwasio.so, second part: the O(markers x annotations) work of the annotation prior update
 * (MCMC/annotation_updates.jl:43-123, 177-189, 260-304) as threaded host C.  Declarations in include/jwas_io.h.
 * Results do not depend on the thread count: reductions run over fixed 4096-element chunks whose partial sums are
 * added in chunk order. */
#include "../../../include/jwas_io.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define JWANN_CHUNK 4096

static inline double phi_cdf(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }

/* inverse of the standard normal CDF: rational start (Acklam, relative error 1.2e-9), one Halley step on Phi(x) - p
 * (third-order: the step leaves ~1e-16) */
static double phi_inv(double p) {
    static const double a[6] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                                1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b[5] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                                6.680131188771972e+01, -1.328068155288572e+01};
    static const double c[6] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                                -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
    static const double d[4] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                                3.754408661907416e+00};
    if (!(p > 0.0)) return -INFINITY;
    if (!(p < 1.0)) return INFINITY;
    double x;
    if (p < 0.02425) {
        double q = sqrt(-2.0 * log(p));
        x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
            ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0);
    } else if (p <= 1.0 - 0.02425) {
        double q = p - 0.5, r = q * q;
        x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
            (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1.0);
    } else {
        double q = sqrt(-2.0 * log1p(-p));
        x = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
             ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0);
    }
    if (fabs(x) <= 37.0) {                               /* beyond, exp(x^2/2) would overflow; the start is good to 1e-9 */
        double e = phi_cdf(x) - p;
        double u = e * 2.50662827463100050242 * exp(0.5 * x * x);
        x -= u / (1.0 + 0.5 * x * u);
    }
    return x;
}

static void set_threads(int n_threads) {
#ifdef _OPENMP
    omp_set_num_threads(n_threads > 0 ? n_threads : omp_get_num_procs());
#else
    (void)n_threads;
#endif
}

/* sum_i v[i] * w(i), w = 1 when x == NULL, else x[idx ? idx[i] : i]; deterministic */
static double chunked_dot(const double* v, const double* x, const int64_t* idx, int64_t n) {
    const int64_t nch = (n + JWANN_CHUNK - 1) / JWANN_CHUNK;
    double* part = (double*)malloc(sizeof(double) * (size_t)(nch > 0 ? nch : 1));
    if (!part) return NAN;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < nch; ++c) {
        const int64_t i0 = c * JWANN_CHUNK, i1 = i0 + JWANN_CHUNK < n ? i0 + JWANN_CHUNK : n;
        double s = 0.0;
        if (!x) for (int64_t i = i0; i < i1; ++i) s += v[i];
        else if (!idx) for (int64_t i = i0; i < i1; ++i) s += v[i] * x[i];
        else for (int64_t i = i0; i < i1; ++i) s += v[i] * x[idx[i]];
        part[c] = s;
    }
    double tot = 0.0;
    for (int64_t c = 0; c < nch; ++c) tot += part[c];
    free(part);
    return tot;
}

int jwann_probit_step(int64_t m, int k, const double* Xc, const int64_t* active, int64_t n_active,
                      const int32_t* response, double* coeffs, double prior_var,
                      const double* uniforms, const double* normals,
                      double* liability, double* mu, int n_threads) {
    if (!Xc || !response || !coeffs || !uniforms || !normals || !liability || !mu || m <= 0 || k <= 0) return 1;
    set_threads(n_threads);
    const int64_t n = active ? n_active : m;
    if (n > 0) {
        double* resid = (double*)malloc(sizeof(double) * (size_t)n);
        if (!resid) return 2;
        /* liabilities: l ~ N(mu, 1) truncated to [0, inf) when z = 1, (-inf, 0] when z = 0, by inversion in the
         * lower tail of the side that is kept (annotation_updates.jl:43-59) */
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            const int64_t j = active ? active[i] : i;
            const double mj = mu[j];
            const int one = response[j] != 0;
            const double s = one ? mj : -mj;
            double u = uniforms[i];
            if (u < 1e-300) u = 1e-300;
            if (u > 1.0) u = 1.0;
            double w = phi_inv(u * phi_cdf(s));
            if (!isfinite(w)) w = s < -37.0 ? s : -37.0;
            double l = one ? mj - w : mj + w;
            if (one) { if (l < 0.0) l = 0.0; } else { if (l > 0.0) l = 0.0; }
            liability[j] = l;
            resid[i] = l - mj;
        }
        /* coordinate Gibbs (annotation_updates.jl:98-123): flat-prior intercept, slopes ~ N(0, prior_var) */
        {
            const double old = coeffs[0], inv_lhs = 1.0 / (double)n;
            const double ahat = inv_lhs * (chunked_dot(resid, NULL, NULL, n) + (double)n * old);
            coeffs[0] = normals[0] * sqrt(inv_lhs) + ahat;
            const double shift = old - coeffs[0];
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < n; ++i) resid[i] += shift;
        }
        for (int q = 1; q < k; ++q) {
            const double* xq = Xc + (int64_t)q * m;
            double diag = 0.0;
            {   /* x_q'x_q over the markers of this step, same chunking */
                const int64_t nch = (n + JWANN_CHUNK - 1) / JWANN_CHUNK;
                double* part = (double*)malloc(sizeof(double) * (size_t)nch);
                if (!part) { free(resid); return 2; }
#pragma omp parallel for schedule(static)
                for (int64_t c = 0; c < nch; ++c) {
                    const int64_t i0 = c * JWANN_CHUNK, i1 = i0 + JWANN_CHUNK < n ? i0 + JWANN_CHUNK : n;
                    double s = 0.0;
                    for (int64_t i = i0; i < i1; ++i) { const double v = xq[active ? active[i] : i]; s += v * v; }
                    part[c] = s;
                }
                for (int64_t c = 0; c < nch; ++c) diag += part[c];
                free(part);
            }
            const double old = coeffs[q];
            const double inv_lhs = 1.0 / (diag + 1.0 / prior_var);
            const double ahat = inv_lhs * (chunked_dot(resid, xq, active, n) + diag * old);
            coeffs[q] = normals[q] * sqrt(inv_lhs) + ahat;
            const double shift = old - coeffs[q];
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < n; ++i) resid[i] += xq[active ? active[i] : i] * shift;
        }
        free(resid);
    }
    /* linear predictor of ALL markers for the new coefficients */
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < m; ++j) {
        double s = 0.0;
        for (int q = 0; q < k; ++q) s += Xc[(int64_t)q * m + j] * coeffs[q];
        mu[j] = s;
    }
    return 0;
}

int jwann_probit_probability(const double* mu, int64_t m, int complement, double* prob, int n_threads) {
    if (!mu || !prob || m < 0) return 1;
    set_threads(n_threads);
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < m; ++j) {
        double p = phi_cdf(mu[j]);
        if (complement) p = 1.0 - p;
        if (p < DBL_EPSILON) p = DBL_EPSILON;
        if (p > 1.0 - DBL_EPSILON) p = 1.0 - DBL_EPSILON;
        prob[j] = p;
    }
    return 0;
}

/* exported for the tests: the two special functions */
double jwann_phi_inv(double p) { return phi_inv(p); }
double jwann_phi_cdf(double x) { return phi_cdf(x); }
