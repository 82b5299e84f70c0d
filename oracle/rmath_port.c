/*
 * oracle/rmath_port.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (log-density only) of the distribution functions that the reference's hot path
 * reaches through nimble's `calculateDiff` / `calculate`
 * (call sites: /root/reference/nimbleAPT/R/APT_functions.R:711,715,825-829, via nimble, which calls
 * R's Rmath).  Neither nimble nor Rmath is vendored in /root/reference (DESCRIPTION:19-21 gives
 * `Depends: nimble`, no version) and neither is installed here, so this file restates the PUBLISHED
 * algorithms of Rmath (R <= 4.0.x flavour: Loader's saddle-point `stirlerr`/`bd0` for dbinom/dpois,
 * dgamma and dbeta built on them, dnorm4, dunif) and nimble's dmnorm_chol.
 *
 * PARITY UNPINNED by the reference (it ships no tests / golden vectors): these functions are pinned
 * instead against scipy.stats (independent implementation) in tests/test_oracle_dists.py.
 *
 * Known, documented deviations: lgamma() is glibc's (Rmath has its own lgammafn), lbeta() is
 * lgamma(a)+lgamma(b)-lgamma(a+b) for min(a,b) < 10 and the asymptotic form otherwise without Rmath's
 * lgammacor Chebyshev series (replaced by the equivalent Stirling series), and R >= 4.1's `ebd0`
 * refinement of dpois_raw is not reproduced; all agree with Rmath to ~1e-14 relative.
 */
#include "rmath_port.h"
#include <float.h>
#include <math.h>
#include "sferr_table.h"

#define LN_SQRT_2PI 0.918938533204672741780329736406 /* log(sqrt(2*pi)) */
#define LN_2PI 1.837877066409345483560659472811      /* log(2*pi) */
#define TWO_PI 6.283185307179586476925286766559

static const double NEG_INF = -INFINITY;

/* stirlerr(n) = log(n!) - log( sqrt(2*pi*n)*(n/e)^n ) */
double apt_stirlerr(double n) {
    const double S0 = 1.0 / 12.0, S1 = 1.0 / 360.0, S2 = 1.0 / 1260.0, S3 = 1.0 / 1680.0, S4 = 1.0 / 1188.0;
    double nn;
    if (n <= 15.0) {
        nn = n + n;
        if (nn == (double)(int)nn) return APT_SFERR_HALVES[(int)nn];
        return lgamma(n + 1.0) - (n + 0.5) * log(n) + n - LN_SQRT_2PI;
    }
    nn = n * n;
    if (n > 500) return (S0 - S1 / nn) / n;
    if (n > 80) return (S0 - (S1 - S2 / nn) / nn) / n;
    if (n > 35) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
    return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}

/* bd0(x, np) = x log(x/np) + np - x, evaluated without cancellation near x ~ np */
double apt_bd0(double x, double np) {
    if (!isfinite(x) || !isfinite(np) || np == 0.0) return NAN;
    if (fabs(x - np) < 0.1 * (x + np)) {
        double v = (x - np) / (x + np);
        double s = (x - np) * v;
        if (fabs(s) < DBL_MIN) return s;
        double ej = 2 * x * v;
        v = v * v;
        for (int j = 1; j < 1000; j++) {
            ej *= v;
            double s1 = s + ej / ((j << 1) + 1);
            if (s1 == s) return s1;
            s = s1;
        }
    }
    return x * log(x / np) + np - x;
}

double apt_dpois_raw_log(double x, double lambda) {
    if (lambda == 0) return (x == 0) ? 0.0 : NEG_INF;
    if (!isfinite(lambda)) return NEG_INF;
    if (x < 0) return NEG_INF;
    if (x <= lambda * DBL_MIN) return -lambda;
    if (lambda < x * DBL_MIN) {
        if (!isfinite(x)) return NEG_INF;
        return -lambda + x * log(lambda) - lgamma(x + 1);
    }
    return -0.5 * log(TWO_PI * x) + (-apt_stirlerr(x) - apt_bd0(x, lambda));
}

double apt_dbinom_raw_log(double x, double n, double p, double q) {
    double lc, lf;
    if (p == 0) return (x == 0) ? 0.0 : NEG_INF;
    if (q == 0) return (x == n) ? 0.0 : NEG_INF;
    if (x == 0) {
        if (n == 0) return 0.0;
        lc = (p < 0.1) ? -apt_bd0(n, n * q) - n * p : n * log(q);
        return lc;
    }
    if (x == n) {
        lc = (q < 0.1) ? -apt_bd0(n, n * p) - n * q : n * log(p);
        return lc;
    }
    if (x < 0 || x > n) return NEG_INF;
    lc = apt_stirlerr(n) - apt_stirlerr(x) - apt_stirlerr(n - x) - apt_bd0(x, n * p) - apt_bd0(n - x, n * q);
    lf = LN_2PI + log(x) + log1p(-x / n);
    return lc - 0.5 * lf;
}

static int nonint(double x) { return fabs(x - nearbyint(x)) > 1e-7 * fmax(1.0, fabs(x)); }

double apt_dnorm_log(double x, double mu, double sigma) {
    if (isnan(x) || isnan(mu) || isnan(sigma)) return x + mu + sigma;
    if (!isfinite(sigma)) return NEG_INF;
    if (!isfinite(x) && mu == x) return NAN;
    if (sigma <= 0) {
        if (sigma < 0) return NAN;
        return (x == mu) ? INFINITY : NEG_INF;
    }
    x = (x - mu) / sigma;
    if (!isfinite(x)) return NEG_INF;
    x = fabs(x);
    if (x >= 2 * sqrt(DBL_MAX)) return NEG_INF;
    return -(LN_SQRT_2PI + 0.5 * x * x + log(sigma));
}

double apt_dunif_log(double x, double a, double b) {
    if (isnan(x) || isnan(a) || isnan(b)) return x + a + b;
    if (b <= a) return NAN;
    if (a <= x && x <= b) return -log(b - a);
    return NEG_INF;
}

double apt_dgamma_log(double x, double shape, double scale) {
    double pr;
    if (isnan(x) || isnan(shape) || isnan(scale)) return x + shape + scale;
    if (shape < 0 || scale <= 0) return NAN;
    if (x < 0) return NEG_INF;
    if (shape == 0) return (x == 0) ? INFINITY : NEG_INF;
    if (x == 0) {
        if (shape < 1) return INFINITY;
        if (shape > 1) return NEG_INF;
        return -log(scale);
    }
    if (shape < 1) {
        pr = apt_dpois_raw_log(shape, x / scale);
        return pr + (isfinite(shape / x) ? log(shape / x) : log(shape) - log(x));
    }
    pr = apt_dpois_raw_log(shape - 1, x / scale);
    return pr - log(scale);
}

/* Stirling-series correction lgamma(x) - [(x-.5)log x - x + log sqrt(2pi)], x >= 10 */
static double lgammacor_series(double x) {
    double x2 = x * x;
    return (1.0 / 12.0 - (1.0 / 360.0 - (1.0 / 1260.0 - (1.0 / 1680.0 - (1.0 / 1188.0) / x2) / x2) / x2) / x2) / x;
}

double apt_lbeta(double a, double b) {
    double p = fmin(a, b), q = fmax(a, b), corr;
    if (p < 0) return NAN;
    if (p == 0) return INFINITY;
    if (!isfinite(q)) return NEG_INF;
    if (p >= 10) {
        corr = lgammacor_series(p) + lgammacor_series(q) - lgammacor_series(p + q);
        return log(q) * -0.5 + LN_SQRT_2PI + corr + (p - 0.5) * log(p / (p + q)) + q * log1p(-p / (p + q));
    } else if (q >= 10) {
        corr = lgammacor_series(q) - lgammacor_series(p + q);
        return lgamma(p) + corr + p - p * log(p + q) + (q - 0.5) * log1p(-p / (p + q));
    }
    return lgamma(p) + lgamma(q) - lgamma(p + q);
}

double apt_dbeta_log(double x, double a, double b) {
    double lval;
    if (isnan(x) || isnan(a) || isnan(b)) return x + a + b;
    if (a < 0 || b < 0) return NAN;
    if (x < 0 || x > 1) return NEG_INF;
    if (a == 0 || b == 0 || !isfinite(a) || !isfinite(b)) {
        if (a == 0 && b == 0) return (x == 0 || x == 1) ? INFINITY : NEG_INF;
        if (a == 0 || a / b == 0) return (x == 0) ? INFINITY : NEG_INF;
        if (b == 0 || b / a == 0) return (x == 1) ? INFINITY : NEG_INF;
        return (x == 0.5) ? INFINITY : NEG_INF;
    }
    if (x == 0) {
        if (a > 1) return NEG_INF;
        if (a < 1) return INFINITY;
        return log(b);
    }
    if (x == 1) {
        if (b > 1) return NEG_INF;
        if (b < 1) return INFINITY;
        return log(a);
    }
    if (a <= 2 || b <= 2)
        lval = (a - 1) * log(x) + (b - 1) * log1p(-x) - apt_lbeta(a, b);
    else
        lval = log(a + b - 1) + apt_dbinom_raw_log(a - 1, a + b - 2, x, 1 - x);
    return lval;
}

double apt_dbinom_log(double x, double n, double p) {
    if (isnan(x) || isnan(n) || isnan(p)) return x + n + p;
    if (p < 0 || p > 1 || n < 0 || nonint(n)) return NAN;
    if (nonint(x) || x < 0 || !isfinite(x)) return NEG_INF;
    n = nearbyint(n);
    x = nearbyint(x);
    return apt_dbinom_raw_log(x, n, p, 1 - p);
}

double apt_dpois_log(double x, double lambda) {
    if (isnan(x) || isnan(lambda)) return x + lambda;
    if (lambda < 0) return NAN;
    if (nonint(x)) return NEG_INF;
    if (x < 0 || !isfinite(x)) return NEG_INF;
    x = nearbyint(x);
    return apt_dpois_raw_log(x, lambda);
}

double apt_dexp_log(double x, double scale) {
    if (isnan(x) || isnan(scale)) return x + scale;
    if (scale <= 0.0) return NAN;
    if (x < 0.) return NEG_INF;
    return (-x / scale) - log(scale);
}

/*
 * nimble's dcat(x, prob, K, give_log) [NIMBLE-UNVERIFIED, restated from its documented behaviour: "prob" need not
 * be normalised; values outside 1..K and non-integers have probability 0]: log(prob[x] / sum(prob)).
 * Used by the demo model of the reference (nimbleAPT/demo/APT_demo.R:84, `k ~ dcat(pk[1:K])`).
 */
double apt_dcat_log(double x, const double* prob, int K) {
    if (isnan(x)) return x;
    double sum = 0.0;
    for (int i = 0; i < K; i++) {
        if (isnan(prob[i])) return prob[i];
        sum += prob[i];
    }
    if (x < 1.0 || x > (double)K || x != floor(x)) return NEG_INF;
    return log(prob[(int)x - 1] / sum);
}

/*
 * nimble's dmulti(x, size, prob, K, give_log) [NIMBLE-UNVERIFIED, restated from its documented behaviour: "prob" is
 * normalised internally; sum(x) must equal size]: lgamma(size + 1) + sum_i (x_i log(prob_i / sum(prob)) - lgamma(x_i + 1)),
 * terms with x_i = 0 and prob_i = 0 skipped.  Target distribution of sampler_RW_multinomial_tempered (APT:1046).
 */
double apt_dmulti_log(const double* x, const double* prob, int K, double size) {
    double sumProb = 0.0, sumX = 0.0;
    if (isnan(size)) return size;
    for (int i = 0; i < K; i++) {
        if (isnan(prob[i])) return prob[i];
        if (isnan(x[i])) return x[i];
        sumProb += prob[i];
        sumX += x[i];
    }
    for (int i = 0; i < K; i++)
        if (x[i] < 0 || x[i] != floor(x[i])) return NEG_INF;
    if (sumX > size + 10 * 2.220446049250313e-16 || sumX < size - 10 * 2.220446049250313e-16) return NEG_INF;
    double lp = lgamma(size + 1.0);
    for (int i = 0; i < K; i++)
        if (!(x[i] == 0.0 && prob[i] == 0.0)) lp += x[i] * log(prob[i] / sumProb) - lgamma(x[i] + 1.0);
    return lp;
}

/*
 * rbinom(1, size, prob) by inversion of ONE uniform (R's rbinom consumes a data-dependent number of draws from its
 * sequential stream, which a slot-addressed generator cannot reproduce; include/aptb200.h documents the convention):
 * the smallest k with u <= P(X <= k), the pmf walked from k = 0 by its recurrence, on the smaller of prob / 1 - prob.
 * size <= 1000 (no underflow of P(X = 0)).
 */
double apt_rbinom_inv(double n, double p, double u) {
    if (n <= 0 || p <= 0) return 0.0;
    if (p >= 1) return n;
    const int flip = p > 0.5;
    const double pp = flip ? 1.0 - p : p;
    const double r = pp / (1.0 - pp);
    double f = exp(n * log1p(-pp));
    double k = 0.0;
    while (u > f && k < n) {
        u -= f;
        k += 1.0;
        f *= r * ((n - k + 1.0) / k);
    }
    return flip ? n - k : k;
}

/*
 * nimble's dmnorm_chol(x, mean, chol, n, prec_param, give_log) [NIMBLE-UNVERIFIED, restated from its
 * documented behaviour]: chol is the UPPER Cholesky factor U (U'U = cov or precision), column-major.
 *   dens = -n*log(sqrt(2pi)) + (prec_param ? +1 : -1) * sum_i log U_ii
 *   prec_param:  z = U (x-mean)          (triangular multiply)
 *   otherwise :  z = U^-T (x-mean)       (forward solve with U')
 *   dens += -0.5 * z'z
 */
double apt_dmnorm_chol_log(const double* x, const double* mean, const double* chol, int n, int prec_param,
                           double* work) {
    double dens = -n * LN_SQRT_2PI;
    int i, j;
    for (i = 0; i < n; i++) {
        if (prec_param)
            dens += log(chol[i * n + i]);
        else
            dens -= log(chol[i * n + i]);
    }
    for (i = 0; i < n; i++) work[i] = x[i] - mean[i];
    if (prec_param) {
        /* z = U * xCopy ; U upper, column-major U[r + c*n] */
        for (i = 0; i < n; i++) {
            double s = 0;
            for (j = i; j < n; j++) s += chol[i + j * n] * work[j];
            work[i] = s; /* safe: row i only reads work[j>=i], and work[i] is consumed first */
        }
    } else {
        /* solve U' z = xCopy : forward substitution, U'[i][j] = U[j + i*n] for j<=i */
        for (i = 0; i < n; i++) {
            double s = work[i];
            for (j = 0; j < i; j++) s -= chol[j + i * n] * work[j];
            work[i] = s / chol[i + i * n];
        }
    }
    double ss = 0;
    for (i = 0; i < n; i++) ss += work[i] * work[i];
    dens += -0.5 * ss;
    return dens;
}

/* Upper Cholesky U with U'U = A (column-major, n x n). Returns 0 on success, -1 if not positive definite. */
int apt_chol_upper(const double* A, double* U, int n) {
    int i, j, k;
    for (i = 0; i < n * n; i++) U[i] = 0.0;
    for (j = 0; j < n; j++) {
        double s = A[j + j * n];
        for (k = 0; k < j; k++) s -= U[k + j * n] * U[k + j * n];
        if (!(s > 0)) return -1;
        double d = sqrt(s);
        U[j + j * n] = d;
        for (i = j + 1; i < n; i++) {
            double t = A[j + i * n];
            for (k = 0; k < j; k++) t -= U[k + j * n] * U[k + i * n];
            U[j + i * n] = t / d;
        }
    }
    return 0;
}
