// apt_dists.cuh — device log-densities for the lowered BUGS subset (sm_100a, fp64).
//
// These stand in for what the reference reaches through nimble's calculateDiff()/calculate()
// (call sites /root/reference/nimbleAPT/R/APT_functions.R:711,715,825-829): nimble evaluates each
// stochastic node with R's Rmath d*(…, log=TRUE).  The algorithms below follow the published Rmath
// algorithms (Loader's saddle-point stirlerr/bd0 for dbinom/dpois, dgamma/dbeta built on them, dnorm4,
// dunif) so that logProb agrees with the reference path to ~1e-14 relative; they are written for the GPU
// (branch-light, no error signalling: invalid parameters yield NaN exactly as Rmath's ML_ERR_return_NAN does,
// which the samplers then reject, nimble `decide`: NaN -> FALSE).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "apt_sferr_table.cuh"
#include "apt_softplus.cuh"

namespace apt {

#define APT_LN_SQRT_2PI 0.918938533204672741780329736406
#define APT_LN_2PI 1.837877066409345483560659472811
#define APT_TWO_PI 6.283185307179586476925286766559
#define APT_DBL_MIN 2.2250738585072014e-308
#define APT_NEG_INF (-CUDART_INF)
#ifndef CUDART_INF
#define CUDART_INF __longlong_as_double(0x7ff0000000000000LL)
#endif
#ifndef CUDART_NAN
#define CUDART_NAN __longlong_as_double(0xfff8000000000000LL)
#endif

__device__ __forceinline__ double stirlerr(double n) {
    const double S0 = 1.0 / 12.0, S1 = 1.0 / 360.0, S2 = 1.0 / 1260.0, S3 = 1.0 / 1680.0, S4 = 1.0 / 1188.0;
    if (n <= 15.0) {
        double nn = n + n;
        if (nn == (double)(int)nn) return APT_SFERR_HALVES_DEV[(int)nn];
        return lgamma(n + 1.0) - (n + 0.5) * log(n) + n - APT_LN_SQRT_2PI;
    }
    double nn = n * n;
    if (n > 500) return (S0 - S1 / nn) / n;
    if (n > 80) return (S0 - (S1 - S2 / nn) / nn) / n;
    if (n > 35) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
    return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}

__device__ __forceinline__ double bd0(double x, double np) {
    if (!isfinite(x) || !isfinite(np) || np == 0.0) return CUDART_NAN;
    if (fabs(x - np) < 0.1 * (x + np)) {
        double v = (x - np) / (x + np);
        double s = (x - np) * v;
        if (fabs(s) < APT_DBL_MIN) return s;
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

__device__ __forceinline__ double dpois_raw_log(double x, double lambda) {
    if (lambda == 0) return (x == 0) ? 0.0 : APT_NEG_INF;
    if (!isfinite(lambda)) return APT_NEG_INF;
    if (x < 0) return APT_NEG_INF;
    if (x <= lambda * APT_DBL_MIN) return -lambda;
    if (lambda < x * APT_DBL_MIN) {
        if (!isfinite(x)) return APT_NEG_INF;
        return -lambda + x * log(lambda) - lgamma(x + 1);
    }
    return -0.5 * log(APT_TWO_PI * x) + (-stirlerr(x) - bd0(x, lambda));
}

__device__ __forceinline__ double dbinom_raw_log(double x, double n, double p, double q) {
    if (p == 0) return (x == 0) ? 0.0 : APT_NEG_INF;
    if (q == 0) return (x == n) ? 0.0 : APT_NEG_INF;
    if (x == 0) {
        if (n == 0) return 0.0;
        return (p < 0.1) ? -bd0(n, n * q) - n * p : n * log(q);
    }
    if (x == n) return (q < 0.1) ? -bd0(n, n * p) - n * q : n * log(p);
    if (x < 0 || x > n) return APT_NEG_INF;
    double lc = stirlerr(n) - stirlerr(x) - stirlerr(n - x) - bd0(x, n * p) - bd0(n - x, n * q);
    double lf = APT_LN_2PI + log(x) + log1p(-x / n);
    return lc - 0.5 * lf;
}

__device__ __forceinline__ bool nonint(double x) { return fabs(x - nearbyint(x)) > 1e-7 * fmax(1.0, fabs(x)); }

// dnorm4(x, mu, sigma, log = TRUE) with all of Rmath's special cases; out of line: the callers below take it only when the
// standardised value is not an ordinary number (the checks were 8 % of the small-model kernel's instructions)
__device__ __noinline__ double dnorm_log_slow(double x, double mu, double sigma) {
    if (isnan(x) || isnan(mu) || isnan(sigma)) return x + mu + sigma;
    if (!isfinite(sigma)) return APT_NEG_INF;
    if (!isfinite(x) && mu == x) return CUDART_NAN;
    if (sigma <= 0) {
        if (sigma < 0) return CUDART_NAN;
        return (x == mu) ? CUDART_INF : APT_NEG_INF;
    }
    x = (x - mu) / sigma;
    if (!isfinite(x)) return APT_NEG_INF;
    x = fabs(x);
    if (x >= 2.6815615859885194e154) return APT_NEG_INF;  // 2*sqrt(DBL_MAX)
    return -(APT_LN_SQRT_2PI + 0.5 * x * x + apt_log_pos(sigma));
}
// One test on the high word of z = (x - mu) / sigma covers every special case of the function above: NaN or Inf in any
// argument, sigma == 0 (z is Inf or NaN), |z| >= 2^510 (the -Inf cut-off lies above).  sigma < 0 or denormal is caught on
// sigma's own high word.  On the ordinary path the arithmetic is that of dnorm4, operation for operation.
__device__ __forceinline__ double dnorm_log(double x, double mu, double sigma) {
    const double z = (x - mu) / sigma;
    const unsigned hs = (unsigned)__double2hiint(sigma) - 0x00100000u;  // positive normal finite <=> hs < 0x7fe00000
    if ((__double2hiint(z) & 0x7fffffff) >= 0x5fd00000 || hs >= 0x7fe00000u) return dnorm_log_slow(x, mu, sigma);
    return -(APT_LN_SQRT_2PI + 0.5 * z * z + apt_log_pos(sigma));
}

// dnorm with a compile-time constant, positive, finite sd: log(sd) is supplied by the lowering as a literal
// (the library log() is not constant-folded by nvcc and dominated the small-model kernel, profiles/)
__device__ __noinline__ double dnorm_log_csd_slow(double x, double mu, double sigma, double log_sigma) {
    if (isnan(x) || isnan(mu)) return x + mu;
    if (!isfinite(x) && mu == x) return CUDART_NAN;
    x = (x - mu) / sigma;
    if (!isfinite(x)) return APT_NEG_INF;
    x = fabs(x);
    if (x >= 2.6815615859885194e154) return APT_NEG_INF;
    return -(APT_LN_SQRT_2PI + 0.5 * x * x + log_sigma);
}
__device__ __forceinline__ double dnorm_log_csd(double x, double mu, double sigma, double log_sigma) {
    const double z = (x - mu) / sigma;
    if ((__double2hiint(z) & 0x7fffffff) >= 0x5fd00000) return dnorm_log_csd_slow(x, mu, sigma, log_sigma);
    return -(APT_LN_SQRT_2PI + 0.5 * z * z + log_sigma);
}

__device__ __forceinline__ double dunif_log(double x, double a, double b) {
    if (isnan(x) || isnan(a) || isnan(b)) return x + a + b;
    if (b <= a) return CUDART_NAN;
    if (a <= x && x <= b) return -log(b - a);
    return APT_NEG_INF;
}

__device__ __forceinline__ double dgamma_log(double x, double shape, double scale) {
    if (isnan(x) || isnan(shape) || isnan(scale)) return x + shape + scale;
    if (shape < 0 || scale <= 0) return CUDART_NAN;
    if (x < 0) return APT_NEG_INF;
    if (shape == 0) return (x == 0) ? CUDART_INF : APT_NEG_INF;
    if (x == 0) {
        if (shape < 1) return CUDART_INF;
        if (shape > 1) return APT_NEG_INF;
        return -log(scale);
    }
    if (shape < 1) {
        double pr = dpois_raw_log(shape, x / scale);
        return pr + (isfinite(shape / x) ? log(shape / x) : log(shape) - log(x));
    }
    return dpois_raw_log(shape - 1, x / scale) - log(scale);
}

__device__ __forceinline__ double lgammacor_series(double x) {
    double x2 = x * x;
    return (1.0 / 12.0 - (1.0 / 360.0 - (1.0 / 1260.0 - (1.0 / 1680.0 - (1.0 / 1188.0) / x2) / x2) / x2) / x2) / x;
}

__device__ __forceinline__ double lbeta_fn(double a, double b) {
    double p = fmin(a, b), q = fmax(a, b);
    if (p < 0) return CUDART_NAN;
    if (p == 0) return CUDART_INF;
    if (!isfinite(q)) return APT_NEG_INF;
    if (p >= 10) {
        double corr = lgammacor_series(p) + lgammacor_series(q) - lgammacor_series(p + q);
        return log(q) * -0.5 + APT_LN_SQRT_2PI + corr + (p - 0.5) * log(p / (p + q)) + q * log1p(-p / (p + q));
    } else if (q >= 10) {
        double corr = lgammacor_series(q) - lgammacor_series(p + q);
        return lgamma(p) + corr + p - p * log(p + q) + (q - 0.5) * log1p(-p / (p + q));
    }
    return lgamma(p) + lgamma(q) - lgamma(p + q);
}

__device__ __forceinline__ double dbeta_log(double x, double a, double b) {
    if (isnan(x) || isnan(a) || isnan(b)) return x + a + b;
    if (a < 0 || b < 0) return CUDART_NAN;
    if (x < 0 || x > 1) return APT_NEG_INF;
    if (a == 0 || b == 0 || !isfinite(a) || !isfinite(b)) {
        if (a == 0 && b == 0) return (x == 0 || x == 1) ? CUDART_INF : APT_NEG_INF;
        if (a == 0 || a / b == 0) return (x == 0) ? CUDART_INF : APT_NEG_INF;
        if (b == 0 || b / a == 0) return (x == 1) ? CUDART_INF : APT_NEG_INF;
        return (x == 0.5) ? CUDART_INF : APT_NEG_INF;
    }
    if (x == 0) {
        if (a > 1) return APT_NEG_INF;
        if (a < 1) return CUDART_INF;
        return log(b);
    }
    if (x == 1) {
        if (b > 1) return APT_NEG_INF;
        if (b < 1) return CUDART_INF;
        return log(a);
    }
    if (a <= 2 || b <= 2) return (a - 1) * log(x) + (b - 1) * log1p(-x) - lbeta_fn(a, b);
    return log(a + b - 1) + dbinom_raw_log(a - 1, a + b - 2, x, 1 - x);
}

// canonical BUGS order (prob, size)
__device__ __forceinline__ double dbinom_log(double x, double p, double n) {
    if (isnan(x) || isnan(n) || isnan(p)) return x + n + p;
    if (p < 0 || p > 1 || n < 0 || nonint(n)) return CUDART_NAN;
    if (nonint(x) || x < 0 || !isfinite(x)) return APT_NEG_INF;
    n = nearbyint(n);
    x = nearbyint(x);
    return dbinom_raw_log(x, n, p, 1 - p);
}

__device__ __forceinline__ double dpois_log(double x, double lambda) {
    if (isnan(x) || isnan(lambda)) return x + lambda;
    if (lambda < 0) return CUDART_NAN;
    if (nonint(x)) return APT_NEG_INF;
    if (x < 0 || !isfinite(x)) return APT_NEG_INF;
    x = nearbyint(x);
    return dpois_raw_log(x, lambda);
}

// ---- dpois with a DATA value x.  What depends on x alone is computed once per observation by the module's
// derived-array kernel (pre0 = -0.5 log(2 pi x), pre1 = stirlerr(x); pre1 = NaN marks an x that is not a positive
// integer, which takes the generic path); per evaluation remain exp(eta), one division and the bd0 series / log.
// Deviations from the literal restatement, each within one ulp of an intermediate: ej * (1/(2j+1)) for
// ej / (2j+1) in the bd0 series, apt_log_pos for log.
__constant__ double APT_ODD_RCP_DEV[64] = {
    1.0 / 1, 1.0 / 3, 1.0 / 5, 1.0 / 7, 1.0 / 9, 1.0 / 11, 1.0 / 13, 1.0 / 15, 1.0 / 17, 1.0 / 19, 1.0 / 21, 1.0 / 23, 1.0 / 25, 1.0 / 27, 1.0 / 29, 1.0 / 31,
    1.0 / 33, 1.0 / 35, 1.0 / 37, 1.0 / 39, 1.0 / 41, 1.0 / 43, 1.0 / 45, 1.0 / 47, 1.0 / 49, 1.0 / 51, 1.0 / 53, 1.0 / 55, 1.0 / 57, 1.0 / 59, 1.0 / 61, 1.0 / 63,
    1.0 / 65, 1.0 / 67, 1.0 / 69, 1.0 / 71, 1.0 / 73, 1.0 / 75, 1.0 / 77, 1.0 / 79, 1.0 / 81, 1.0 / 83, 1.0 / 85, 1.0 / 87, 1.0 / 89, 1.0 / 91, 1.0 / 93, 1.0 / 95,
    1.0 / 97, 1.0 / 99, 1.0 / 101, 1.0 / 103, 1.0 / 105, 1.0 / 107, 1.0 / 109, 1.0 / 111, 1.0 / 113, 1.0 / 115, 1.0 / 117, 1.0 / 119, 1.0 / 121, 1.0 / 123, 1.0 / 125, 1.0 / 127};
__device__ __noinline__ double dpois_log_slow(double x, double lambda) { return dpois_log(x, lambda); }
__device__ __forceinline__ void dpois_pre(double x, double* out) {  // derived-array entry of one observation
    const bool ok = isfinite(x) && x >= 1.0 && x == nearbyint(x);
    out[0] = ok ? -0.5 * log(APT_TWO_PI * x) : 0.0;
    out[1] = ok ? stirlerr(x) : CUDART_NAN;
}
__device__ __forceinline__ double dpois_log_pre(double x, double lambda, double pre0, double pre1) {
    if (!(lambda > 0.0) || !(lambda <= 1.0e300)) return dpois_log_slow(x, lambda);  // 0, negative, NaN, Inf, huge
    if (x == 0.0) return -lambda;                                                   // dpois_raw: x <= lambda * DBL_MIN
    if (!(pre1 == pre1) || lambda < x * APT_DBL_MIN) return dpois_log_slow(x, lambda);
    // bd0(x, lambda), x >= 1 and lambda finite
    const double d = x - lambda, sm = x + lambda;
    double b;
    if (fabs(d) < 0.1 * sm) {
        double v = d / sm;
        double s = d * v;
        if (fabs(s) < APT_DBL_MIN) b = s;
        else {
            double ej = 2 * x * v;
            v = v * v;
            int j = 1;
            for (; j < 64; j++) {
                ej *= v;
                const double s1 = s + ej * APT_ODD_RCP_DEV[j];
                if (s1 == s) break;
                s = s1;
            }
            if (j == 64) return dpois_log_slow(x, lambda);  // not converged in 63 terms: never for |v| < 0.1
            b = s;
        }
    } else b = x * apt_log_pos(x / lambda) + lambda - x;
    return pre0 + (-pre1 - b);
}

__device__ __forceinline__ double dexp_log(double x, double scale) {
    if (isnan(x) || isnan(scale)) return x + scale;
    if (scale <= 0.0) return CUDART_NAN;
    if (x < 0.) return APT_NEG_INF;
    return (-x / scale) - log(scale);
}

// nimble's dcat(x, prob) [NIMBLE-UNVERIFIED: "prob" need not be normalised; values outside 1..K and non-integers
// have probability 0]; the reference's demo model uses it for its discrete node (demo/APT_demo.R:84).
template <int K>
__device__ __forceinline__ double dcat_log(double x, const double (&prob)[K]) {
    if (isnan(x)) return x;
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i < K; i++) sum += prob[i];
    if (isnan(sum)) return sum;
    if (x < 1.0 || x > (double)K || x != floor(x)) return APT_NEG_INF;
    double px = prob[0];
    const int ix = (int)x - 1;
#pragma unroll
    for (int i = 1; i < K; i++) px = (i == ix) ? prob[i] : px;  // no dynamic indexing of a register array
    return log(px / sum);
}
// nimble's dmulti(x, size, prob) [NIMBLE-UNVERIFIED: "prob" is normalised internally; sum(x) must equal size]; the
// target distribution of sampler_RW_multinomial_tempered (APT_functions.R:1046).
template <int K>
__device__ __forceinline__ double dmulti_log(const double (&x)[K], const double (&prob)[K], double size) {
    double sumProb = 0.0, sumX = 0.0;
    if (isnan(size)) return size;
#pragma unroll
    for (int i = 0; i < K; i++) {
        if (isnan(prob[i])) return prob[i];
        if (isnan(x[i])) return x[i];
        sumProb += prob[i];
        sumX += x[i];
    }
#pragma unroll
    for (int i = 0; i < K; i++)
        if (x[i] < 0 || x[i] != floor(x[i])) return APT_NEG_INF;
    if (sumX > size + 10 * 2.220446049250313e-16 || sumX < size - 10 * 2.220446049250313e-16) return APT_NEG_INF;
    double lp = lgamma(size + 1.0);
#pragma unroll
    for (int i = 0; i < K; i++)
        if (!(x[i] == 0.0 && prob[i] == 0.0)) lp += x[i] * log(prob[i] / sumProb) - lgamma(x[i] + 1.0);
    return lp;
}

// rbinom(1, size, prob) by inversion of ONE slot-addressed uniform (include/aptb200.h): the smallest k with
// u <= P(X <= k), the pmf walked from k = 0 by its recurrence on the smaller of prob / 1 - prob.  size <= 1000.
__device__ __noinline__ double rbinom_inv(double n, double p, double u) {
    if (n <= 0 || p <= 0) return 0.0;
    if (p >= 1) return n;
    const bool flip = p > 0.5;
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
// element of a constant array selected by a sampled node (e.g. rfx[k]): clamped into the array, so that a state the
// prior excludes (its logProb is -Inf) cannot read out of bounds
__device__ __forceinline__ int dyn_index(double v, int n) {
    return (int)fmin(fmax(v, 1.0), (double)n);
}

// Fused form of  y ~ dbinom(prob = ilogit(eta), size = 1)  (lowering peephole, build option fuse_links):
//   log p(y | eta) = -(max(t, 0) + log(1 + exp(-|t|))),  t = (1 - 2y) eta          (apt_softplus.cuh)
// instead of the reference's chain  p = 1/(1+exp(-eta)); q = 1-p; dbinom_raw(y, 1, p, q).  Walking through
// dbinom_raw for n = 1 shows that the chain equals the exact value to ~1e-16 ABSOLUTE in every branch but one:
//   * "good sign" (y=1, eta>0 | y=0, eta<0): -bd0(1, big) - small = -small - small^2/2 - ..., small = 1-big known to
//     1.1e-16 absolute  -> equals -log1p(exp(-|eta|)) to 1e-16;
//   * y=1, eta<0: log(p), p = 1/(1+exp(|eta|)) computed directly -> relative 2e-16;  -Inf once exp(-eta) overflows;
//   * y=0, eta>0: log(q), q = 1 - 1/(1+exp(-eta)): q carries an absolute error 1.1e-16, i.e. the term is wrong by
//     1.1e-16 e^eta, and saturates to -Inf beyond eta = 36.74 (p rounds to 1).
// To give the SAME accept/reject decisions as the reference even for hot chains that wander there, that one branch
// is reproduced literally for eta > 8 (below 8 both agree to < 3.3e-13); everything else takes the fused form.
// threshold of the literal branch and its high word (a power of two or 1.5 times one, so that the high-word test is exact)
#ifndef APT_LIT_ETA
#define APT_LIT_ETA 8.0
#define APT_LIT_HI 0x40200000
#endif
__device__ __noinline__ double bern_logit_invalid(double y, double eta) {
    if (isnan(y) || isnan(eta)) return y + eta;
    return APT_NEG_INF;  // dbinom: non-integer or out-of-range x
}
// the reference's chain verbatim; only reached for NaN / Inf / |eta| > 709 (exp about to overflow)
__device__ __noinline__ double bern_logit_generic(double y, double eta) {
    if (isnan(eta)) return eta;
    const double p = 1.0 / (1.0 + exp(-eta));
    return dbinom_raw_log(y, 1.0, p, 1.0 - p);
}
// |eta| > 709 (exp about to overflow) with a valid response.  Hot chains live here: a rung at T ~ 1e5 under a tempered
// N(0, 10^2) prior has |beta| in the thousands, and every row of such a chain used to take the out-of-line library chain.
// Beyond 710 that chain has only two outcomes: exp(-eta) is 0 or Inf, p is exactly 1 or 0, and dbinom_raw(y, 1, p, q) returns
// log(1) = 0 for the matching response and log(0) = -Inf for the other one.  (Between 709 and 710 p can still be a denormal
// whose logarithm is finite, and Inf / NaN propagate: those stay with the verbatim chain.)
__device__ __forceinline__ double bern_logit_far(double y, double eta) {
    const double a = fabs(eta);
    if (a > 710.0 && a < CUDART_INF) return ((y == 0.0) == (eta > 0.0)) ? APT_NEG_INF : 0.0;
    return bern_logit_generic(y, eta);
}

__device__ __noinline__ double bern_logit_literal_q(double eta) {  // y = 0, eta > 8; out of line: one copy instead of one per term
    const double e = apt_exp_neg(eta);
    const double p = 1.0 / (1.0 + e);
    const double q = 1.0 - p;
    return apt_log_pos(q);  // q == 0 -> -Inf
}
__device__ __forceinline__ double bern_logit_log(double y, double eta) {
    if (!(y == 0.0 || y == 1.0)) return bern_logit_invalid(y, eta);
    if (!(fabs(eta) <= 709.0)) return bern_logit_far(y, eta);
    const double t = (y == 1.0) ? -eta : eta;
    double r = -(fmax(t, 0.0) + apt_softplus_neg(fmin(fabs(t), 64.0)));
    if (y == 0.0 && eta > APT_LIT_ETA) r = bern_logit_literal_q(eta);
    return r;
}

// every rule that overrides the fused value, in one out-of-line function (one copy in the kernel, reached rarely
// for cold rungs and for a fraction of the rows of hot ones)
__device__ __noinline__ double bern_logit_fix(double y, double eta, double v) {
    if (!(y == 0.0 || y == 1.0)) return bern_logit_invalid(y, eta);
    if (!(fabs(eta) <= 709.0)) return bern_logit_generic(y, eta);
    if (y == 0.0 && eta > APT_LIT_ETA) return bern_logit_literal_q(eta);
    return v;
}

// N chains at once (grid engine): acc[c] += log p(y | eta[c]); fused terms in batches of 8 in lock step.  The two
// fix-ups (literal q branch, non-finite / overflowing eta) are tested once per batch, not once per term.
template <int N>
__device__ __forceinline__ void bern_logit_acc(double y, const double (&eta)[N], double (&acc)[N]) {
    constexpr int M = N >= 8 ? 8 : N;
    const bool y0 = (y == 0.0);
    if (!(y0 || y == 1.0)) {
#pragma unroll
        for (int j = 0; j < N; j++) acc[j] += bern_logit_invalid(y, eta[j]);
        return;
    }
    const double sgn = y0 ? 1.0 : -1.0;
#pragma unroll
    for (int h = 0; h < N; h += M) {
        // Bookkeeping on the high words with integer min/max (one issue slot each) instead of FP64 compares and
        // selects: amax = largest |eta| of the batch, smax = largest eta (as signed integers, which order finite
        // doubles of equal sign like the doubles themselves; NaN and Inf land above every finite value).
        double a[M], v[M];
        int amax = 0, smax = (int)0x80000000;
#pragma unroll
        for (int j = 0; j < M; j++) {
            const int hi = __double2hiint(eta[h + j]);
            amax = max(amax, hi & 0x7fffffff);
            smax = max(smax, hi);
            a[j] = fabs(eta[h + j]);
        }
        apt_softplus_neg_batch<M>(a, v);
        // -(max(t, 0) + softplus), t = sgn * eta:  t + |t| is exactly 2 max(t, 0), so one rounding as before
#pragma unroll
        for (int j = 0; j < M; j++) v[j] = fma(-0.5, fma(sgn, eta[h + j], a[j]), -v[j]);
        if ((y0 && smax >= APT_LIT_HI) || amax >= 0x40862800) {  // conservative on the high words; exact tests inside
#pragma unroll
            for (int j = 0; j < M; j++)
                if (!(fabs(eta[h + j]) <= 709.0)) v[j] = bern_logit_far(y, eta[h + j]);
                else if (y0 && eta[h + j] > APT_LIT_ETA) v[j] = bern_logit_literal_q(eta[h + j]);
        }
#pragma unroll
        for (int j = 0; j < M; j++) acc[h + j] += v[j];
    }
}

// Thread tile of the grid engine: NR rows (each with its own y) x NC chains.  A batch is CPB chains x all NR rows (8 terms in
// lock step); per chain the rows are accumulated in ascending order.  Same arithmetic and fix-ups as bern_logit_acc.
// The literal branch (y = 0, eta > 8) is taken per CHAIN, i.e. per column of the tile: whether it is needed depends almost
// only on how hot the chain is (a chain at T ~ 1e5 roams the prior and has |eta| > 8 on most rows, a cold chain never), and
// the grid kernel deals the rungs to the threads with a stride (apt_grid.cuh), so that every thread of a warp holds exactly one
// of the hottest rungs in the same column: the branch is warp-uniform in the steady state and its 1/w and log run on all
// lanes for that one column, instead of on a quarter of the lanes for every column as with consecutive rungs per thread.
template <int NR, int NC>
__device__ __forceinline__ void bern_logit_tile(const double (&y)[NR], const double (&eta)[NR][NC], double (&acc)[NC]) {
#ifndef APT_TILE_BATCH
#define APT_TILE_BATCH 8
#endif
    constexpr int CPB = (NR >= APT_TILE_BATCH) ? 1 : ((APT_TILE_BATCH / NR) < NC ? (APT_TILE_BATCH / NR) : NC);  // chains per batch
    constexpr int M = CPB * NR;
    static_assert(NC % CPB == 0, "chains per thread must be a multiple of the batch");
    bool valid = true;
#pragma unroll
    for (int r = 0; r < NR; r++) valid = valid && (y[r] == 0.0 || y[r] == 1.0);
#pragma unroll
    for (int c0 = 0; c0 < NC; c0 += CPB) {
        double a[M], v[M];
        int amax = 0, smax[CPB];  // largest |eta|; per chain the largest eta among the y = 0 rows (high words)
#pragma unroll
        for (int cc = 0; cc < CPB; cc++) {
            smax[cc] = (int)0x80000000;
#pragma unroll
            for (int r = 0; r < NR; r++) {
                const double e = eta[r][c0 + cc];
                const int hi = __double2hiint(e);
                amax = max(amax, hi & 0x7fffffff);
                smax[cc] = max(smax[cc], (y[r] == 0.0) ? hi : (int)0x80000000);
                a[cc * NR + r] = fabs(e);
            }
        }
        double w[M];  // 1 + exp(-|eta|)
        apt_softplus_neg_batch<M>(a, v, w);
#pragma unroll
        for (int cc = 0; cc < CPB; cc++)
#pragma unroll
            for (int r = 0; r < NR; r++) v[cc * NR + r] = fma(-0.5, fma(1.0 - 2.0 * y[r], eta[r][c0 + cc], a[cc * NR + r]), -v[cc * NR + r]);
#pragma unroll
        for (int cc = 0; cc < CPB; cc++)
            if (smax[cc] >= APT_LIT_HI) {
                // y = 0 and eta > 8: the reference's own chain p = 1/(1+exp(-eta)), q = 1-p, log(q), whose rounding the
                // decisions of hot chains depend on (see above) -- the rows of this chain in lock step, reusing 1+exp(-eta)
                double lq[NR];
#pragma unroll
                for (int r = 0; r < NR; r++) lq[r] = 1.0 - apt_rcp_1to2(w[cc * NR + r]);  // w in [1, 2]: bit for bit the quotient 1 / w
#pragma unroll
                for (int r = 0; r < NR; r++) lq[r] = apt_log_q(lq[r]);  // q == 0 -> -Inf; q is 0 or normal (a multiple of 2^-53)
#pragma unroll
                for (int r = 0; r < NR; r++)
                    if (y[r] == 0.0 && eta[r][c0 + cc] > APT_LIT_ETA) v[cc * NR + r] = lq[r];
            }
        if (!valid || amax >= 0x40862800) {  // invalid response, |eta| >= 709, Inf, NaN: exact tests inside
#pragma unroll
            for (int cc = 0; cc < CPB; cc++)
#pragma unroll
                for (int r = 0; r < NR; r++) {
                    const double e = eta[r][c0 + cc];
                    if (!valid) v[cc * NR + r] = bern_logit_fix(y[r], e, v[cc * NR + r]);
                    else if (!(fabs(e) <= 709.0)) v[cc * NR + r] = bern_logit_far(y[r], e);
                }
        }
#pragma unroll
        for (int cc = 0; cc < CPB; cc++)
#pragma unroll
            for (int r = 0; r < NR; r++) acc[c0 + cc] += v[cc * NR + r];
    }
}

__device__ __forceinline__ double ilogit(double x) { return 1.0 / (1.0 + exp(-x)); }
__device__ __forceinline__ double logit(double p) { return log(p / (1.0 - p)); }
__device__ __forceinline__ double iprobit(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }
__device__ __forceinline__ double icloglog(double x) { return 1.0 - exp(-exp(x)); }
__device__ __forceinline__ double cloglog(double p) { return log(-log(1.0 - p)); }
__device__ __forceinline__ double stepfn(double x) { return x >= 0 ? 1.0 : 0.0; }

// nimble dmnorm_chol(x, mean, chol = upper U column-major, n, prec_param): small fixed n, register arrays.
template <int N>
__device__ __forceinline__ double dmnorm_chol_log(const double (&x)[N], const double (&mu)[N], const double (&U)[N * N],
                                                  bool prec_param) {
    double dens = -N * APT_LN_SQRT_2PI;
    double w[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        dens += prec_param ? log(U[i * N + i]) : -log(U[i * N + i]);
        w[i] = x[i] - mu[i];
    }
    if (prec_param) {
#pragma unroll
        for (int i = 0; i < N; i++) {
            double s = 0;
#pragma unroll
            for (int j = i; j < N; j++) s += U[i + j * N] * w[j];
            w[i] = s;
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; i++) {
            double s = w[i];
#pragma unroll
            for (int j = 0; j < i; j++) s -= U[j + i * N] * w[j];
            w[i] = s / U[i + i * N];
        }
    }
    double ss = 0;
#pragma unroll
    for (int i = 0; i < N; i++) ss += w[i] * w[i];
    return dens - 0.5 * ss;
}

}  // namespace apt
