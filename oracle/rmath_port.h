/* oracle/rmath_port.h — TEST INFRASTRUCTURE (see rmath_port.c). Log-densities only. */
#ifndef APT_ORACLE_RMATH_PORT_H
#define APT_ORACLE_RMATH_PORT_H
#ifdef __cplusplus
extern "C" {
#endif
double apt_stirlerr(double n);
double apt_bd0(double x, double np);
double apt_dpois_raw_log(double x, double lambda);
double apt_dbinom_raw_log(double x, double n, double p, double q);
double apt_dnorm_log(double x, double mu, double sigma);
double apt_dunif_log(double x, double a, double b);
double apt_dgamma_log(double x, double shape, double scale);
double apt_lbeta(double a, double b);
double apt_dbeta_log(double x, double a, double b);
double apt_dbinom_log(double x, double n, double p);
double apt_dpois_log(double x, double lambda);
double apt_dexp_log(double x, double scale);
double apt_dcat_log(double x, const double* prob, int K);
double apt_dmulti_log(const double* x, const double* prob, int K, double size);
double apt_rbinom_inv(double n, double p, double u);
double apt_dmnorm_chol_log(const double* x, const double* mean, const double* chol, int n, int prec_param,
                           double* work);
int apt_chol_upper(const double* A, double* U, int n);
#ifdef __cplusplus
}
#endif
#endif
