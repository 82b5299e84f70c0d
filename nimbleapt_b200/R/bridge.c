/* bridge.c — thin .Call bridge between R and libaptb200.so (include/aptb200.h).
 *
 * This image has no R; build where R exists with
 *     R CMD SHLIB bridge.c -I../../include -L.. -laptb200
 * Here it is compiled (-Wall -Werror) and executed against tests/rstub/, a stand-in for the entry points of R's C
 * API used below (tests/test_r_bridge.py: marshalling, errors, interrupts, finalizers; CPU and GPU).
 * Pure marshalling: every R object is allocated by R, the library only fills caller-allocated buffers, a non-zero
 * status becomes Rf_error() AFTER the library call has returned (no longjmp across the C ABI).
 * Replaces the external-pointer plumbing that nimble generates for a compiled buildAPT object
 * (/root/reference/nimbleAPT/R/APT_functions.R:157-162, compileNimble). */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Utils.h>
#include <string.h>
#include "aptb200.h"

static void chk(int rc) { if (rc) Rf_error("%s", apt_last_error()); }

static void model_finalizer(SEXP p) { apt_model* m = R_ExternalPtrAddr(p); if (m) { apt_model_free(m); R_ClearExternalPtr(p); } }
static void engine_finalizer(SEXP p) { apt_engine* e = R_ExternalPtrAddr(p); if (e) { apt_free(e); R_ClearExternalPtr(p); } }

SEXP aptR_model_create(SEXP code) {
    apt_model* m = NULL;
    chk(apt_model_create(CHAR(STRING_ELT(code, 0)), &m));
    SEXP p = PROTECT(R_MakeExternalPtr(m, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(p, model_finalizer, TRUE);
    UNPROTECT(1);
    return p;
}

/* a user-supplied deterministic function the model code calls: name, deparse()d text of the nimbleFunction's run function */
SEXP aptR_model_add_function(SEXP p, SEXP name, SEXP src) {
    chk(apt_model_add_function(R_ExternalPtrAddr(p), CHAR(STRING_ELT(name, 0)), CHAR(STRING_ELT(src, 0))));
    return R_NilValue;
}

/* role: 0 constants, 1 data, 2 inits; x is a numeric vector/matrix (column-major, as R stores it) */
SEXP aptR_model_set_array(SEXP p, SEXP name, SEXP role, SEXP x) {
    SEXP dim = Rf_getAttrib(x, R_DimSymbol);
    int64_t dims[3] = {0, 0, 0};
    int nd;
    if (dim == R_NilValue) { nd = 1; dims[0] = XLENGTH(x); }
    else { nd = LENGTH(dim); if (nd > 3) Rf_error("arrays of more than 3 dimensions are not supported"); for (int i = 0; i < nd; i++) dims[i] = INTEGER(dim)[i]; }
    SEXP xr = PROTECT(Rf_coerceVector(x, REALSXP));
    int rc = apt_model_set_array(R_ExternalPtrAddr(p), CHAR(STRING_ELT(name, 0)), Rf_asInteger(role), REAL(xr), dims, nd);
    UNPROTECT(1);
    chk(rc);
    return R_NilValue;
}

static SEXP list_get(SEXP lst, const char* nm) {
    SEXP names = Rf_getAttrib(lst, R_NamesSymbol);
    for (R_xlen_t i = 0; i < XLENGTH(lst); i++) if (!strcmp(CHAR(STRING_ELT(names, i)), nm)) return VECTOR_ELT(lst, i);
    return R_NilValue;
}

/* control: the sampler's control list, keys as the reference's code spells them (APT:654-660, 775-781) */
SEXP aptR_model_add_sampler(SEXP p, SEXP target, SEXP type, SEXP control) {
    apt_sampler_control c;
    apt_sampler_control_default(&c);
    SEXP v;
    if ((v = list_get(control, "log")) != R_NilValue) c.log = Rf_asLogical(v);
    if ((v = list_get(control, "reflective")) != R_NilValue) c.reflective = Rf_asLogical(v);
    if ((v = list_get(control, "adaptive")) != R_NilValue) c.adaptive = Rf_asLogical(v);
    if ((v = list_get(control, "adaptScaleOnly")) != R_NilValue) c.adaptScaleOnly = Rf_asLogical(v);
    if ((v = list_get(control, "adaptInterval")) != R_NilValue) c.adaptInterval = Rf_asInteger(v);
    if ((v = list_get(control, "adaptFactorExponent")) != R_NilValue) c.adaptFactorExponent = Rf_asReal(v);
    if ((v = list_get(control, "scale")) != R_NilValue) c.scale = Rf_asReal(v);
    if ((v = list_get(control, "temperPriors")) != R_NilValue) c.temperPriors = Rf_asLogical(v);
    if ((v = list_get(control, "width")) != R_NilValue) c.width = Rf_asReal(v);                    /* APT:911 */
    if ((v = list_get(control, "sliceMaxSteps")) != R_NilValue) c.sliceMaxSteps = Rf_asInteger(v);  /* APT:912 */
    if ((v = list_get(control, "useTempering")) != R_NilValue) c.useTempering = Rf_asLogical(v);    /* APT:1015 */
    SEXP pc = list_get(control, "propCov");
    int prot = 0;
    if (pc != R_NilValue && Rf_isNumeric(pc)) {
        SEXP dim = Rf_getAttrib(pc, R_DimSymbol);
        if (dim == R_NilValue || LENGTH(dim) != 2) Rf_error("propCov must be a matrix\n");
        pc = PROTECT(Rf_coerceVector(pc, REALSXP)); prot = 1;
        c.propCov = REAL(pc);
        c.propCov_dim = INTEGER(dim)[0];
    }
    int rc = apt_model_add_sampler(R_ExternalPtrAddr(p), CHAR(STRING_ELT(target, 0)), CHAR(STRING_ELT(type, 0)), &c);
    UNPROTECT(prot);
    chk(rc);
    return R_NilValue;
}

SEXP aptR_model_set_monitors(SEXP p, SEXP which, SEXP names, SEXP thin, SEXP thin2) {
    int n = LENGTH(names);
    const char** nm = (const char**)R_alloc(n > 0 ? n : 1, sizeof(char*));
    for (int i = 0; i < n; i++) nm[i] = CHAR(STRING_ELT(names, i));
    chk(apt_model_set_monitors(R_ExternalPtrAddr(p), Rf_asInteger(which), nm, n));
    chk(apt_model_set_thin(R_ExternalPtrAddr(p), Rf_asInteger(thin), Rf_asInteger(thin2)));
    return R_NilValue;
}

SEXP aptR_model_names(SEXP p, SEXP which) {
    char* buf = R_alloc(1 << 20, 1);
    chk(apt_model_names(R_ExternalPtrAddr(p), Rf_asInteger(which), buf, 1 << 20));
    return Rf_mkString(buf); /* '\n'-separated; split on the R side */
}

SEXP aptR_build(SEXP p, SEXP temps, SEXP monitorTmax, SEXP ULT, SEXP nLadders, SEXP seed, SEXP device, SEXP enableWAIC) {
    apt_build_opts o;
    apt_build_opts_default(&o);
    o.monitorTmax = Rf_asLogical(monitorTmax); o.ULT = Rf_asReal(ULT); o.n_ladders = Rf_asInteger(nLadders);
    o.seed = (uint32_t)Rf_asInteger(seed); o.device = Rf_asInteger(device); o.verbose = 1;
    o.enable_waic = Rf_asLogical(enableWAIC);
    apt_engine* e = NULL;
    SEXP t = PROTECT(Rf_coerceVector(temps, REALSXP));
    int rc = apt_build(R_ExternalPtrAddr(p), REAL(t), LENGTH(t), &o, &e);
    UNPROTECT(1);
    chk(rc);
    SEXP h = PROTECT(R_MakeExternalPtr(e, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(h, engine_finalizer, TRUE);
    UNPROTECT(1);
    return h;
}

/* checkInterrupt() (APT:415): R_CheckUserInterrupt may longjmp, so it is probed inside R_ToplevelExec */
static void probe_interrupt(void* flag) { R_CheckUserInterrupt(); *(int*)flag = 0; }
static int should_abort(void* unused) { int interrupted = 1; R_ToplevelExec(probe_interrupt, &interrupted); return interrupted; }

/* samplesOut: R_NilValue, or a numeric vector of n_ladders x (niter / thin) x n_monitors doubles that receives THIS run's
 * rows of mvSamples while the run is still sampling (apt_run_args.samples_out): ladder-major, each ladder's block row-major
 * (the R side reads ladder l as t(matrix(x[block l], n_monitors, rows))). */
SEXP aptR_run(SEXP h, SEXP niter, SEXP flags /* logical(7): reset resetMV resetTempering simulateAll time adaptTemps printTemps */,
              SEXP tune /* numeric(2) */, SEXP thins /* integer(2) */, SEXP samplesOut) {
    apt_run_args a;
    apt_run_args_default(&a);
    a.niter = (int64_t)Rf_asReal(niter);
    int* f = LOGICAL(flags);
    a.reset = f[0]; a.resetMV = f[1]; a.resetTempering = f[2]; a.simulateAll = f[3]; a.time = f[4]; a.adaptTemps = f[5]; a.printTemps = f[6];
    a.tuneTemper1 = REAL(tune)[0]; a.tuneTemper2 = REAL(tune)[1];
    a.thin = INTEGER(thins)[0]; a.thin2 = INTEGER(thins)[1];
    a.should_abort = should_abort;
    if (samplesOut != R_NilValue) { a.samples_out = REAL(samplesOut); a.samples_out_capacity = (int64_t)XLENGTH(samplesOut); }
    chk(apt_run(R_ExternalPtrAddr(h), &a));
    return R_NilValue;
}

/* what: APT_SAMPLES ... ; returns a numeric matrix rows x cols.  R allocates it; the library writes it in R's own
 * column-major layout (apt_get_layout, APT_COLUMN_MAJOR: transposed on the device), so REAL(out) is passed straight through. */
SEXP aptR_get(SEXP h, SEXP what, SEXP ladder, SEXP rows, SEXP cols) {
    R_xlen_t r = (R_xlen_t)Rf_asReal(rows), c = (R_xlen_t)Rf_asReal(cols);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, (int)r, (int)c));
    int64_t n = 0;
    int rc = apt_get_layout(R_ExternalPtrAddr(h), Rf_asInteger(what), Rf_asInteger(ladder), APT_COLUMN_MAJOR, REAL(out), (int64_t)(r * c), &n);
    if (!rc && n != (int64_t)(r * c)) { UNPROTECT(1); Rf_error("aptR_get: the item has %lld elements, not %lld x %lld", (long long)n, (long long)r, (long long)c); }
    UNPROTECT(1);
    chk(rc);
    return out;
}

/* overwrite the model's current values (the inits of the next reset run): theta numeric(n_params); ladder -1 = all ladders */
SEXP aptR_set_inits(SEXP h, SEXP theta, SEXP ladder) {
    SEXP t = PROTECT(Rf_coerceVector(theta, REALSXP));
    int rc = (int64_t)XLENGTH(t) != apt_info(R_ExternalPtrAddr(h), APT_INFO_N_PARAMS) ? -1 : apt_set_inits(R_ExternalPtrAddr(h), REAL(t), Rf_asInteger(ladder));
    UNPROTECT(1);
    if (rc == -1) Rf_error("aptR_set_inits: theta must have one value per parameter");
    chk(rc);
    return R_NilValue;
}

/* re-upload a constant / data array of the same shape (model$setData for a resident array) */
SEXP aptR_set_array(SEXP h, SEXP name, SEXP x) {
    SEXP xr = PROTECT(Rf_coerceVector(x, REALSXP));
    int rc = apt_set_array(R_ExternalPtrAddr(h), CHAR(STRING_ELT(name, 0)), REAL(xr), (int64_t)XLENGTH(xr));
    UNPROTECT(1);
    chk(rc);
    return R_NilValue;
}

/* model$calculate() at given states: thetaT is an n_params x n matrix (one state per COLUMN, i.e. t(theta) of the usual
 * n x n_params layout, which makes R's column-major storage the [n][n_params] the library takes); returns an
 * (n_groups + 1) x n matrix: row 1 the total logProb, the other rows the per-declaration sums */
SEXP aptR_logprob(SEXP h, SEXP thetaT) {
    SEXP dim = Rf_getAttrib(thetaT, R_DimSymbol);
    const int64_t P = apt_info(R_ExternalPtrAddr(h), APT_INFO_N_PARAMS), ND = apt_info(R_ExternalPtrAddr(h), APT_INFO_N_LOGPROB_GROUPS);
    if (dim == R_NilValue || LENGTH(dim) != 2 || INTEGER(dim)[0] != P) Rf_error("aptR_logprob: thetaT must be an n_params x n matrix");
    const int n = INTEGER(dim)[1];
    SEXP t = PROTECT(Rf_coerceVector(thetaT, REALSXP));
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, (int)ND + 1, n));
    double* tot = (double*)R_alloc((size_t)(n > 0 ? n : 1), sizeof(double));
    double* grp = (double*)R_alloc((size_t)(n > 0 ? n * ND : 1), sizeof(double));
    int rc = apt_logprob(R_ExternalPtrAddr(h), REAL(t), n, tot, grp);
    if (!rc) for (int i = 0; i < n; i++) { REAL(out)[(size_t)i * (ND + 1)] = tot[i]; for (int d = 0; d < ND; d++) REAL(out)[(size_t)i * (ND + 1) + 1 + d] = grp[(size_t)i * ND + d]; }
    UNPROTECT(2);
    chk(rc);
    return out;
}

SEXP aptR_info(SEXP h, SEXP what) { return Rf_ScalarReal((double)apt_info(R_ExternalPtrAddr(h), Rf_asInteger(what))); }

/* cAPT$calculateWAIC(nburnin)  (APT:593-635): numeric(3) = WAIC, lppd, pWAIC */
SEXP aptR_waic(SEXP h, SEXP nburnin, SEXP ladder) {
    SEXP out = PROTECT(Rf_allocVector(REALSXP, 3));
    int rc = apt_waic(R_ExternalPtrAddr(h), (int64_t)Rf_asReal(nburnin), Rf_asInteger(ladder), REAL(out));
    UNPROTECT(1);
    chk(rc);
    return out;
}
