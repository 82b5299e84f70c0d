/* rstub.c — a few dozen lines standing in for R's runtime so that nimbleapt_b200/R/bridge.c can be compiled and run
 * in an image without R (test infrastructure; see Rinternals.h here).  Objects are never freed before stub_reset():
 * there is no garbage collector, PROTECT / UNPROTECT only keep a balance that the tests check.  Rf_error() records its
 * message and longjmps to the innermost stub_call(), as R's error() unwinds to top level. */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "Rinternals.h"

struct stub_sexp {
    int type;
    R_xlen_t n;
    void* data;            /* double / int / SEXP array, char string, or external address */
    SEXP names, dim;
    R_CFinalizer_t fin;
};
static struct stub_sexp nil_obj = {NILSXP, 0, NULL, NULL, NULL, NULL}, dim_sym = {SYMSXP}, names_sym = {SYMSXP};
SEXP R_NilValue = &nil_obj, R_DimSymbol = &dim_sym, R_NamesSymbol = &names_sym;

static void* pool[1 << 16];
static int npool = 0, protect_depth = 0, interrupt_pending = 0;
static SEXP extptrs[256];
static int nextptr = 0;
static char errmsg[4096];
static jmp_buf* cur_jmp = NULL;

static void* keep(void* p) { if (npool < (1 << 16)) pool[npool++] = p; return p; }
static SEXP mk(int type, R_xlen_t n, size_t elt) {
    SEXP s = (SEXP)keep(calloc(1, sizeof *s));
    s->type = type; s->n = n; s->names = s->dim = R_NilValue;
    s->data = elt ? keep(calloc(n > 0 ? (size_t)n : 1, elt)) : NULL;
    return s;
}
double* REAL(SEXP s) { if (s->type != REALSXP) Rf_error("REAL() on a non-double"); return (double*)s->data; }
int* INTEGER(SEXP s) { if (s->type != INTSXP && s->type != LGLSXP) Rf_error("INTEGER() on a non-integer"); return (int*)s->data; }
int* LOGICAL(SEXP s) { if (s->type != LGLSXP) Rf_error("LOGICAL() on a non-logical"); return (int*)s->data; }
const char* CHAR(SEXP s) { return (const char*)s->data; }
SEXP STRING_ELT(SEXP s, R_xlen_t i) { if (s->type != STRSXP || i >= s->n) Rf_error("STRING_ELT out of range"); return ((SEXP*)s->data)[i]; }
SEXP VECTOR_ELT(SEXP s, R_xlen_t i) { if (s->type != VECSXP || i >= s->n) Rf_error("VECTOR_ELT out of range"); return ((SEXP*)s->data)[i]; }
int LENGTH(SEXP s) { return (int)s->n; }
R_xlen_t XLENGTH(SEXP s) { return s->n; }
SEXP PROTECT(SEXP s) { protect_depth++; return s; }
void UNPROTECT(int n) { protect_depth -= n; }
SEXP Rf_getAttrib(SEXP s, SEXP sym) { return sym == R_DimSymbol ? s->dim : sym == R_NamesSymbol ? s->names : R_NilValue; }
SEXP Rf_allocVector(int type, R_xlen_t n) { return mk(type, n, type == REALSXP ? sizeof(double) : (type == INTSXP || type == LGLSXP) ? sizeof(int) : sizeof(SEXP)); }
SEXP Rf_allocMatrix(int type, int r, int c) {
    SEXP s = Rf_allocVector(type, (R_xlen_t)r * c);
    s->dim = Rf_allocVector(INTSXP, 2);
    INTEGER(s->dim)[0] = r; INTEGER(s->dim)[1] = c;
    return s;
}
static SEXP mkchar(const char* c) { SEXP s = mk(CHARSXP, (R_xlen_t)strlen(c), 0); s->data = keep(strdup(c)); return s; }
SEXP Rf_mkString(const char* c) { SEXP s = Rf_allocVector(STRSXP, 1); ((SEXP*)s->data)[0] = mkchar(c); return s; }
SEXP Rf_ScalarReal(double x) { SEXP s = Rf_allocVector(REALSXP, 1); REAL(s)[0] = x; return s; }
SEXP Rf_coerceVector(SEXP s, int type) {
    if (s->type == type) return s;
    if (type != REALSXP || (s->type != INTSXP && s->type != LGLSXP)) Rf_error("stub: unsupported coercion");
    SEXP r = Rf_allocVector(REALSXP, s->n);
    for (R_xlen_t i = 0; i < s->n; i++) REAL(r)[i] = (double)((int*)s->data)[i];
    r->dim = s->dim; r->names = s->names;
    return r;
}
double Rf_asReal(SEXP s) { if (s->n < 1) return 0.0 / 0.0; return s->type == REALSXP ? ((double*)s->data)[0] : (s->type == INTSXP || s->type == LGLSXP) ? (double)((int*)s->data)[0] : 0.0 / 0.0; }
int Rf_asInteger(SEXP s) { return (int)Rf_asReal(s); }
int Rf_asLogical(SEXP s) { return Rf_asReal(s) != 0.0; }
int Rf_isNumeric(SEXP s) { return s->type == REALSXP || s->type == INTSXP; }
void Rf_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(errmsg, sizeof errmsg, fmt, ap); va_end(ap);
    if (!cur_jmp) { fprintf(stderr, "rstub: error outside stub_call: %s\n", errmsg); abort(); }
    longjmp(*cur_jmp, 1);
}
char* R_alloc(size_t n, int size) { return (char*)keep(calloc(n ? n : 1, (size_t)size)); }
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot) { (void)tag; (void)prot; SEXP s = mk(EXTPTRSXP, 0, 0); s->data = p; if (nextptr < 256) extptrs[nextptr++] = s; return s; }
void* R_ExternalPtrAddr(SEXP s) { if (s->type != EXTPTRSXP) Rf_error("not an external pointer"); return s->data; }
void R_ClearExternalPtr(SEXP s) { s->data = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t f, Rboolean onexit) { (void)onexit; s->fin = f; }
void R_CheckUserInterrupt(void) { if (interrupt_pending) { interrupt_pending = 0; Rf_error("interrupt"); } }
Rboolean R_ToplevelExec(void (*fun)(void*), void* data) {  /* FALSE if fun() raised an error (here: the interrupt) */
    jmp_buf jb, *outer = cur_jmp;
    cur_jmp = &jb;
    int failed = setjmp(jb);
    if (!failed) fun(data);
    cur_jmp = outer;
    return !failed;
}

/* ---- what the test driver (Python, ctypes) uses to stand in for the R interpreter */
SEXP stub_real(const double* v, R_xlen_t n) { SEXP s = Rf_allocVector(REALSXP, n); if (n) memcpy(s->data, v, (size_t)n * sizeof(double)); return s; }
SEXP stub_int(const int* v, R_xlen_t n) { SEXP s = Rf_allocVector(INTSXP, n); if (n) memcpy(s->data, v, (size_t)n * sizeof(int)); return s; }
SEXP stub_lgl(const int* v, R_xlen_t n) { SEXP s = Rf_allocVector(LGLSXP, n); if (n) memcpy(s->data, v, (size_t)n * sizeof(int)); return s; }
SEXP stub_str(const char* const* v, R_xlen_t n) { SEXP s = Rf_allocVector(STRSXP, n); for (R_xlen_t i = 0; i < n; i++) ((SEXP*)s->data)[i] = mkchar(v[i]); return s; }
SEXP stub_list(const char* const* names, const SEXP* elems, R_xlen_t n) {
    SEXP s = Rf_allocVector(VECSXP, n);
    for (R_xlen_t i = 0; i < n; i++) ((SEXP*)s->data)[i] = elems[i];
    s->names = stub_str(names, n);
    return s;
}
void stub_set_dim(SEXP s, const int* d, int nd) { s->dim = stub_int(d, nd); }
SEXP stub_nil(void) { return R_NilValue; }
int stub_type(SEXP s) { return s->type; }
R_xlen_t stub_length(SEXP s) { return s->n; }
const void* stub_data(SEXP s) { return s->data; }
const char* stub_string(SEXP s, R_xlen_t i) { return CHAR(STRING_ELT(s, i)); }
int stub_dim(SEXP s, int i) { return s->dim == R_NilValue ? -1 : INTEGER(s->dim)[i]; }
const char* stub_errmsg(void) { return errmsg; }
int stub_protect_depth(void) { return protect_depth; }
void stub_set_interrupt(int on) { interrupt_pending = on; }
/* .Call(fn, args...): NULL when the callee raised an R error (message in stub_errmsg) */
SEXP stub_call(SEXP (*fn)(), int nargs, SEXP* a) {
    jmp_buf jb, *outer = cur_jmp;
    cur_jmp = &jb;
    errmsg[0] = 0;
    SEXP r = NULL;
    if (!setjmp(jb)) {
        switch (nargs) {
            case 1: r = fn(a[0]); break;
            case 2: r = fn(a[0], a[1]); break;
            case 3: r = fn(a[0], a[1], a[2]); break;
            case 4: r = fn(a[0], a[1], a[2], a[3]); break;
            case 5: r = fn(a[0], a[1], a[2], a[3], a[4]); break;
            case 6: r = fn(a[0], a[1], a[2], a[3], a[4], a[5]); break;
            case 8: r = fn(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); break;
            default: snprintf(errmsg, sizeof errmsg, "stub_call: unsupported arity %d", nargs);
        }
    } else {
        protect_depth = 0; /* R resets the protection stack when an error unwinds to top level */
    }
    cur_jmp = outer;
    return r;
}
/* gc at exit: run the finalizers of all external pointers, free everything */
void stub_reset(void) {
    for (int i = 0; i < nextptr; i++) if (extptrs[i]->fin && extptrs[i]->data) extptrs[i]->fin(extptrs[i]);
    nextptr = 0;
    for (int i = 0; i < npool; i++) free(pool[i]);
    npool = 0; protect_depth = 0;
}
