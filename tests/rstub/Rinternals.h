/* Stand-in for R's C API (test infrastructure): this image has no R, so the .Call bridge
 * (nimbleapt_b200/R/bridge.c) is compiled and EXECUTED against this minimal re-declaration of the handful of entry
 * points it uses.  Semantics follow "Writing R Extensions" for those entry points only; nothing here comes from R's
 * sources.  Implemented in rstub.c. */
#ifndef APT_RSTUB_RINTERNALS_H
#define APT_RSTUB_RINTERNALS_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct stub_sexp* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
enum { NILSXP = 0, LGLSXP = 10, INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19, EXTPTRSXP = 22, CHARSXP = 9, SYMSXP = 1 };
extern SEXP R_NilValue, R_DimSymbol, R_NamesSymbol;
double* REAL(SEXP);
int* INTEGER(SEXP);
int* LOGICAL(SEXP);
const char* CHAR(SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
int LENGTH(SEXP);
R_xlen_t XLENGTH(SEXP);
SEXP PROTECT(SEXP);
void UNPROTECT(int);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_coerceVector(SEXP, int);
SEXP Rf_allocVector(int, R_xlen_t);
SEXP Rf_allocMatrix(int, int, int);
SEXP Rf_mkString(const char*);
SEXP Rf_ScalarReal(double);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
int Rf_isNumeric(SEXP);
void Rf_error(const char*, ...) __attribute__((noreturn));
char* R_alloc(size_t, int);
SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
void* R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
Rboolean R_ToplevelExec(void (*fun)(void*), void* data);
void R_CheckUserInterrupt(void);
#ifdef __cplusplus
}
#endif
#endif
