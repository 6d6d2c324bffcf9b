/* oracle/shim/R.h -- TEST INFRASTRUCTURE ONLY.
 * Minimal stand-in for the R C API so that the reference's src/chebpol.c can
 * be compiled, unmodified and in place, without an R installation.  Only the
 * declarations the reference TU names are provided.  The per-point numeric
 * kernels (C_evalcheb, C_evalmlip, FH, fhweights, chebcoef) never touch the R
 * API; every accessor declared here is a link-time stub in ref_wrap.c that
 * aborts if it is ever executed.
 */
#ifndef CHEBPOL_ORACLE_SHIM_R_H
#define CHEBPOL_ORACLE_SHIM_R_H
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

typedef struct shim_sexprec *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
#define R_INLINE inline
#define DOUBLE_EPS DBL_EPSILON

/* SEXP type tags actually named by the reference */
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define VECSXP 19
#define STRSXP 16

extern SEXP R_NilValue, R_DimSymbol, R_DimNamesSymbol, R_BaseEnv, R_NamesSymbol;
extern double R_NaReal;
#define NA_REAL R_NaReal

double *REAL(SEXP);
int *INTEGER(SEXP);
int *LOGICAL(SEXP);
int LENGTH(SEXP);
R_xlen_t XLENGTH(SEXP);
int Rf_length(SEXP);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
int Rf_isNull(SEXP);
int Rf_isLogical(SEXP);
int Rf_isMatrix(SEXP);
int Rf_isFunction(SEXP);
int Rf_isReal(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
SEXP Rf_allocVector(unsigned int, R_xlen_t);
SEXP Rf_allocMatrix(unsigned int, int, int);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP STRING_ELT(SEXP, R_xlen_t);
const char *CHAR(SEXP);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
/* external pointers (declarations only: used by the compile check of chebpol_b200/csrc/r_glue.c) */
typedef void (*R_CFinalizer_t)(SEXP);
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot);
void *R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
SEXP Rf_mkChar(const char *);
#define mkChar Rf_mkChar
SEXP Rf_coerceVector(SEXP, unsigned int);
SEXP Rf_eval(SEXP, SEXP);
SEXP Rf_lang2(SEXP, SEXP);
SEXP SETCADR(SEXP, SEXP);
SEXP Rf_ScalarLogical(int);
SEXP Rf_ScalarReal(double);
int MAYBE_REFERENCED(SEXP);
char *R_alloc(size_t, int);
void Rf_error(const char *, ...) __attribute__((noreturn));
void Rf_warning(const char *, ...);

/* the un-prefixed spellings the reference uses */
#define length Rf_length
#define getAttrib Rf_getAttrib
#define setAttrib Rf_setAttrib
#define isNull Rf_isNull
#define isLogical Rf_isLogical
#define isMatrix Rf_isMatrix
#define isFunction Rf_isFunction
#define nrows Rf_nrows
#define ncols Rf_ncols
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
#define allocVector Rf_allocVector
#define allocMatrix Rf_allocMatrix
#define coerceVector Rf_coerceVector
#define eval Rf_eval
#define lang2 Rf_lang2
#define ScalarLogical Rf_ScalarLogical
#define ScalarReal Rf_ScalarReal
#define error Rf_error
#define warning Rf_warning
#endif
