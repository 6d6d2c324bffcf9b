/* oracle/shim/Rdefines.h -- TEST INFRASTRUCTURE ONLY (see R.h). */
#ifndef CHEBPOL_ORACLE_SHIM_RDEFINES_H
#define CHEBPOL_ORACLE_SHIM_RDEFINES_H
#include "R.h"
#define NEW_NUMERIC(n) Rf_allocVector(REALSXP, n)
#define NEW_INTEGER(n) Rf_allocVector(INTSXP, n)
#define NEW_LIST(n) Rf_allocVector(VECSXP, n)
#define AS_INTEGER(x) Rf_coerceVector(x, INTSXP)
#define AS_LOGICAL(x) Rf_coerceVector(x, LGLSXP)
#define AS_NUMERIC(x) Rf_coerceVector(x, REALSXP)
#define IS_NUMERIC(x) Rf_isReal(x)
#define SET_NAMES(x, n) Rf_setAttrib(x, R_NamesSymbol, n)
#endif
