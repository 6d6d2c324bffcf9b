/* oracle/shim/R_ext/Rdynload.h -- TEST INFRASTRUCTURE ONLY (see R.h). */
#ifndef CHEBPOL_ORACLE_SHIM_RDYNLOAD_H
#define CHEBPOL_ORACLE_SHIM_RDYNLOAD_H
typedef void *(*DL_FUNC)(void);
typedef struct { const char *name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct shim_dllinfo DllInfo;
int R_registerRoutines(DllInfo *, const void *, const R_CallMethodDef *, const void *, const void *);
int R_useDynamicSymbols(DllInfo *, int);
void R_RegisterCCallable(const char *, const char *, DL_FUNC);
#endif
