/* oracle/shim/R_ext/BLAS.h -- TEST INFRASTRUCTURE ONLY (see R.h). */
#ifndef CHEBPOL_ORACLE_SHIM_BLAS_H
#define CHEBPOL_ORACLE_SHIM_BLAS_H
#define F77_CALL(x) x##_
#define F77_NAME(x) x##_
void dgemm_(const char *, const char *, const int *, const int *, const int *,
            const double *, const double *, const int *, const double *, const int *,
            const double *, double *, const int *);
#endif
