/* oracle/shim/R_ext/Lapack.h -- TEST INFRASTRUCTURE ONLY (see R.h). */
#ifndef CHEBPOL_ORACLE_SHIM_LAPACK_H
#define CHEBPOL_ORACLE_SHIM_LAPACK_H
void dgetrf_(const int *, const int *, double *, const int *, int *, int *);
void dgetrs_(const char *, const int *, const int *, const double *, const int *,
             const int *, double *, const int *, int *);
#endif
