/* oracle/shim/Rmath.h -- TEST INFRASTRUCTURE ONLY (see R.h). */
#ifndef CHEBPOL_ORACLE_SHIM_RMATH_H
#define CHEBPOL_ORACLE_SHIM_RMATH_H
double cospi(double);
double R_pow_di(double, int);
double Rf_sign(double);
#define sign Rf_sign
#endif
