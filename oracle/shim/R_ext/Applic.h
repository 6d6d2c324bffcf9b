/* oracle/shim/R_ext/Applic.h -- TEST INFRASTRUCTURE ONLY (see R.h). */
#ifndef CHEBPOL_ORACLE_SHIM_APPLIC_H
#define CHEBPOL_ORACLE_SHIM_APPLIC_H
#include "../R.h"
int findInterval(double *xt, int n, double x, Rboolean rightmost_closed, Rboolean all_inside, int ilo, int *mflag);
#endif
