/* oracle/shim/R_ext/Applic.h -- TEST INFRASTRUCTURE ONLY (see R.h). */
#ifndef CHEBPOL_ORACLE_SHIM_APPLIC_H
#define CHEBPOL_ORACLE_SHIM_APPLIC_H
#endif
