/* oracle/shim/R_ext/Visibility.h -- TEST INFRASTRUCTURE ONLY (see R.h). */
#ifndef CHEBPOL_ORACLE_SHIM_VISIBILITY_H
#define CHEBPOL_ORACLE_SHIM_VISIBILITY_H
#define attribute_visible __attribute__((visibility("default")))
#define attribute_hidden __attribute__((visibility("hidden")))
#endif
