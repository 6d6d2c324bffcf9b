/* oracle/shim/config.h -- TEST INFRASTRUCTURE ONLY.
 * Stands in for the autoconf output of the reference (src/config.h.in):
 * no FFTW, no GSL, no ALGLIB, so chebcoef takes its BLAS-matrix path
 * (src/chebpol.c:65-122) and the ALGLIB stubs (:926-941) are compiled. */
