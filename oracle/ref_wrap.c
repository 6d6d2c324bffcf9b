/* oracle/ref_wrap.c -- TEST INFRASTRUCTURE ONLY; never linked into the product.
 *
 * Builds oracle/_ref/libchebpol_ref.so from the reference's OWN source file,
 * compiled where it lies (the Makefile passes -I$(REF)/src; nothing is copied
 * into this repository).  Because this TU #includes the reference .c file,
 * its `static` per-point kernels are reachable directly:
 *   C_evalcheb  src/chebpol.c:314-338     FH         src/chebpol.c:182-218
 *   C_evalmlip  src/chebpol.c:617-667     fhweights  src/chebpol.c:267-282
 *   chebcoef    src/chebpol.c:8-141 (BLAS-matrix path, :65-122)
 * The batch drivers below repeat the reference's outer loops
 * (src/chebpol.c:375-378, :701-704, :259-262) with 64-bit point indices
 * (the reference's `int i`, `i*rank` overflow once P*K >= 2^31).
 *
 * The R API is a set of stubs that abort when executed -- the numeric kernels never call
 * into R -- except for a minimal working vector object (REAL/INTEGER/nrows/ncols/VECTOR_ELT/
 * allocVector ...) that lets the two simplex-linear wrappers at the end run the reference's
 * real .Call entry points, and the LAPACK/BLAS routines restated below.
 */
#include <stdio.h>
#include <stdarg.h>
#include <stdint.h>
#include "chebpol.c" /* resolved through -I/root/reference/src at build time */
#include "stalker.c" /* the reference's second TU: makehyp/evalhyp, makestalk/evalstalk (no GSL here) */

/* ---- R API link-time stubs (never executed on the paths we call) ---- */
#undef error
#undef warning
#undef length
#define SHIM_DIE(name) do { fprintf(stderr, "oracle shim: R API '%s' called\n", name); abort(); } while (0)
SEXP R_NilValue, R_DimSymbol, R_DimNamesSymbol, R_BaseEnv;
double R_NaReal;
/* R's NA_real_: a quiet NaN whose low word is 1954 (R src/main/arithmetic.c, R_NaReal) */
__attribute__((constructor)) static void shim_init_na(void) {
  union { double d; unsigned long long u; } v; v.u = 0x7FF00000000007A2ULL; R_NaReal = v.d;
}
/* A minimal working vector object, used ONLY by the two simplex-linear wrappers at the end of
 * this file, which run the reference's real .Call entry points R_analyzesimplex / R_evalsl
 * (src/chebpol.c:803-875): type, length, matrix shape and a data pointer (borrowed from the
 * caller or owned).  Every other wrapper passes plain C pointers to the static kernels and
 * never reaches these accessors. */
struct shim_sexprec { unsigned int type; R_xlen_t len; int nrow, ncol; void *data; int owned; };
static SEXP shim_wrap(unsigned int type, R_xlen_t len, int nrow, int ncol, void *data) {
  SEXP s = calloc(1, sizeof *s); s->type = type; s->len = len; s->nrow = nrow; s->ncol = ncol; s->data = data; return s;
}
static void shim_free(SEXP s) {
  if (!s) return;
  if (s->type == VECSXP && s->data) for (R_xlen_t i = 0; i < s->len; i++) shim_free(((SEXP *)s->data)[i]);
  if (s->owned) free(s->data);
  free(s);
}
double *REAL(SEXP s) { if (!s || s->type != REALSXP) SHIM_DIE("REAL"); return s->data; }
int *INTEGER(SEXP s) { if (!s || s->type != INTSXP) SHIM_DIE("INTEGER"); return s->data; }
int *LOGICAL(SEXP s) { if (!s || s->type != LGLSXP) SHIM_DIE("LOGICAL"); return s->data; }
int LENGTH(SEXP s) { if (!s) SHIM_DIE("LENGTH"); return (int)s->len; }
R_xlen_t XLENGTH(SEXP s) { if (!s) SHIM_DIE("XLENGTH"); return s->len; }
int Rf_length(SEXP s) { if (!s) SHIM_DIE("length"); return (int)s->len; }
SEXP Rf_getAttrib(SEXP a, SEXP b) { (void)a; (void)b; SHIM_DIE("getAttrib"); }
SEXP Rf_setAttrib(SEXP a, SEXP b, SEXP c) { (void)a; (void)b; (void)c; SHIM_DIE("setAttrib"); }
int Rf_isNull(SEXP s) { return s == NULL; } /* R_NilValue is the NULL pointer here */
int Rf_isLogical(SEXP s) { (void)s; SHIM_DIE("isLogical"); }
int Rf_isMatrix(SEXP s) { (void)s; SHIM_DIE("isMatrix"); }
int Rf_isFunction(SEXP s) { (void)s; SHIM_DIE("isFunction"); }
int Rf_isReal(SEXP s) { (void)s; SHIM_DIE("isReal"); }
int Rf_nrows(SEXP s) { if (!s || s->nrow < 0) SHIM_DIE("nrows"); return s->nrow; }
int Rf_ncols(SEXP s) { if (!s || s->ncol < 0) SHIM_DIE("ncols"); return s->ncol; }
SEXP Rf_protect(SEXP s) { return s; }
void Rf_unprotect(int n) { (void)n; }
SEXP Rf_allocVector(unsigned int t, R_xlen_t n) {
  size_t el = (t == REALSXP) ? sizeof(double) : (t == VECSXP) ? sizeof(SEXP) : sizeof(int);
  SEXP s = shim_wrap(t, n, -1, -1, calloc(n > 0 ? n : 1, el)); s->owned = 1; return s;
}
SEXP Rf_allocMatrix(unsigned int t, int a, int b) {
  SEXP s = Rf_allocVector(t, (R_xlen_t)a * b); s->nrow = a; s->ncol = b; return s;
}
SEXP VECTOR_ELT(SEXP s, R_xlen_t i) { if (!s || s->type != VECSXP || i >= s->len) SHIM_DIE("VECTOR_ELT"); return ((SEXP *)s->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP s, R_xlen_t i, SEXP v) { if (!s || s->type != VECSXP || i >= s->len) SHIM_DIE("SET_VECTOR_ELT"); ((SEXP *)s->data)[i] = v; return v; }
/* AS_INTEGER / AS_LOGICAL of an object that already has the type: the object itself */
SEXP Rf_coerceVector(SEXP s, unsigned int t) { if (!s || s->type != t) SHIM_DIE("coerceVector"); return s; }
SEXP Rf_eval(SEXP a, SEXP b) { (void)a; (void)b; SHIM_DIE("eval"); }
SEXP Rf_lang2(SEXP a, SEXP b) { (void)a; (void)b; SHIM_DIE("lang2"); }
SEXP SETCADR(SEXP a, SEXP b) { (void)a; (void)b; SHIM_DIE("SETCADR"); }
SEXP Rf_ScalarLogical(int v) { (void)v; SHIM_DIE("ScalarLogical"); }
SEXP Rf_ScalarReal(double v) { (void)v; SHIM_DIE("ScalarReal"); }
int MAYBE_REFERENCED(SEXP s) { (void)s; SHIM_DIE("MAYBE_REFERENCED"); }
int R_registerRoutines(DllInfo *d, const void *a, const R_CallMethodDef *b, const void *c, const void *e)
{ (void)d; (void)a; (void)b; (void)c; (void)e; SHIM_DIE("R_registerRoutines"); }
int R_useDynamicSymbols(DllInfo *d, int v) { (void)d; (void)v; SHIM_DIE("R_useDynamicSymbols"); }
void R_RegisterCCallable(const char *a, const char *b, DL_FUNC f) { (void)a; (void)b; (void)f; SHIM_DIE("R_RegisterCCallable"); }
void Rf_error(const char *fmt, ...) {
  va_list ap; va_start(ap, fmt); fprintf(stderr, "oracle shim: error(): "); vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n"); va_end(ap); abort();
}
void Rf_warning(const char *fmt, ...) {
  va_list ap; va_start(ap, fmt); fprintf(stderr, "oracle shim: warning(): "); vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n"); va_end(ap);
}
void SET_STRING_ELT(SEXP s, R_xlen_t i, SEXP v) { (void)s; (void)i; (void)v; SHIM_DIE("SET_STRING_ELT"); }
SEXP Rf_mkChar(const char *c) { (void)c; SHIM_DIE("mkChar"); }
SEXP R_NamesSymbol;
/* R's sign() (nmath/sign.c) and findInterval() (appl/interv.c) are not vendored under
 * /root/reference; restated from their documented semantics.  findInterval is only called with
 * rightmost_closed = all_inside = TRUE (src/stalker.c:382,553,633): the 1-based index i with
 * xt[i-1] <= x < xt[i], clamped to 1..n-1. */
double Rf_sign(double x) { if (isnan(x)) return x; return (x > 0) ? 1.0 : ((x == 0) ? 0.0 : -1.0); }
int findInterval(double *xt, int n, double x, Rboolean rightmost_closed, Rboolean all_inside, int ilo, int *mflag)
{
  (void)rightmost_closed; (void)all_inside; (void)ilo;
  int lo = 0, hi = n; /* count of knots <= x by bisection */
  while (lo < hi) { int mid = (lo + hi) / 2; if (xt[mid] <= x) lo = mid + 1; else hi = mid; }
  *mflag = (lo == 0) ? -1 : ((lo == n) ? 1 : 0);
  if (lo < 1) lo = 1;
  if (lo > n - 1) lo = n - 1;
  return lo;
}
/* LAPACK dgetrf / dgetrs (and the BLAS they call) are not vendored under /root/reference and no
 * LAPACK is installed here.  Restated from the published reference (netlib) algorithm: partial
 * pivoting with the first largest |a| (idamax), row interchange, the multipliers scaled by the
 * reciprocal of the pivot (dgetf2 / dgetrf2: dscal by 1/pivot when |pivot| >= sfmin), rank-one
 * update of the trailing block (dger); dgetrs 'N': dlaswp, unit-lower then upper dtrsm, each
 * skipping zero right-hand-side entries.  The recursive dgetrf2 of LAPACK >= 3.6 performs the
 * same elementary operations on every entry in the same order, so it rounds identically. */
void dgetrf_(const int *M, const int *N, double *A, const int *lda, int *ipiv, int *info)
{
  const int m = *M, n = *N, ld = *lda, mn = m < n ? m : n;
  *info = 0;
  for (int j = 0; j < mn; j++) {
    int jp = j; double mx = fabs(A[j + (size_t)j * ld]);
    for (int i = j + 1; i < m; i++) { const double v = fabs(A[i + (size_t)j * ld]); if (v > mx) { mx = v; jp = i; } }
    ipiv[j] = jp + 1;
    if (A[jp + (size_t)j * ld] != 0.0) {
      if (jp != j) for (int c = 0; c < n; c++) { double t = A[j + (size_t)c * ld]; A[j + (size_t)c * ld] = A[jp + (size_t)c * ld]; A[jp + (size_t)c * ld] = t; }
      if (j < m - 1) {
        if (fabs(A[j + (size_t)j * ld]) >= DBL_MIN) { const double r = 1.0 / A[j + (size_t)j * ld]; for (int i = j + 1; i < m; i++) A[i + (size_t)j * ld] *= r; }
        else for (int i = j + 1; i < m; i++) A[i + (size_t)j * ld] /= A[j + (size_t)j * ld];
      }
    } else if (*info == 0) *info = j + 1;
    if (j < mn - 1)
      for (int c = j + 1; c < n; c++) {
        const double y = A[j + (size_t)c * ld];
        if (y != 0.0) { const double t = -y; for (int i = j + 1; i < m; i++) A[i + (size_t)c * ld] += A[i + (size_t)j * ld] * t; }
      }
  }
}
void dgetrs_(const char *tr, const int *N, const int *nrhs, const double *A, const int *lda, const int *ipiv, double *B, const int *ldb, int *info)
{
  const int n = *N, ld = *lda;
  if (tr[0] != 'N') SHIM_DIE("dgetrs (trans != N)");
  *info = 0;
  for (int r = 0; r < *nrhs; r++) {
    double *b = B + (size_t)r * *ldb;
    for (int i = 0; i < n; i++) { const int ip = ipiv[i] - 1; if (ip != i) { double t = b[i]; b[i] = b[ip]; b[ip] = t; } }
    for (int k = 0; k < n; k++) if (b[k] != 0.0) for (int i = k + 1; i < n; i++) b[i] = b[i] - b[k] * A[i + (size_t)k * ld];
    for (int k = n - 1; k >= 0; k--) if (b[k] != 0.0) { b[k] = b[k] / A[k + (size_t)k * ld]; for (int i = 0; i < k; i++) b[i] = b[i] - b[k] * A[i + (size_t)k * ld]; }
  }
}

/* ---- the three non-stub services chebcoef's matrix path needs ---- */
/* R_alloc: transient storage R frees at .Call exit; here a small arena the
 * wrapper resets after each call. */
static char **arena = NULL; static int arena_n = 0, arena_cap = 0;
char *R_alloc(size_t n, int size) {
  if (arena_n == arena_cap) { arena_cap = arena_cap ? 2 * arena_cap : 16; arena = realloc(arena, arena_cap * sizeof(char *)); }
  char *p = malloc(n * (size_t)size + 16); arena[arena_n++] = p; return p;
}
static void arena_reset(void) { for (int i = 0; i < arena_n; i++) free(arena[i]); arena_n = 0; }
/* R's cospi(x) = cos(pi*x) with exact values at multiples of 1/2 */
double cospi(double x) {
  x = fmod(fabs(x), 2.0);
  if (fmod(x, 1.0) == 0.5) return 0.0;
  if (x == 1.0) return -1.0;
  if (x == 0.0) return 1.0;
  return cos(M_PI * x);
}
/* R's R_pow_di (R >= 3.x, src/main/arithmetic.c): x^n by repeated squaring -- a dependency of the
 * reference that is not vendored under /root/reference; restated from the published R source so the
 * polyharmonic basis r^k (src/chebpol.c:402,410) rounds as it does inside R. */
double R_pow_di(double x, int n) {
  double xn = 1.0;
  if (isnan(x)) return x;
  if (n != 0) {
    if (!isfinite(x)) return pow(x, (double)n);
    int is_neg = (n < 0);
    if (is_neg) n = -n;
    for (;;) {
      if (n & 01) xn *= x;
      if (n >>= 1) x *= x; else break;
    }
    if (is_neg) xn = 1. / xn;
  }
  return xn;
}
/* dgemm, only the ("T","T") case reached from src/chebpol.c:119:
 * C(MxN) = alpha * A^T * B^T + beta*C, column-major, A is KxM (lda), B is NxK (ldb) */
void dgemm_(const char *ta, const char *tb, const int *M, const int *N, const int *K,
            const double *alpha, const double *A, const int *lda, const double *B, const int *ldb,
            const double *beta, double *C, const int *ldc) {
  if (ta[0] != 'T' || tb[0] != 'T') SHIM_DIE("dgemm (non T,T)");
  for (int j = 0; j < *N; j++)
    for (int i = 0; i < *M; i++) {
      double s = 0.0;
      for (int k = 0; k < *K; k++) s += A[k + (size_t)i * *lda] * B[j + (size_t)k * *ldb];
      C[i + (size_t)j * *ldc] = *alpha * s + (*beta == 0.0 ? 0.0 : *beta * C[i + (size_t)j * *ldc]);
    }
}

/* ---- exported batch drivers over the reference's own kernels ---- */
#define EXPORT __attribute__((visibility("default")))

EXPORT void ref_evalcheb(const double *cf, const int *dims, int rank, const double *x,
                         int64_t npts, int threads, double *out) {
#pragma omp parallel for num_threads(threads) schedule(static) if (npts > 1 && threads > 1)
  for (int64_t i = 0; i < npts; i++)
    out[i] = C_evalcheb((double *)cf, (double *)x + i * rank, (int *)dims, rank);
}

EXPORT void ref_evalmlip(int rank, const double *const *grid, const int *dims, const double *values,
                         const double *x, int64_t npts, int blend, int threads, double *out) {
#pragma omp parallel for num_threads(threads) schedule(static) if (npts > 1 && threads > 1)
  for (int64_t i = 0; i < npts; i++)
    out[i] = C_evalmlip(rank, (double *)x + i * rank, (double **)grid, (int *)dims, (double *)values, blend);
}

EXPORT void ref_fh(const double *vals, const double *x, const double *const *knots, const int *dims,
                   int rank, const double *const *weights, int64_t npts, int threads, double *out) {
#pragma omp parallel for num_threads(threads) schedule(static) if (npts > 1 && threads > 1)
  for (int64_t i = 0; i < npts; i++)
    out[i] = FH((double *)vals, (double *)x + i * rank, (double **)knots, (int *)dims, rank, (double **)weights);
}

/* the loop of R_evalpolyh (src/chebpol.c:443-446) over the reference's C_evalpolyh (:383-415) */
EXPORT void ref_evalpolyh(const double *x, int64_t npts, const double *knots, const double *weights,
                          const double *lweights, int rank, int nknots, double k, int threads, double *out) {
#pragma omp parallel for num_threads(threads) schedule(static) if (npts > 1 && threads > 1)
  for (int64_t i = 0; i < npts; i++)
    out[i] = C_evalpolyh(x + i * rank, knots, weights, lweights, rank, nknots, k);
}

/* stalker splines: the fit (makestalk :271-364, makehyp :446-535) and the point loops of
 * R_evalstalk (:789-792) and R_evalhyp (:824-827) over evalstalk (:366-442) / evalhyp (:537-613).
 * Tables are rank x prod(dims) column-major, as R_makestalk / R_makehyp allocate them. */
EXPORT void ref_makestalk(int rank, const int *dims, const double *const *grid, const double *val,
                          double *b, double *c, double *r) {
  int point[rank > 0 ? rank : 1];
  makestalk(rank, dims, (double **)grid, val, b, c, r, 0, point);
}
EXPORT void ref_makehyp(int rank, const int *dims, const double *const *grid, const double *val,
                        double *a, double *b, double *c, double *d) {
  int point[rank > 0 ? rank : 1];
  makehyp(rank, dims, (double **)grid, val, a, b, c, d, 0, point);
}
EXPORT void ref_evalstalk(const double *x, int64_t npts, int rank, const int *dims, const double *const *grid,
                          const double *val, int blend, const double *b, const double *c, const double *r,
                          int threads, double *out) {
#pragma omp parallel for num_threads(threads) schedule(guided) if (npts > 1 && threads > 1)
  for (int64_t i = 0; i < npts; i++)
    out[i] = evalstalk(x + i * rank, rank, dims, (double **)grid, val, blend, (double *)b, (double *)c, (double *)r);
}
EXPORT void ref_evalhyp(const double *x, int64_t npts, int rank, const int *dims, const double *const *grid,
                        const double *val, int blend, const double *a, const double *b, const double *c,
                        const double *d, int threads, double *out) {
#pragma omp parallel for num_threads(threads) schedule(guided) if (npts > 1 && threads > 1)
  for (int64_t i = 0; i < npts; i++)
    out[i] = evalhyp(x + i * rank, rank, dims, (double **)grid, val, blend, (double *)a, (double *)b, (double *)c,
                     (double *)d);
}

EXPORT void ref_chebcoef(const double *val, const int *dims, int rank, double *F, int dct) {
  chebcoef((double *)val, (int *)dims, rank, F, dct, 1);
  arena_reset();
}

/* n = number of knots (the reference is called with dims[r]-1, src/chebpol.c:307) */
EXPORT void ref_fhweights(int nknots, const double *grid, int d, double *w) {
  fhweights(nknots - 1, grid, d, w, 1);
}

EXPORT double ref_blendfun(double w, int blend) { return blendfun(w, blend); }
EXPORT int ref_binsearch(int n, const double *arr, double val) { return binsearch(n, (double *)arr, val); }

/* Simplex-linear interpolation (slappx.real, R/simplexlinear.R): the reference's real .Call
 * entry points, R_analyzesimplex (src/chebpol.c:830-875) and R_evalsl (:803-828) over
 * findsimplex (:738-801), run on caller buffers wrapped as vector objects.  dtri is the
 * (dim+1) x numsimplex matrix of 1-based knot indices, knots dim x nknots.  R_evalsl indexes
 * points with `int`, so the batch is passed in chunks. */
EXPORT void ref_analyzesimplex(const int *dtri, int numsimplex, const double *knots, int dim, int nknots,
                               double *lumats, int *pivots, double *bbox) {
  int one = 1;
  SEXP Sd = shim_wrap(INTSXP, (R_xlen_t)(dim + 1) * numsimplex, dim + 1, numsimplex, (void *)dtri);
  SEXP Sk = shim_wrap(REALSXP, (R_xlen_t)dim * nknots, dim, nknots, (void *)knots);
  SEXP St = shim_wrap(INTSXP, 1, -1, -1, &one);
  SEXP res = R_analyzesimplex(Sd, Sk, St);
  const int N = dim + 1;
  memcpy(lumats, REAL(VECTOR_ELT(res, 0)), sizeof(double) * (size_t)numsimplex * N * N);
  memcpy(pivots, INTEGER(VECTOR_ELT(res, 1)), sizeof(int) * (size_t)numsimplex * N);
  memcpy(bbox, REAL(VECTOR_ELT(res, 2)), sizeof(double) * (size_t)numsimplex * 2 * dim);
  shim_free(res); shim_free(Sd); shim_free(Sk); shim_free(St);
}
EXPORT void ref_evalsl(const double *x, int64_t npts, const double *knots, int dim, int nknots, const int *dtri,
                       int numsimplex, const double *lumats, const int *pivots, const double *bbox,
                       const double *val, int epol, int threads, int blend, double *out) {
  const int N = dim + 1;
  SEXP Sk = shim_wrap(REALSXP, (R_xlen_t)dim * nknots, dim, nknots, (void *)knots);
  SEXP Sd = shim_wrap(INTSXP, (R_xlen_t)N * numsimplex, N, numsimplex, (void *)dtri);
  SEXP ad = Rf_allocVector(VECSXP, 3);
  SET_VECTOR_ELT(ad, 0, shim_wrap(REALSXP, (R_xlen_t)numsimplex * N * N, -1, -1, (void *)lumats));
  SET_VECTOR_ELT(ad, 1, shim_wrap(INTSXP, (R_xlen_t)numsimplex * N, -1, -1, (void *)pivots));
  SET_VECTOR_ELT(ad, 2, shim_wrap(REALSXP, (R_xlen_t)numsimplex * 2 * dim, 2 * dim, numsimplex, (void *)bbox));
  SEXP Sv = shim_wrap(REALSXP, nknots, -1, -1, (void *)val);
  SEXP Se = shim_wrap(LGLSXP, 1, -1, -1, &epol);
  SEXP St = shim_wrap(INTSXP, 1, -1, -1, &threads);
  SEXP Sb = shim_wrap(INTSXP, 1, -1, -1, &blend);
  const int64_t chunk = 1 << 24;
  for (int64_t lo = 0; lo < npts; lo += chunk) {
    const int64_t n = (npts - lo < chunk) ? npts - lo : chunk;
    SEXP Sx = shim_wrap(REALSXP, (R_xlen_t)dim * n, dim, (int)n, (void *)(x + lo * dim));
    SEXP r = R_evalsl(Sx, Sk, Sd, ad, Sv, Se, St, Sb);
    memcpy(out + lo, REAL(r), sizeof(double) * (size_t)n);
    shim_free(r); shim_free(Sx);
  }
  shim_free(ad); shim_free(Sk); shim_free(Sd); shim_free(Sv); shim_free(Se); shim_free(St); shim_free(Sb);
}

/* The deprecated recursive stalker evaluator (ipol(method='oldstalker'), R/stalker.R:11-49): the point loop
 * of R_evalstalker (src/stalker.c:881-884) over the reference's evalstalker (:617-678).  det, pmin, pplus
 * are the per-dimension vectors of stalkercontext (R/stalker.R:2-10); this build has no GSL, so a NaN
 * degree is only legal on a uniform grid (:871-874). */
EXPORT void ref_evalstalker(const double *x, int64_t npts, int rank, const int *dims, const double *const *grid,
                            const double *val, int blend, const double *degree, const double *const *det,
                            const double *const *pmin, const double *const *pplus, const int *uniform, int threads,
                            double *out) {
  precomp pc[rank > 0 ? rank : 1];
  for (int i = 0; i < rank; i++) {
    pc[i].det = (double *)det[i]; pc[i].pmin = (double *)pmin[i]; pc[i].pplus = (double *)pplus[i]; pc[i].uniform = uniform[i];
  }
#pragma omp parallel for num_threads(threads) schedule(guided) if (npts > 1 && threads > 1)
  for (int64_t i = 0; i < npts; i++)
    out[i] = evalstalker(x + i * rank, rank, dims, (double **)grid, val, blend, degree, pc);
}
