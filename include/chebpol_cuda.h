/* chebpol_cuda.h -- C ABI of libchebpol_cuda, the B200 (sm_100a) batched
 * evaluation engine behind chebpol's interpolant closures.
 *
 * Plain pointers and sizes only; no R, torch or C++ types.  Every function
 * returns 0 on success or a CHEBPOL_CUDA_E* code; nothing longjmps, nothing
 * calls back into R, and there is no CPU fallback: without a usable CUDA
 * device every compute entry point fails with CHEBPOL_CUDA_ENODEV.
 *
 * Conventions shared with the reference (/root/reference/src/chebpol.c):
 *  - tensors (coef / values / vals) are R arrays: column-major, dim 0 fastest;
 *  - points are the K x P column-major matrix R passes to .Call, i.e. point i
 *    occupies x[i*rank .. i*rank+rank-1] (reference: `xp+i*rank`, :377);
 *  - results are a length-P vector of doubles (reference: NEW_NUMERIC(numvec)).
 * Differences, all deliberate: point counts are int64_t (the reference's `int`
 * index overflows once P*K >= 2^31, src/chebpol.c:368,376-377); NaN
 * coordinates give NaN results instead of the endless loop of the reference's
 * binsearch (src/chebpol.c:547-554).
 */
#ifndef CHEBPOL_CUDA_H
#define CHEBPOL_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CHEBPOL_CUDA_API __attribute__((visibility("default")))
#else
#define CHEBPOL_CUDA_API
#endif

#define CHEBPOL_CUDA_MAXRANK 31

/* ---- error codes ------------------------------------------------------- */
#define CHEBPOL_CUDA_OK 0
#define CHEBPOL_CUDA_EINVAL 1  /* NULL pointer / negative count / bad option        */
#define CHEBPOL_CUDA_ERANK 2   /* rank <= 0 or > MAXRANK (ref: "rank must be positive") */
#define CHEBPOL_CUDA_ESIZE 3   /* a dimension <= 0 or product overflows              */
#define CHEBPOL_CUDA_EBLEND 4  /* blend not in 0..5 (src/chebpol.h:24-45)            */
#define CHEBPOL_CUDA_ENODEV 5  /* no CUDA device / requested device absent           */
#define CHEBPOL_CUDA_ENOMEM 6  /* device or pinned-host allocation failed            */
#define CHEBPOL_CUDA_ECUDA 7   /* any other CUDA runtime failure (see last_error)    */
#define CHEBPOL_CUDA_EARCH 8   /* device is not compute capability 10.x              */

CHEBPOL_CUDA_API const char *chebpol_cuda_strerror(int code);
/* detail text of the calling thread's most recent failure ("" if none) */
CHEBPOL_CUDA_API const char *chebpol_cuda_last_error(void);
CHEBPOL_CUDA_API int chebpol_cuda_device_count(int *count);
/* version of this ABI: major*10000 + minor*100 + patch */
CHEBPOL_CUDA_API int chebpol_cuda_version(void);

/* ---- stateless entry points: host pointers in, host pointers out --------
 * These are what the .Call glue binds (csrc/r_glue.c); one call replaces
 * one call of the reference routine named beside it.  The tensor is uploaded
 * (replicated on each GPU used), the point batch is split into contiguous
 * ranges over `ngpus` devices (<=0: all visible), streamed through each GPU
 * in chunks, and every GPU writes its own slice of `out`.  No state survives
 * the call, as in the reference (tensors live in the R closure environment).
 */

/* Replaces R_evalcheb (src/chebpol.c:340-381) + C_evalcheb (:314-338), and
 * the R-level pre-maps that run on every call of the closure:
 *   mid/ispan != NULL : x <- (x - mid)*ispan          imap, R/chebyshev.R:223-227
 *   ugm_n     != NULL : x <- sin(0.5*pi*x*(1-n)/n)    ugm,  R/uniform.R:6,63-73
 *                       (applied after the interval map, as ucappx.real does)
 * dims[rank] are the dims of `coef`; length(coef) == prod(dims). */
CHEBPOL_CUDA_API int chebpol_cuda_evalcheb(const double *coef, const int *dims, int rank,
                          const double *x, int64_t npts,
                          const double *mid, const double *ispan, const int *ugm_n,
                          int ngpus, double *out);

/* The 'general' grid method: replaces the closure of chebappxg.real (R/chebyshev.R:333-359),
 *   function(x, threads) ch(gridmap(x), threads)
 * i.e. R_evalcheb preceded by the R-level grid map that runs on every call: coordinate d goes
 * through gridmaps[[d]] = stats::splinefun(grid[[d]], chebknots(dims)[[d]], method='hyman'),
 * applied per point column with mapply (:351-354, interpreter-bound like ugm).  The tables are
 * the ones the splinefun closure already holds -- environment(gridmaps[[d]])$z: x (sn[d] strictly
 * increasing knots), y, b, c, d -- and the map is evaluated as R's spline_eval does for a scalar
 * call: piece i = 0 if x[0] <= u <= x[1], else the last knot <= u; y + dx*(b + dx*(c + dx*d));
 * the first / last piece extrapolates. */
CHEBPOL_CUDA_API int chebpol_cuda_evalchebg(const double *coef, const int *dims, int rank,
                           const double *x, int64_t npts, const int *sn,
                           const double *const *sx, const double *const *sy, const double *const *sb,
                           const double *const *sc, const double *const *sd, int ngpus, double *out);

/* Replaces R_evalmlip (src/chebpol.c:670-707) + C_evalmlip (:617-667),
 * binsearch (:543-556) and blendfun (src/chebpol.h:24-45).
 * grid[d] has dims[d] strictly increasing knots; blend: 0 linear, 1 sigmoid,
 * 2 parodic, 3 cubic, 4 square/step, 5 mean (R/multilinear.R:64-65). */
CHEBPOL_CUDA_API int chebpol_cuda_evalmlip(const double *const *grid, const int *dims, int rank,
                          const double *values, const double *x, int64_t npts,
                          int blend, int ngpus, double *out);

/* Replaces R_FH (src/chebpol.c:220-265) + FH (:182-218).  weights[d] are the
 * Floater-Hormann weights of grid[d] (reference: C_FHweights, :267-312);
 * grids may be increasing or decreasing. */
CHEBPOL_CUDA_API int chebpol_cuda_fh(const double *vals, const double *const *grid,
                    const double *const *weights, const int *dims, int rank,
                    const double *x, int64_t npts, int ngpus, double *out);

/* ---- plans: a fitted interpolant resident on one GPU --------------------
 * For callers that evaluate the same interpolant repeatedly, and for
 * device-resident benchmarking.  A plan owns the device copy of the tensor
 * (re-laid-out for the kernels), knot/weight vectors and pre-map constants.
 */
typedef struct chebpol_cuda_plan chebpol_cuda_plan;

CHEBPOL_CUDA_API int chebpol_cuda_plan_cheb(const double *coef, const int *dims, int rank,
                           const double *mid, const double *ispan, const int *ugm_n,
                           int device, chebpol_cuda_plan **plan);
/* Chebyshev plan with the 'general' grid maps (see chebpol_cuda_evalchebg) */
CHEBPOL_CUDA_API int chebpol_cuda_plan_chebg(const double *coef, const int *dims, int rank, const int *sn,
                            const double *const *sx, const double *const *sy, const double *const *sb,
                            const double *const *sc, const double *const *sd, int device,
                            chebpol_cuda_plan **plan);
CHEBPOL_CUDA_API int chebpol_cuda_plan_mlip(const double *const *grid, const int *dims, int rank,
                           const double *values, int blend, int device,
                           chebpol_cuda_plan **plan);
CHEBPOL_CUDA_API int chebpol_cuda_plan_fh(const double *vals, const double *const *grid,
                         const double *const *weights, const int *dims, int rank,
                         int device, chebpol_cuda_plan **plan);
CHEBPOL_CUDA_API void chebpol_cuda_plan_destroy(chebpol_cuda_plan *plan);

/* d_x (rank*npts doubles, point-major) and d_out (npts doubles) are DEVICE
 * pointers on the plan's device; `stream` is a cudaStream_t (NULL = default
 * stream).  Asynchronous: returns after enqueueing -- except that a plan may first build a
 * derived table it has not built yet (the multilinear cell-major table on the stream; the
 * simplex-linear candidate lists on the host, which synchronises the device once). */
CHEBPOL_CUDA_API int chebpol_cuda_plan_eval_device(chebpol_cuda_plan *plan, const double *d_x,
                                  int64_t npts, double *d_out, void *stream);
/* Host pointers (pinned or pageable); chunked, double-buffered H2D -> kernel
 * -> D2H pipeline on the plan's device.  Synchronous. */
CHEBPOL_CUDA_API int chebpol_cuda_plan_eval_host(chebpol_cuda_plan *plan, const double *x,
                                int64_t npts, double *out);

/* plan options (set) and facts (get) */
#define CHEBPOL_CUDA_OPT_KERNEL 1      /* see CHEBPOL_CUDA_KERNEL_*                       */
#define CHEBPOL_CUDA_OPT_BLEND 2       /* multilinear plans: blend 0..5                   */
#define CHEBPOL_CUDA_OPT_CHUNK_POINTS 3 /* eval_host chunk size in points (0 = auto)      */
#define CHEBPOL_CUDA_OPT_VARIANT 4     /* which fast kernel / layout (0 = auto), see below  */
#define CHEBPOL_CUDA_OPT_EPOL 5        /* simplex-linear plans: extrapolate outside the hull */
#define CHEBPOL_CUDA_OPT_SL_CELLS 6    /* simplex-linear plans: pin the candidate grid to this many cells
                                          per dimension and build it now (0: back to automatic sizing)  */
#define CHEBPOL_CUDA_GET_KERNEL_USED 101  /* kernel id of the most recent eval            */
#define CHEBPOL_CUDA_GET_LAUNCHES 102     /* kernels launched by this plan so far         */
#define CHEBPOL_CUDA_GET_TENSOR_BYTES 103 /* device bytes of the resident tensor          */
#define CHEBPOL_CUDA_GET_DEVICE 104
#define CHEBPOL_CUDA_GET_SMEM_BYTES 105   /* dynamic smem of the most recent launch       */
#define CHEBPOL_CUDA_GET_GRID 106         /* grid size of the most recent launch          */
#define CHEBPOL_CUDA_GET_RANK 107         /* number of coordinates per point               */
#define CHEBPOL_CUDA_GET_MMA_MACS 108     /* multiply-adds per point the DMMA engine executes for this plan (tensor-core
                                             MACs incl. padding, plus the per-row fold and per-column close); 0 if the
                                             plan has no DMMA configuration.  The reference's Clenshaw count is
                                             S = sum_k prod_{j>=k} n_j; the engine needs fewer when it folds two or more
                                             dimensions into K, so an algorithmic roofline fraction can exceed 1 */

/* kernel selection.  AUTO picks the fastest kernel that supports the shape.
 * STRICT is the reference's operation order with no fused multiply-add:
 * bit-identical to the reference C (up to libm exp/sin for blends 1,2 and the
 * uniform map).  The others are IEEE FP64 kernels that reorder the sums
 * (basis-contraction instead of Clenshaw, FMA), within 1e-12*max|f|. */
#define CHEBPOL_CUDA_KERNEL_AUTO 0
#define CHEBPOL_CUDA_KERNEL_STRICT 1
#define CHEBPOL_CUDA_KERNEL_FAST 2
/* CHEBPOL_CUDA_OPT_VARIANT (tests, tuning, ncu evidence; 0 lets the library choose):
 *   Chebyshev plans   1            DFMA slab kernel, default layout
 *                     2..99999     DFMA slab kernel, layout ROWS*10000 + PT*100 + WARPS (n0 = 10, 12 only)
 *                     100000       DMMA contraction engine, default layout
 *                     100000+c     DMMA engine, layout c = MT*100 + WARPS
 *   multilinear plans 1 pair-gather kernel on the R layout; 2 cell-major kernel (2^D x the table);
 *                     3 quad kernel (row pairs interleaved, 2 x the table); 4 binned path (points
 *                     reordered by cell, sub-tables staged in shared memory, 1 x the table plus a
 *                     workspace proportional to the batch); 20..199 cell-major
 *                     kernel with WARPS*10 + STAGES; 300+c quad kernel with c CTAs per SM
 *   FH plans          100000+c as for Chebyshev
 * CHEBPOL_CUDA_GET_KERNEL_USED: 10 cheb_strict, 11 cheb_slab (DFMA), 12 cheb_mma (DMMA),
 *   20 mlip_strict, 21 mlip_pair, 22 mlip_cell, 23 mlip_quad, 24 mlip_bin, 30 fh_strict, 32 fh_mma, 40/41 polyh, 50/51 stalk,
 *   60 sl_strict, 61 sl_cell. */
CHEBPOL_CUDA_API int chebpol_cuda_plan_set(chebpol_cuda_plan *plan, int option, int64_t value);
CHEBPOL_CUDA_API int chebpol_cuda_plan_get(chebpol_cuda_plan *plan, int what, int64_t *value);

/* ---- fit side (the step before the evaluation path) ----------------------
 * chebcoef: values on the Chebyshev grid (R array `val`, dims[rank]) -> Chebyshev
 * coefficients, replacing R_chebcoef / chebcoef (src/chebpol.c:145-180, :8-141); dct != 0
 * returns the raw DCT-II as chebcoef(val, TRUE) does.  fhweights: Floater-Hormann weights of
 * one knot vector for degree d, replacing fhweights (src/chebpol.c:267-282; R_fhweights
 * :285-312 calls it once per dimension).  Host pointers in and out. */
CHEBPOL_CUDA_API int chebpol_cuda_chebcoef(const double *val, const int *dims, int rank, int dct,
                                           int device, double *coef);
CHEBPOL_CUDA_API int chebpol_cuda_fhweights(const double *grid, int nknots, int d, int device, double *w);

/* ---- polyharmonic splines on scattered knots ------------
 * Replaces R_evalpolyh / C_evalpolyh, src/chebpol.c:383-451, as called by the closure of
 * polyh.real (R/polyharmonic.R:129): f(x) = sum_i w_i phi(|x - k_i|) + v_0 + sum_j v_{j+1} x_j,
 * phi(r) = exp(k r^2) for k < 0, r^k for odd (int)k, log(r) r^k for even (int)k.
 * knots: rank x nknots R matrix (knot i at knots + i*rank); weights: nknots; lweights: rank+1.
 * norm_a/norm_b (both or neither, `rank` entries): the closure's normfun x <- a*(x - b)
 * (R/polyharmonic.R:77,82) applied inside the kernel instead of in R.
 * Kernel ids: 40 polyh_strict, 41 polyh_tile.  rank <= CHEBPOL_CUDA_MAXRANK. */
CHEBPOL_CUDA_API int chebpol_cuda_evalpolyh(const double *knots, int rank, int nknots, const double *weights,
                                            const double *lweights, double k, const double *x,
                                            int64_t npts, const double *norm_a, const double *norm_b,
                                            int ngpus, double *out);
CHEBPOL_CUDA_API int chebpol_cuda_plan_polyh(const double *knots, int rank, int nknots, const double *weights,
                                             const double *lweights, double k, const double *norm_a,
                                             const double *norm_b, int device, chebpol_cuda_plan **plan);

/* ---- stalker splines ------------------------------------
 * Replace R_evalstalk / R_evalhyp, src/stalker.c:765-832 (evalstalk :366-442, evalhyp :537-613),
 * as called by the closures of stalkerappx / hstalkerappx (R/stalker.R:80-85,62-67).  The tables
 * are what R_makestalk (b, c, r) / R_makehyp (a, b, c, d) return: a has prod(dims) entries, the
 * others are rank x prod(dims) R matrices (entry j of grid point i at [rank*i + j]); val is the
 * R array of values, grid[d] the dims[d] increasing knots of dimension d.  blend as for
 * chebpol_cuda_evalmlip (0 linear .. 5 mean; the closures default to 3, cubic).
 * Kernel ids: 50 stalk_strict, 51 stalk_cell (rank <= 6).
 * chebpol_cuda_plan_stalk: kind 0 stalker (a ignored, t1 = c, t2 = r), kind 1 hstalker (t1 = c, t2 = d). */
CHEBPOL_CUDA_API int chebpol_cuda_evalstalk(const double *val, const double *const *grid, const int *dims, int rank,
                                            const double *b, const double *c, const double *r, int blend,
                                            const double *x, int64_t npts, int ngpus, double *out);
CHEBPOL_CUDA_API int chebpol_cuda_evalhyp(const double *val, const double *const *grid, const int *dims, int rank,
                                          const double *a, const double *b, const double *c, const double *d,
                                          int blend, const double *x, int64_t npts, int ngpus, double *out);
CHEBPOL_CUDA_API int chebpol_cuda_plan_stalk(int kind, const double *val, const double *const *grid,
                                             const int *dims, int rank, const double *a, const double *b,
                                             const double *t1, const double *t2, int blend, int device,
                                             chebpol_cuda_plan **plan);

/* ---- the deprecated recursive stalker evaluator (ipol(method='oldstalker')) -----------------
 * Replaces R_evalstalker / evalstalker, src/stalker.c:617-678, 834-891, as called by the closure of
 * oldstalkerappx (R/stalker.R:35-46).  val: R array of values; grid[d]: dims[d] >= 2 increasing knots;
 * degree[rank]: the degree r of every dimension (0: hyperbolic, NaN: per-basis degree); det, pmin,
 * pplus: the per-dimension vectors of stalkercontext (R/stalker.R:2-10), dims[d] entries each;
 * uniform[d] != 0 when grid d is uniform.  Like a reference build without GSL, a NaN degree on a
 * non-uniform grid is refused (src/stalker.c:871-874).  Kernel id 52 (oldstalk). */
CHEBPOL_CUDA_API int chebpol_cuda_evalstalker(const double *val, const double *const *grid, const int *dims, int rank,
                                              const double *degree, const double *const *det,
                                              const double *const *pmin, const double *const *pplus,
                                              const int *uniform, int blend, const double *x, int64_t npts,
                                              int ngpus, double *out);
CHEBPOL_CUDA_API int chebpol_cuda_plan_oldstalk(const double *val, const double *const *grid, const int *dims,
                                                int rank, const double *degree, const double *const *det,
                                                const double *const *pmin, const double *const *pplus,
                                                const int *uniform, int blend, int device,
                                                chebpol_cuda_plan **plan);

/* ---- simplex-linear interpolation on a Delaunay triangulation -------------
 * Replaces R_evalsl / findsimplex (src/chebpol.c:738-828), as called by the closure of
 * slappx.real (R/simplexlinear.R:34-39), and the fit step R_analyzesimplex (:830-875).
 * knots: dim x nknots R matrix (knot i at knots + i*dim); dtri: (dim+1) x numsimplex matrix of
 * 1-based knot indices (t(geometry::delaunayn(...))); lumats (numsimplex x (dim+1)^2), pivots
 * (numsimplex x (dim+1)) and bbox (2*dim x numsimplex) are the three elements of the list
 * R_analyzesimplex returns; val: nknots values.  The result for a point is that of the FIRST
 * simplex (index order) whose bounding box holds it and whose barycentric coordinates are all
 * >= -1e-10, blended (0 linear .. 5 mean; the closure defaults to 3, cubic); outside the hull
 * NA_real_, or with epol != 0 the barycentric extrapolation from the first simplex holding the
 * nearest knot.  Both kernels (60 sl_strict: the reference's linear scan; 61 sl_cell: candidates
 * from a uniform cell grid, dim <= 6) issue the reference's floating-point operations in the
 * reference's order and are bit-identical to it.
 * chebpol_cuda_analyzesimplex: LAPACK dgetrf semantics (first largest pivot, reciprocal-scaled
 * multipliers), one GPU thread per simplex. */
CHEBPOL_CUDA_API int chebpol_cuda_evalsl(const double *x, int64_t npts, const double *knots, int dim, int nknots,
                                         const int *dtri, int numsimplex, const double *lumats, const int *pivots,
                                         const double *bbox, const double *val, int epol, int blend, int ngpus,
                                         double *out);
CHEBPOL_CUDA_API int chebpol_cuda_plan_sl(const double *knots, int dim, int nknots, const int *dtri, int numsimplex,
                                          const double *lumats, const int *pivots, const double *bbox,
                                          const double *val, int device, chebpol_cuda_plan **plan);
CHEBPOL_CUDA_API int chebpol_cuda_analyzesimplex(const int *dtri, int numsimplex, const double *knots, int dim,
                                                 int nknots, int device, double *lumats, int *pivots, double *bbox);

/* Host-only (no GPU needed) view of the candidate lists sl_cell scans, for verification: cell
 * c_d = clamp(floor((x_d - lo[d]) * inv[d]), 0, cells_used-1), cell id = sum_d c_d cells_used^d,
 * candidates items[start[id] .. start[id+1]) in ascending simplex index. */
CHEBPOL_CUDA_API int chebpol_cuda_debug_sl_lists(const double *knots, int dim, int nknots, const int *dtri,
                                                 int numsimplex, const double *bbox, int cells_per_dim,
                                                 int *cells_used, double *lo, double *inv, int *start,
                                                 int64_t start_cap, int *items, int64_t items_cap, int64_t *nitems);

/* ---- evaluation on a tensor grid (evalongridV of an interpolant, R/tools.R:132-137) ----
 * The reference expands the grid (expand.grid) and evaluates point by point; here the
 * structure is used: one basis matrix per dimension and one mode product per dimension.
 * pts[d] holds gdims[d] coordinates of dimension d; out is the R array of dims gdims
 * (first dimension fastest, the order of expand.grid). */
CHEBPOL_CUDA_API int chebpol_cuda_evalgrid_cheb(const double *coef, const int *dims, int rank,
                                                const double *const *pts, const int *gdims,
                                                const double *mid, const double *ispan, const int *ugm_n,
                                                int device, double *out);
CHEBPOL_CUDA_API int chebpol_cuda_evalgrid_mlip(const double *const *grid, const int *dims, int rank,
                                                const double *values, const double *const *pts,
                                                const int *gdims, int blend, int device, double *out);
CHEBPOL_CUDA_API int chebpol_cuda_evalgrid_fh(const double *vals, const double *const *grid,
                                              const double *const *weights, const int *dims, int rank,
                                              const double *const *pts, const int *gdims, int device,
                                              double *out);

/* ---- measurement support ------------------------------------------------ */
/* Counter-based synthetic points: fills d_x (DEVICE pointer, npts*rank doubles,
 * point-major) with global points first_point .. first_point+npts-1 of the batch `seed`, every
 * coordinate uniform in [lo, hi) and a pure function of (seed, point index, dimension), so any
 * split of a batch over GPUs or onto the CPU regenerates identical values.  Asynchronous on
 * `stream`.  chebpol_b200/synth.py is the bit-identical numpy twin. */
CHEBPOL_CUDA_API int chebpol_cuda_synth_points(uint64_t seed, int64_t first_point, int64_t npts, int rank,
                                               double lo, double hi, double *d_x, int device, void *stream);

/* ---- diagnostics -------------------------------------------------------- */
/* The FP64 roofline denominator (MEASURED_PEAKS.json has none): the better of a
 * register-resident DFMA chain and a DMMA-only kernel on every SM of `device`.
 * Returns TFLOP/s (2 flop/FMA). */
CHEBPOL_CUDA_API int chebpol_cuda_fp64_peak(int device, int repeats, double *tflops);
/* The stateless entry points keep interpolants resident between calls (per device up to 8
 * entries and CHEBPOL_CUDA_CACHE_MAX_GB, default 64, of device memory; least recently used out
 * first), so the closure that re-passes its tensor on every call (src/chebpol.c:340-381) does not
 * re-upload it, and a multilinear plan keeps its expanded cell table.  The key is the CONTENT:
 * kind, device, dims, byte count and a 128-bit fingerprint of every defining array (tensor,
 * knots, weights, pre-map).  A hit is verified: entries up to 32 MB retain a host copy and the
 * bytes are compared, larger ones rely on the fingerprint (collision probability ~2^-128 per
 * pair).  Clearing is never required for correctness; CHEBPOL_CUDA_PLAN_CACHE=0 disables the
 * cache.  Entries another thread is evaluating are destroyed when that evaluation returns. */
CHEBPOL_CUDA_API void chebpol_cuda_cache_clear(void);
/* cache facts (any pointer may be NULL): live entries, verified hits, misses, fingerprint matches
 * whose bytes differed, device bytes held */
CHEBPOL_CUDA_API int chebpol_cuda_debug_cache_stats(int64_t *entries, int64_t *hits, int64_t *misses,
                                                    int64_t *rejected, int64_t *device_bytes);
/* test hook: on != 0 makes every fingerprint equal, so that only the byte comparison separates
 * interpolants of the same shape */
CHEBPOL_CUDA_API void chebpol_cuda_debug_cache_weak_hash(int on);
/* host only: the 128-bit fingerprint of one array, computed by `threads` host threads (the value
 * does not depend on the thread count) */
CHEBPOL_CUDA_API int chebpol_cuda_debug_fingerprint(const void *data, int64_t nbytes, int threads, uint64_t *out2);
/* host only: would a call defined by array b be served by the cache entry made from array a?
 * retain = 0 models an entry above the 32 MB retention limit (fingerprint only) */
CHEBPOL_CUDA_API int chebpol_cuda_debug_cache_would_hit(const void *a, int64_t na, const void *b, int64_t nb,
                                                        int retain, int *hit);
/* host only (no device needed): the geometry the DMMA contraction engine chooses for a Chebyshev (kind 0) or
 * Floater-Hormann (kind 1) tensor and `npt` points per pass (a multiple of 8, e.g. 128).  out[16]:
 * 0 engine takes the shape (0/1), 1 folded dims, 2 K, 3 k-chunks KC, 4 row stride rs (doubles), 5 rows,
 * 6 rows per slab, 7 slabs, 8 ring stages, 9 stage bytes, 10 tensor resident in shared memory (0/1),
 * 11 rows stored pair-interleaved (0/1), 12 dynamic shared memory of the ring + basis table (bytes),
 * 13 level-A size, 14 level-A size padded, 15 r' columns.  Used by the layout tests. */
CHEBPOL_CUDA_API int chebpol_cuda_debug_mma_geometry(int kind, const int *dims, int rank, int npt, int64_t *out16);
/* host only: position inside a stored row of coefficient k of a tensor with `kc` k-chunks */
CHEBPOL_CUDA_API int chebpol_cuda_debug_mma_row_pos(int kc, int k);
/* total kernels launched by this library in this process */
CHEBPOL_CUDA_API int64_t chebpol_cuda_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* CHEBPOL_CUDA_H */
