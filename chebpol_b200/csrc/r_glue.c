/* r_glue.c -- the .Call boundary of chebpol rebound to libchebpol_cuda.
 *
 * Drop-in replacements for the three hot-path entry points of the reference,
 * with the SAME argument lists, the same validation and error messages, and the
 * same return value (a fresh REALSXP of length P, no dim attribute):
 *
 *   R_evalcheb_cuda(coef, inx, Rthreads, spare)            <- R_evalcheb, src/chebpol.c:340-381
 *   R_evalmlip_cuda(sgrid, values, x, Rthreads, Sblend)    <- R_evalmlip, src/chebpol.c:670-707
 *   R_FH_cuda(inx, vals, grid, Sweights, Rthreads, spare)  <- R_FH,       src/chebpol.c:220-265
 *
 * `spare` (unused in the reference, :347 and :227) carries the per-call pre-map of the
 * closure so that it runs inside the kernel instead of in R:
 *   NULL                          no pre-map (chebappx without intervals)
 *   list(mid=, ispan=)            imap, R/chebyshev.R:223-227
 *   list(ugm=dims[, mid=, ispan=]) uniform-grid map, R/uniform.R:6,58-77
 *   list(splines=lapply(gridmaps, function(gm) environment(gm)$z))
 *                                 the 'general' grid maps, R/chebyshev.R:349-354: the x, y, b, c, d
 *                                 tables each stats::splinefun(method='hyman') closure holds
 * `Rthreads` keeps its slot; a value <= 0 means "all visible GPUs", otherwise it is the
 * number of GPUs to spread the point batch over (CHEBPOL_CUDA_GPUS overrides).
 *
 * Errors: the C ABI never longjmps.  The glue validates on R's main thread, calls the
 * library, and only then -- with nothing of its own left to free -- turns a non-zero
 * return code into error(), exactly where the reference raises its errors.
 *
 * This file needs R's headers (R.h, Rdefines.h, R_ext/Rdynload.h).  R is not installed
 * in the build image; it is compile-checked against the declarations in oracle/shim
 * (`make -C chebpol_b200/csrc rglue-check`) and is NOT part of libchebpol_cuda.so.
 * INTEGRATION.md shows the registration-table and closure patches that go with it.
 */
#include <R.h>
#include <Rdefines.h>
#include <R_ext/Rdynload.h>
#include <stdlib.h>
#include "../../include/chebpol_cuda.h"

#define UNUSED(x) (void)(x)

static int gpus_from_threads(SEXP Rthreads)
{
    const char *env = getenv("CHEBPOL_CUDA_GPUS");
    if (env) return atoi(env);
    (void)INTEGER(AS_INTEGER(Rthreads))[0]; /* same coercion (and same error) as the reference */
    return 0;                               /* all visible GPUs */
}

static SEXP list_elt(SEXP list, const char *name)
{
    if (isNull(list)) return R_NilValue;
    SEXP names = getAttrib(list, R_NamesSymbol);
    for (int i = 0; i < LENGTH(list); i++)
        if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(list, i);
    return R_NilValue;
}

static void raise(int rc)
{
    if (rc != CHEBPOL_CUDA_OK)
        error("chebpol_cuda: %s (%s)", chebpol_cuda_strerror(rc), chebpol_cuda_last_error());
}

SEXP R_evalcheb_cuda(SEXP coef, SEXP inx, SEXP Rthreads, SEXP spare)
{
    int *dims;
    int siz = 1;
    double *cf = REAL(coef);
    SEXP dim;
    int rank;
    /* checks and messages of src/chebpol.c:349-365 */
    dim = getAttrib(coef, R_DimSymbol);
    dims = INTEGER(dim);
    rank = LENGTH(dim);
    if (rank <= 0) error("rank must be positive");
    if (isMatrix(inx)) {
        if (rank != nrows(inx))
            error("coeffcient rank(%d) must match number of rows(%d)", LENGTH(dim), nrows(inx));
    } else {
        if (rank != LENGTH(inx))
            error("coefficient rank(%d) does not match argument length(%d)", LENGTH(dim), LENGTH(inx));
    }
    for (int i = 0; i < rank; i++) siz *= dims[i];
    if (LENGTH(coef) != siz)
        error("coefficient length(%d) does not match data length(%d)", LENGTH(coef), siz);

    const double *mid = NULL, *ispan = NULL;
    const int *ugm = NULL;
    SEXP smid = list_elt(spare, "mid"), sispan = list_elt(spare, "ispan"), sugm = list_elt(spare, "ugm");
    if (!isNull(smid) && !isNull(sispan)) {
        if (LENGTH(smid) != rank || LENGTH(sispan) != rank) error("mid/ispan must have one entry per dimension");
        mid = REAL(smid);
        ispan = REAL(sispan);
    }
    if (!isNull(sugm)) {
        if (LENGTH(sugm) != rank) error("ugm must have one entry per dimension");
        ugm = INTEGER(sugm);
    }
    const int ngpus = gpus_from_threads(Rthreads);
    const R_xlen_t numvec = isMatrix(inx) ? ncols(inx) : 1;
    SEXP ssp = list_elt(spare, "splines");
    if (!isNull(ssp)) {
        /* chebappxg.real: ch(gridmap(x), threads) with the grid maps evaluated in the kernel */
        static const char *nm[5] = {"x", "y", "b", "c", "d"};
        const double *tab[5][CHEBPOL_CUDA_MAXRANK];
        int sn[CHEBPOL_CUDA_MAXRANK];
        if (rank > CHEBPOL_CUDA_MAXRANK) error("rank must be positive");
        if (LENGTH(ssp) != rank) error("splines must have one entry per dimension");
        for (int k = 0; k < rank; k++) {
            SEXP z = VECTOR_ELT(ssp, k);
            sn[k] = LENGTH(list_elt(z, "x"));
            for (int j = 0; j < 5; j++) {
                SEXP v = list_elt(z, nm[j]);
                if (isNull(v) || !IS_NUMERIC(v) || LENGTH(v) != sn[k]) error("spline table %s of dimension %d is malformed", nm[j], k + 1);
                tab[j][k] = REAL(v);
            }
        }
        SEXP res = PROTECT(NEW_NUMERIC(numvec));
        const int rcg = chebpol_cuda_evalchebg(cf, dims, rank, REAL(inx), (int64_t)numvec, sn, tab[0], tab[1], tab[2],
                                               tab[3], tab[4], ngpus, REAL(res));
        UNPROTECT(1);
        raise(rcg);
        return res;
    }
    SEXP resvec = PROTECT(NEW_NUMERIC(numvec));
    const int rc = chebpol_cuda_evalcheb(cf, dims, rank, REAL(inx), (int64_t)numvec, mid, ispan, ugm, ngpus,
                                         REAL(resvec));
    UNPROTECT(1);
    raise(rc);
    return resvec;
}

SEXP R_evalmlip_cuda(SEXP sgrid, SEXP values, SEXP x, SEXP Rthreads, SEXP Sblend)
{
    const int rank = LENGTH(sgrid);
    int gridsize = 1;
    int blend = 0;
    /* checks and messages of src/chebpol.c:677-692 */
    if (!IS_NUMERIC(values)) error("values must be numeric");
    if (!IS_NUMERIC(x)) error("argument x must be numeric");
    if (isMatrix(x) ? (nrows(x) != rank) : (LENGTH(x) != rank))
        error("grid has dimension %d, you supplied a length %d vector", rank, isMatrix(x) ? nrows(x) : LENGTH(x));
    if (rank <= 0 || rank > CHEBPOL_CUDA_MAXRANK) error("rank must be positive");
    int dims[CHEBPOL_CUDA_MAXRANK];
    const double *grid[CHEBPOL_CUDA_MAXRANK];
    for (int i = 0; i < rank; i++) {
        dims[i] = LENGTH(VECTOR_ELT(sgrid, i));
        grid[i] = REAL(VECTOR_ELT(sgrid, i));
        gridsize *= dims[i];
    }
    if (LENGTH(values) != gridsize) error("grid has size %d, you supplied %d values", gridsize, LENGTH(values));
    if (!isNull(Sblend)) blend = INTEGER(AS_INTEGER(Sblend))[0];
    const int ngpus = gpus_from_threads(Rthreads);
    const R_xlen_t numvec = isMatrix(x) ? ncols(x) : 1;
    SEXP resvec = PROTECT(NEW_NUMERIC(numvec));
    const int rc = chebpol_cuda_evalmlip(grid, dims, rank, REAL(values), REAL(x), (int64_t)numvec, blend, ngpus,
                                         REAL(resvec));
    UNPROTECT(1);
    raise(rc);
    return resvec;
}

SEXP R_FH_cuda(SEXP inx, SEXP vals, SEXP grid, SEXP Sweights, SEXP Rthreads, SEXP spare)
{
    int *dims;
    int siz = 1;
    SEXP dim;
    int rank;
    UNUSED(spare);
    /* checks and messages of src/chebpol.c:229-248 */
    dim = getAttrib(vals, R_DimSymbol);
    dims = INTEGER(dim);
    rank = LENGTH(dim);
    if (rank <= 0) error("rank must be positive");
    if (isMatrix(inx)) {
        if (rank != nrows(inx))
            error("coeffcient rank(%d) must match number of rows(%d)", LENGTH(dim), nrows(inx));
    } else {
        if (rank != LENGTH(inx))
            error("coefficient rank(%d) does not match argument length(%d)", LENGTH(dim), LENGTH(inx));
    }
    for (int i = 0; i < rank; i++) siz *= dims[i];
    if (LENGTH(vals) != siz) error("coefficient length(%d) does not match data length(%d)", LENGTH(vals), siz);
    if (LENGTH(grid) != rank) error("There must be one value for each knot");
    if (rank > CHEBPOL_CUDA_MAXRANK) error("rank must be positive");
    const double *knots[CHEBPOL_CUDA_MAXRANK], *weights[CHEBPOL_CUDA_MAXRANK];
    for (int i = 0; i < rank; i++) {
        knots[i] = REAL(VECTOR_ELT(grid, i));
        weights[i] = REAL(VECTOR_ELT(Sweights, i));
    }
    const int ngpus = gpus_from_threads(Rthreads);
    const R_xlen_t numvec = isMatrix(inx) ? ncols(inx) : 1;
    SEXP resvec = PROTECT(NEW_NUMERIC(numvec));
    const int rc = chebpol_cuda_fh(REAL(vals), knots, weights, dims, rank, REAL(inx), (int64_t)numvec, ngpus,
                                   REAL(resvec));
    UNPROTECT(1);
    raise(rc);
    return resvec;
}

/* Polyharmonic splines: same arguments and return as R_evalpolyh (src/chebpol.c:416-451).
 * `spare` (unused in the reference, :424) may carry the closure's normalisation
 * list(a=, b=) so that normfun(x) = a*(x-b) (R/polyharmonic.R:77,82,129) runs inside the kernel. */
SEXP R_evalpolyh_cuda(SEXP inx, SEXP Sknots, SEXP weights, SEXP lweights, SEXP Sk, SEXP Sthreads, SEXP spare)
{
    const double k = REAL(AS_NUMERIC(Sk))[0];
    const int rank = nrows(Sknots);
    R_xlen_t numvec;
    /* checks and messages of src/chebpol.c:425-441 */
    if (LENGTH(lweights) != rank + 1)
        error("linear weights (%d) should be one longer than rank(%d)", LENGTH(lweights), rank);
    if (LENGTH(weights) != ncols(Sknots))
        error("number of weights (%d) should equal number of knots (%d)", LENGTH(weights), ncols(Sknots));
    if (isMatrix(inx)) {
        if (nrows(inx) != nrows(Sknots))
            error("Dimension of input (%d) must match dimension of interpolant (%d)", nrows(inx), nrows(Sknots));
        numvec = ncols(inx);
    } else {
        if (length(inx) != nrows(Sknots))
            error("Dimension of input (%d) must match dimension of interpolant (%d)", length(inx), nrows(Sknots));
        numvec = 1;
    }
    const double *na = NULL, *nb = NULL;
    SEXP sa = list_elt(spare, "a"), sb = list_elt(spare, "b");
    if (!isNull(sa) && !isNull(sb)) {
        if (LENGTH(sa) != rank || LENGTH(sb) != rank) error("normalisation must have one entry per dimension");
        na = REAL(sa);
        nb = REAL(sb);
    }
    const int ngpus = gpus_from_threads(Sthreads);
    SEXP res = PROTECT(NEW_NUMERIC(numvec));
    const int rc = chebpol_cuda_evalpolyh(REAL(Sknots), rank, ncols(Sknots), REAL(weights), REAL(lweights), k,
                                          REAL(inx), (int64_t)numvec, na, nb, ngpus, REAL(res));
    UNPROTECT(1);
    raise(rc);
    return res;
}

/* Stalker splines: same arguments and return as R_evalstalk / R_evalhyp (src/stalker.c:765-832);
 * `stalker` is the list R_makestalk (b, c, r, val, grid) / R_makehyp (a, b, c, d, val, grid) returned. */
static SEXP stalk_common(int kind, SEXP Sx, SEXP stalker, SEXP Sblend, SEXP Sthreads)
{
    const int off = kind ? 1 : 0;
    SEXP Sgrid = VECTOR_ELT(stalker, 4 + off);
    const int rank = LENGTH(Sgrid);
    if (rank <= 0 || rank > CHEBPOL_CUDA_MAXRANK) error("rank must be positive");
    if (isMatrix(Sx) ? (nrows(Sx) != rank) : (LENGTH(Sx) != rank))
        error("Rank of input(%d) does not match rank of spline", isMatrix(Sx) ? nrows(Sx) : LENGTH(Sx));
    int dims[CHEBPOL_CUDA_MAXRANK];
    const double *grid[CHEBPOL_CUDA_MAXRANK];
    for (int i = 0; i < rank; i++) {
        dims[i] = LENGTH(VECTOR_ELT(Sgrid, i));
        grid[i] = REAL(VECTOR_ELT(Sgrid, i));
    }
    const int blend = INTEGER(Sblend)[0];
    const int ngpus = gpus_from_threads(Sthreads);
    const R_xlen_t N = isMatrix(Sx) ? ncols(Sx) : 1;
    SEXP ret = PROTECT(NEW_NUMERIC(N));
    int rc;
    if (kind == 0)
        rc = chebpol_cuda_evalstalk(REAL(VECTOR_ELT(stalker, 3)), grid, dims, rank, REAL(VECTOR_ELT(stalker, 0)),
                                    REAL(VECTOR_ELT(stalker, 1)), REAL(VECTOR_ELT(stalker, 2)), blend, REAL(Sx),
                                    (int64_t)N, ngpus, REAL(ret));
    else
        rc = chebpol_cuda_evalhyp(REAL(VECTOR_ELT(stalker, 4)), grid, dims, rank, REAL(VECTOR_ELT(stalker, 0)),
                                  REAL(VECTOR_ELT(stalker, 1)), REAL(VECTOR_ELT(stalker, 2)),
                                  REAL(VECTOR_ELT(stalker, 3)), blend, REAL(Sx), (int64_t)N, ngpus, REAL(ret));
    UNPROTECT(1);
    raise(rc);
    return ret;
}

SEXP R_evalstalk_cuda(SEXP Sx, SEXP stalker, SEXP Sblend, SEXP Sthreads)
{
    return stalk_common(0, Sx, stalker, Sblend, Sthreads);
}

SEXP R_evalhyp_cuda(SEXP Sx, SEXP stalker, SEXP Sblend, SEXP Sthreads)
{
    return stalk_common(1, Sx, stalker, Sblend, Sthreads);
}

/* fit side: same arguments and return as R_chebcoef (src/chebpol.c:145-180) */
SEXP R_chebcoef_cuda(SEXP x, SEXP sdct, SEXP Sthreads)
{
    SEXP dim;
    int rank, sdims, *dims, siz = 1;
    UNUSED(Sthreads);
    if (!isLogical(sdct) || LENGTH(sdct) < 1) error("dct must be a logical");
    const int dct = LOGICAL(sdct)[0];
    dim = getAttrib(x, R_DimSymbol);
    if (isNull(dim)) {
        sdims = LENGTH(x);
        dims = &sdims;
        rank = 1;
    } else {
        rank = LENGTH(dim);
        if (rank <= 0) error("Rank must be positive");
        dims = INTEGER(dim);
    }
    for (int i = 0; i < rank; i++) siz *= dims[i];
    if (siz == 0) error("array size must be positive");
    SEXP resvec = PROTECT(NEW_NUMERIC(siz));
    const int rc = chebpol_cuda_chebcoef(REAL(x), dims, rank, dct, 0, REAL(resvec));
    if (rc == CHEBPOL_CUDA_OK) {
        setAttrib(resvec, R_DimSymbol, dim);
        setAttrib(resvec, R_DimNamesSymbol, getAttrib(x, R_DimNamesSymbol));
    }
    UNPROTECT(1);
    raise(rc);
    return resvec;
}

/* same arguments and return as R_fhweights (src/chebpol.c:285-312) */
SEXP R_fhweights_cuda(SEXP Sgrid, SEXP Sd, SEXP Sthreads)
{
    UNUSED(Sthreads);
    if (LENGTH(Sgrid) != LENGTH(Sd))
        error("Length of grid (%d) should equal length of d (%d)", LENGTH(Sgrid), LENGTH(Sd));
    const int rank = LENGTH(Sgrid);
    const int *dd = INTEGER(AS_INTEGER(Sd));
    SEXP ret = PROTECT(NEW_LIST(rank));
    int rc = CHEBPOL_CUDA_OK;
    for (int r = 0; r < rank && rc == CHEBPOL_CUDA_OK; r++) {
        const int n = LENGTH(VECTOR_ELT(Sgrid, r));
        SET_VECTOR_ELT(ret, r, NEW_NUMERIC(n));
        rc = chebpol_cuda_fhweights(REAL(VECTOR_ELT(Sgrid, r)), n, dd[r], 0, REAL(VECTOR_ELT(ret, r)));
    }
    UNPROTECT(1);
    raise(rc);
    return ret;
}

/* same arguments and return as R_evalstalker (src/stalker.c:834-891): stalker is the list
 * c(list(val=, grid=), stalkercontext(...), uniform=) built by oldstalkerappx (R/stalker.R:29) */
SEXP R_evalstalker_cuda(SEXP Sx, SEXP stalker, SEXP Sdegree, SEXP Sblend, SEXP Sthreads)
{
    SEXP Sgrid = VECTOR_ELT(stalker, 1);
    const int nrank = LENGTH(Sgrid);
    const int K = isMatrix(Sx) ? nrows(Sx) : LENGTH(Sx);
    if (K != nrank) error("Rank of input(%d) does not match rank of spline", K, nrank);
    if (nrank <= 0 || nrank > CHEBPOL_CUDA_MAXRANK) error("rank must be positive");
    int dims[CHEBPOL_CUDA_MAXRANK], len = 1;
    const double *grid[CHEBPOL_CUDA_MAXRANK], *det[CHEBPOL_CUDA_MAXRANK], *pmin[CHEBPOL_CUDA_MAXRANK],
        *pplus[CHEBPOL_CUDA_MAXRANK];
    for (int r = 0; r < nrank; r++) {
        dims[r] = LENGTH(VECTOR_ELT(Sgrid, r));
        len *= dims[r];
        grid[r] = REAL(VECTOR_ELT(Sgrid, r));
        det[r] = REAL(VECTOR_ELT(VECTOR_ELT(stalker, 2), r));
        pmin[r] = REAL(VECTOR_ELT(VECTOR_ELT(stalker, 3), r));
        pplus[r] = REAL(VECTOR_ELT(VECTOR_ELT(stalker, 4), r));
    }
    if (len != LENGTH(VECTOR_ELT(stalker, 0)))
        error("number of values (%d) does not match grid size (%d)\n", LENGTH(VECTOR_ELT(stalker, 0)), len);
    if (LENGTH(Sdegree) != nrank) error("degree length must match dimension (%d)", nrank);
    const double *degree = REAL(AS_NUMERIC(Sdegree));
    const int *uniform = LOGICAL(VECTOR_ELT(stalker, 5));
    const int blend = INTEGER(AS_INTEGER(Sblend))[0];
    const int ngpus = gpus_from_threads(Sthreads);
    const R_xlen_t N = isMatrix(Sx) ? ncols(Sx) : 1;
    SEXP ret = PROTECT(NEW_NUMERIC(N));
    const int rc = chebpol_cuda_evalstalker(REAL(VECTOR_ELT(stalker, 0)), grid, dims, nrank, degree, det, pmin, pplus,
                                            uniform, blend, REAL(Sx), (int64_t)N, ngpus, REAL(ret));
    UNPROTECT(1);
    raise(rc);
    return ret;
}

/* same arguments and return as R_evalsl (src/chebpol.c:803-828), called by the closure of
 * slappx.real (R/simplexlinear.R:38) */
SEXP R_evalsl_cuda(SEXP Sx, SEXP Sknots, SEXP Sdtri, SEXP adata, SEXP Sval, SEXP extrapolate, SEXP Sthreads,
                   SEXP Sblend)
{
    const int dim = nrows(Sknots);
    const int nknots = ncols(Sknots);
    const int numsimplex = ncols(Sdtri);
    if (nrows(Sdtri) != dim + 1) error("triangulation has %d rows, knots have dimension %d", nrows(Sdtri), dim);
    if (isMatrix(Sx) ? (nrows(Sx) != dim) : (LENGTH(Sx) != dim)) error("argument x must have %d rows", dim);
    if (LENGTH(Sval) != nknots) error("number of values (%d) should equal number of knots (%d)", LENGTH(Sval), nknots);
    const int epol = LOGICAL(AS_LOGICAL(extrapolate))[0];
    int blend = 0;
    if (!isNull(Sblend)) blend = INTEGER(AS_INTEGER(Sblend))[0];
    const int ngpus = gpus_from_threads(Sthreads);
    const R_xlen_t numvec = isMatrix(Sx) ? ncols(Sx) : 1;
    SEXP ret = PROTECT(NEW_NUMERIC(numvec));
    const int rc = chebpol_cuda_evalsl(REAL(Sx), (int64_t)numvec, REAL(Sknots), dim, nknots, INTEGER(Sdtri), numsimplex,
                                       REAL(VECTOR_ELT(adata, 0)), INTEGER(VECTOR_ELT(adata, 1)),
                                       REAL(VECTOR_ELT(adata, 2)), REAL(Sval), epol, blend, ngpus, REAL(ret));
    UNPROTECT(1);
    raise(rc);
    return ret;
}

/* fit side: same arguments and return as R_analyzesimplex (src/chebpol.c:830-875) */
SEXP R_analyzesimplex_cuda(SEXP Sdtri, SEXP Sknots, SEXP Sthreads)
{
    UNUSED(Sthreads);
    const int dim = nrows(Sknots);
    const int numsimplex = ncols(Sdtri);
    const int N = dim + 1;
    SEXP retlist = PROTECT(NEW_LIST(3));
    SET_VECTOR_ELT(retlist, 0, NEW_NUMERIC(numsimplex * N * N));
    SET_VECTOR_ELT(retlist, 1, NEW_INTEGER(numsimplex * N));
    SET_VECTOR_ELT(retlist, 2, allocMatrix(REALSXP, 2 * dim, numsimplex));
    const int rc = chebpol_cuda_analyzesimplex(INTEGER(Sdtri), numsimplex, REAL(Sknots), dim, ncols(Sknots), 0,
                                               REAL(VECTOR_ELT(retlist, 0)), INTEGER(VECTOR_ELT(retlist, 1)),
                                               REAL(VECTOR_ELT(retlist, 2)));
    UNPROTECT(1);
    raise(rc);
    return retlist;
}

/* ---- resident interpolants ------------------------------------------------------------------
 * The stateless entries above re-receive the tensor on every call, as the reference does, and
 * recognise it by content (one pass over its bytes per call: 3 ms for the 134 MB 64^4 table).
 * A closure that is called many times on few points can instead keep the device copy alive:
 *     plan <- .Call(C_plan_mlip_cuda, grid, values, blend, 0L)       # once, when the closure is made
 *     function(x, threads, blend) .Call(C_plan_eval_cuda, plan, x, blendcode)
 * The plan is an R external pointer; R's garbage collector destroys the device copy with the
 * closure (finalizer), so no C-side state outlives the R object that owns it. */
static void plan_finalizer(SEXP ptr)
{
    chebpol_cuda_plan *p = (chebpol_cuda_plan *)R_ExternalPtrAddr(ptr);
    if (p) {
        chebpol_cuda_plan_destroy(p);
        R_ClearExternalPtr(ptr);
    }
}

static SEXP wrap_plan(chebpol_cuda_plan *p, int rank)
{
    SEXP tag = PROTECT(NEW_INTEGER(1));
    INTEGER(tag)[0] = rank;
    SEXP ptr = PROTECT(R_MakeExternalPtr(p, tag, R_NilValue));
    R_RegisterCFinalizerEx(ptr, plan_finalizer, TRUE);
    UNPROTECT(2);
    return ptr;
}

SEXP R_plan_cheb_cuda(SEXP coef, SEXP spare, SEXP Sdevice)
{
    SEXP dim = getAttrib(coef, R_DimSymbol);
    const int rank = LENGTH(dim);
    int siz = 1;
    if (rank <= 0) error("rank must be positive");
    for (int i = 0; i < rank; i++) siz *= INTEGER(dim)[i];
    if (LENGTH(coef) != siz) error("coefficient length(%d) does not match data length(%d)", LENGTH(coef), siz);
    SEXP smid = list_elt(spare, "mid"), sispan = list_elt(spare, "ispan"), sugm = list_elt(spare, "ugm");
    const int have_map = !isNull(smid) && !isNull(sispan);
    if (have_map && (LENGTH(smid) != rank || LENGTH(sispan) != rank)) error("mid/ispan must have one entry per dimension");
    if (!isNull(sugm) && LENGTH(sugm) != rank) error("ugm must have one entry per dimension");
    chebpol_cuda_plan *p = NULL;
    raise(chebpol_cuda_plan_cheb(REAL(coef), INTEGER(dim), rank, have_map ? REAL(smid) : NULL, have_map ? REAL(sispan) : NULL,
                                 isNull(sugm) ? NULL : INTEGER(sugm), INTEGER(AS_INTEGER(Sdevice))[0], &p));
    return wrap_plan(p, rank);
}

SEXP R_plan_mlip_cuda(SEXP sgrid, SEXP values, SEXP Sblend, SEXP Sdevice)
{
    const int rank = LENGTH(sgrid);
    int gridsize = 1;
    if (!IS_NUMERIC(values)) error("values must be numeric");
    if (rank <= 0 || rank > CHEBPOL_CUDA_MAXRANK) error("rank must be positive");
    int dims[CHEBPOL_CUDA_MAXRANK];
    const double *grid[CHEBPOL_CUDA_MAXRANK];
    for (int i = 0; i < rank; i++) {
        dims[i] = LENGTH(VECTOR_ELT(sgrid, i));
        grid[i] = REAL(VECTOR_ELT(sgrid, i));
        gridsize *= dims[i];
    }
    if (LENGTH(values) != gridsize) error("grid has size %d, you supplied %d values", gridsize, LENGTH(values));
    chebpol_cuda_plan *p = NULL;
    raise(chebpol_cuda_plan_mlip(grid, dims, rank, REAL(values), isNull(Sblend) ? 0 : INTEGER(AS_INTEGER(Sblend))[0],
                                 INTEGER(AS_INTEGER(Sdevice))[0], &p));
    return wrap_plan(p, rank);
}

SEXP R_plan_fh_cuda(SEXP vals, SEXP grid, SEXP Sweights, SEXP Sdevice)
{
    SEXP dim = getAttrib(vals, R_DimSymbol);
    const int rank = LENGTH(dim);
    int siz = 1;
    if (rank <= 0 || rank > CHEBPOL_CUDA_MAXRANK) error("rank must be positive");
    for (int i = 0; i < rank; i++) siz *= INTEGER(dim)[i];
    if (LENGTH(vals) != siz) error("coefficient length(%d) does not match data length(%d)", LENGTH(vals), siz);
    if (LENGTH(grid) != rank || LENGTH(Sweights) != rank) error("There must be one value for each knot");
    const double *knots[CHEBPOL_CUDA_MAXRANK], *weights[CHEBPOL_CUDA_MAXRANK];
    for (int i = 0; i < rank; i++) {
        knots[i] = REAL(VECTOR_ELT(grid, i));
        weights[i] = REAL(VECTOR_ELT(Sweights, i));
    }
    chebpol_cuda_plan *p = NULL;
    raise(chebpol_cuda_plan_fh(REAL(vals), knots, weights, INTEGER(dim), rank, INTEGER(AS_INTEGER(Sdevice))[0], &p));
    return wrap_plan(p, rank);
}

/* evaluate a resident interpolant: x is the K x P matrix (or length-K vector) the closure received;
 * Sblend (multilinear plans; NULL keeps the plan's blend) is the per-call blend of R/multilinear.R:64-66 */
SEXP R_plan_eval_cuda(SEXP plan, SEXP x, SEXP Sblend)
{
    chebpol_cuda_plan *p = (chebpol_cuda_plan *)R_ExternalPtrAddr(plan);
    if (!p) error("the interpolant's device copy has been released (a closure restored by load() must be re-planned)");
    if (!IS_NUMERIC(x)) error("argument x must be numeric");
    int64_t rank = 0;
    raise(chebpol_cuda_plan_get(p, CHEBPOL_CUDA_GET_RANK, &rank));
    if (isMatrix(x) ? (nrows(x) != rank) : (LENGTH(x) != rank))
        error("coefficient rank(%d) does not match argument length(%d)", (int)rank, isMatrix(x) ? nrows(x) : LENGTH(x));
    if (!isNull(Sblend)) raise(chebpol_cuda_plan_set(p, CHEBPOL_CUDA_OPT_BLEND, INTEGER(AS_INTEGER(Sblend))[0]));
    const R_xlen_t numvec = isMatrix(x) ? ncols(x) : 1;
    SEXP resvec = PROTECT(NEW_NUMERIC(numvec));
    const int rc = chebpol_cuda_plan_eval_host(p, REAL(x), (int64_t)numvec, REAL(resvec));
    UNPROTECT(1);
    raise(rc);
    return resvec;
}

SEXP R_plan_free_cuda(SEXP plan)
{
    plan_finalizer(plan);
    return R_NilValue;
}

/* rows to append to R_CallMethodDef callMethods[] (src/chebpol.c:943-967), same arities
 * as "evalcheb", "evalmlip", "FH" */
const R_CallMethodDef chebpol_cuda_callMethods[] = {
    {"evalcheb_cuda", (DL_FUNC)&R_evalcheb_cuda, 4},
    {"evalmlip_cuda", (DL_FUNC)&R_evalmlip_cuda, 5},
    {"FH_cuda", (DL_FUNC)&R_FH_cuda, 6},
    {"chebcoef_cuda", (DL_FUNC)&R_chebcoef_cuda, 3},
    {"FHweights_cuda", (DL_FUNC)&R_fhweights_cuda, 3},
    {"evalpolyh_cuda", (DL_FUNC)&R_evalpolyh_cuda, 7},
    {"evalstalk_cuda", (DL_FUNC)&R_evalstalk_cuda, 4},
    {"evalhyp_cuda", (DL_FUNC)&R_evalhyp_cuda, 4},
    {"evalstalker_cuda", (DL_FUNC)&R_evalstalker_cuda, 5},
    {"evalsl_cuda", (DL_FUNC)&R_evalsl_cuda, 8},
    {"analyzesimplex_cuda", (DL_FUNC)&R_analyzesimplex_cuda, 3},
    {"plan_cheb_cuda", (DL_FUNC)&R_plan_cheb_cuda, 3},
    {"plan_mlip_cuda", (DL_FUNC)&R_plan_mlip_cuda, 4},
    {"plan_fh_cuda", (DL_FUNC)&R_plan_fh_cuda, 4},
    {"plan_eval_cuda", (DL_FUNC)&R_plan_eval_cuda, 3},
    {"plan_free_cuda", (DL_FUNC)&R_plan_free_cuda, 1},
    {NULL, NULL, 0}};
