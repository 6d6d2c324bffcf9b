/* oracle/chebpol_oracle.c -- TEST INFRASTRUCTURE ONLY; never linked into or
 * called from the product (chebpol_b200/, libchebpol_cuda).
 *
 * A CPU restatement, in plain C, of the arithmetic of chebpol's batched
 * evaluation path.  It is written from the reference's behaviour, not from
 * its text: every routine is iterative where the reference recurses, but the
 * floating-point operations are issued in the reference's order so that the
 * results are bit-identical to the reference C compiled without FMA
 * contraction (checked against oracle/_ref/libchebpol_ref.so -- the
 * reference's own src/chebpol.c compiled through oracle/shim -- by
 * tests/test_oracle.py, and against the golden values of the reference's
 * saved test output, tests/all.Rout.save and
 * tests/Examples/chebpol-Ex.Rout.save).  Parity status: PINNED.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off).
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))
#define ORACLE_MAXRANK 32

/* ------------------------------------------------------------------------
 * Chebyshev series, nested Clenshaw.
 * Follows C_evalcheb, /root/reference/src/chebpol.c:314-338:
 *  - dimension rank-1 is the outermost recurrence, dimension 0 the innermost;
 *  - innermost (:324-330): b = 2x*b1 - b2 + c[i] for i = N-1..1, result
 *    x*b - b1 + c[0];
 *  - outer levels (:332-337): b = 2x*b1 - b2 + (value of sub-tensor i) for
 *    i = N-1..0, result b - x*b1.
 * The c[0] coefficients carry no 1/2 factor (chebcoef folds it in, :126-140).
 * ------------------------------------------------------------------------ */
static inline double cheb_row(const double *c, int n, double x0)
{
    const double two_x = 2.0 * x0;
    double b = 0.0, b1 = 0.0, b2;
    for (int i = n - 1; i >= 1; --i) {
        b2 = b1;
        b1 = b;
        b = two_x * b1 - b2 + c[i];
    }
    return x0 * b - b1 + c[0];
}

static double cheb_point(const double *cf, const double *x, const int *dims, int rank,
                         const int64_t *stride)
{
    if (rank == 0) return cf[0];
    if (rank == 1) return cheb_row(cf, dims[0], x[0]);
    int idx[ORACLE_MAXRANK];
    double b[ORACLE_MAXRANK], b1[ORACLE_MAXRANK];
    int64_t off = 0;
    for (int l = 1; l < rank; ++l) {
        idx[l] = dims[l] - 1;
        b[l] = b1[l] = 0.0;
        off += (int64_t)idx[l] * stride[l];
    }
    for (;;) {
        double v = cheb_row(cf + off, dims[0], x[0]);
        int l = 1;
        for (;;) {
            const double b2 = b1[l];
            b1[l] = b[l];
            b[l] = 2.0 * x[l] * b1[l] - b2 + v;
            if (idx[l] > 0) {
                --idx[l];
                off -= stride[l];
                break;
            }
            v = b[l] - x[l] * b1[l];
            b[l] = b1[l] = 0.0;
            idx[l] = dims[l] - 1;
            off += (int64_t)idx[l] * stride[l];
            if (++l == rank) return v;
        }
    }
}

/* Batch driver: the loop of R_evalcheb, src/chebpol.c:375-378, with a 64-bit
 * point index.  x is K x P column-major, i.e. point i at x + i*rank. */
API int oracle_evalcheb(const double *cf, const int *dims, int rank, const double *x,
                        int64_t npts, int threads, double *out)
{
    if (rank < 0 || rank > ORACLE_MAXRANK) return 1;
    int64_t stride[ORACLE_MAXRANK + 1];
    stride[0] = 1;
    for (int l = 0; l < rank; ++l) stride[l + 1] = stride[l] * dims[l];
#pragma omp parallel for num_threads(threads) schedule(static) if (npts > 1 && threads > 1)
    for (int64_t i = 0; i < npts; ++i)
        out[i] = cheb_point(cf, x + i * rank, dims, rank, stride);
    return 0;
}

/* ------------------------------------------------------------------------
 * R-level pre-maps that run on every call of the closure.
 *  imap: (x - mid) * ispan           R/chebyshev.R:223-227
 *  ugm : sin(0.5*pi*x*(1-n)/n)       R/uniform.R:6, applied per coordinate
 *        after the optional interval map is*(xi-mid), R/uniform.R:63-72.
 * R evaluates 0.5*pi*x*(1-n)/n left to right in doubles.
 * ------------------------------------------------------------------------ */
API void oracle_imap(const double *x, int rank, int64_t npts, const double *mid,
                     const double *ispan, int threads, double *out)
{
#pragma omp parallel for num_threads(threads) schedule(static) if (npts > 1 && threads > 1)
    for (int64_t i = 0; i < npts; ++i)
        for (int d = 0; d < rank; ++d)
            out[i * rank + d] = (x[i * rank + d] - mid[d]) * ispan[d];
}

API void oracle_ugm(const double *x, int rank, int64_t npts, const int *n, const double *mid,
                    const double *ispan, int threads, double *out)
{
    const double pi = 3.141592653589793;
#pragma omp parallel for num_threads(threads) schedule(static) if (npts > 1 && threads > 1)
    for (int64_t i = 0; i < npts; ++i)
        for (int d = 0; d < rank; ++d) {
            double t = x[i * rank + d];
            if (mid) t = ispan[d] * (t - mid[d]);
            const double nd = (double)n[d];
            out[i * rank + d] = sin(0.5 * pi * t * (1.0 - nd) / nd);
        }
}

/* ------------------------------------------------------------------------
 * The 'general' grid method's pre-map (chebappxg.real, R/chebyshev.R:333-359):
 *   gridmaps <- mapply(stats::splinefun, grid, chebknots(dim(val)), method='hyman')
 *   ch(gridmap(x), threads)   with gridmap applying gridmaps[[d]] to coordinate d,
 *                             one scalar call per coordinate (:351-354).
 * splinefun is R's stats package -- a dependency of the reference that is NOT
 * under /root/reference (the reference's saved test output was produced with
 * R-devel r77544, i.e. the code base of R 4.0.0).  Its published algorithm is
 * restated here:
 *   - spline_coef, method "fmm" (R src/library/stats/src/splines.c, fmm_spline:
 *     the cubic spline of Forsythe, Malcolm & Moler with end conditions from
 *     the third divided differences of the first / last four points);
 *   - hyman_filter (R src/library/stats/R/splinefun.R): the slopes b are
 *     clamped into [0, 3 min(|S_{i-1}|,|S_i|)] (or the mirrored interval),
 *     Hyman (1983) as simplified by Dougherty, Edelman & Hyman (1989);
 *   - spl_coef_conv (same file): c, d recomputed from the filtered slopes;
 *   - spline_eval, method 3 (splines.c): i = 0 is kept when x[0] <= u <= x[1]
 *     (every call of the closure above is a scalar call, so the search always
 *     starts at i = 0), else bisection for the last knot <= u; Horner
 *     y + dx*(b + dx*(c + dx*d)); outside the knots the first / last cubic
 *     piece extrapolates.
 * Parity anchor: the three 'general' values of tests/all.Rout.save:109,121,126
 * (7 digits); tests/test_oracle.py.  Knots must be strictly increasing (R's
 * regularize.values sorts them; the Python mirror does the same before the fit).
 * ------------------------------------------------------------------------ */
API int oracle_hyman_fit(int n, const double *x, const double *y, double *b, double *c, double *d)
{
    if (n < 2) return 1;
    if (n < 3) {
        const double t = y[1] - y[0];
        b[0] = t / (x[1] - x[0]);
        b[1] = b[0];
        c[0] = c[1] = d[0] = d[1] = 0.0;
    } else {
        /* tridiagonal system: b diagonal, d off-diagonal, c right-hand side */
        const int m = n - 1;
        d[0] = x[1] - x[0];
        c[1] = (y[1] - y[0]) / d[0];
        for (int i = 1; i < m; ++i) {
            d[i] = x[i + 1] - x[i];
            b[i] = 2.0 * (d[i - 1] + d[i]);
            c[i + 1] = (y[i + 1] - y[i]) / d[i];
            c[i] = c[i + 1] - c[i];
        }
        b[0] = -d[0];
        b[m] = -d[m - 1];
        c[0] = c[m] = 0.0;
        if (n > 3) {
            c[0] = c[2] / (x[3] - x[1]) - c[1] / (x[2] - x[0]);
            c[m] = c[m - 1] / (x[m] - x[m - 2]) - c[m - 2] / (x[m - 1] - x[m - 3]);
            c[0] = c[0] * d[0] * d[0] / (x[3] - x[0]);
            c[m] = -c[m] * d[m - 1] * d[m - 1] / (x[m] - x[m - 3]);
        }
        for (int i = 1; i <= m; ++i) { /* forward elimination */
            const double t = d[i - 1] / b[i - 1];
            b[i] = b[i] - t * d[i - 1];
            c[i] = c[i] - t * c[i - 1];
        }
        c[m] = c[m] / b[m]; /* back substitution */
        for (int i = m - 1; i >= 0; --i) c[i] = (c[i] - d[i] * c[i + 1]) / b[i];
        b[m] = (y[m] - y[m - 1]) / d[m - 1] + d[m - 1] * (c[m - 1] + 2.0 * c[m]);
        for (int i = 0; i < m; ++i) {
            b[i] = (y[i + 1] - y[i]) / d[i] - d[i] * (c[i + 1] + 2.0 * c[i]);
            d[i] = (c[i + 1] - c[i]) / d[i];
            c[i] = 3.0 * c[i];
        }
        c[m] = 3.0 * c[m];
        d[m] = d[m - 1];
    }
    /* hyman_filter: S0 = c(ss[1], ss), S1 = c(ss, ss[n-1]) with ss the secant slopes */
    for (int i = 0; i < n; ++i) {
        const int il = (i == 0) ? 0 : i - 1, ir = (i == n - 1) ? n - 2 : i;
        const double s0 = (y[il + 1] - y[il]) / (x[il + 1] - x[il]);
        const double s1 = (y[ir + 1] - y[ir]) / (x[ir + 1] - x[ir]);
        const double t1 = fmin(fabs(s0), fabs(s1));
        double sig = b[i];
        if (s0 * s1 > 0) sig = s1;
        if (sig >= 0) b[i] = fmin(fmax(0.0, b[i]), 3.0 * t1);
        else b[i] = fmax(fmin(0.0, b[i]), -3.0 * t1);
    }
    /* spl_coef_conv */
    for (int i = 0; i < n - 1; ++i) {
        const double h = x[i + 1] - x[i], yy = -(y[i + 1] - y[i]);
        c[i] = -(3.0 * yy + (2.0 * b[i] + b[i + 1]) * h) / (h * h);
        d[i] = (2.0 * yy / h + b[i] + b[i + 1]) / (h * h);
    }
    {
        const int i = n - 2;
        const double h = x[i + 1] - x[i], yy = -(y[i + 1] - y[i]);
        c[n - 1] = (3.0 * yy + (b[i] + 2.0 * b[i + 1]) * h) / (h * h);
        d[n - 1] = d[n - 2];
    }
    return 0;
}

static inline double spline_point(int n, const double *x, const double *y, const double *b,
                                  const double *c, const double *d, double u)
{
    int i = 0;
    if (u < x[0] || (n > 1 && x[1] < u)) {
        int j = n;
        do {
            const int k = (i + j) / 2;
            if (u < x[k]) j = k; else i = k;
        } while (j > i + 1);
    }
    const double dx = u - x[i];
    return y[i] + dx * (b[i] + dx * (c[i] + dx * d[i]));
}

/* out[i*rank + k] = gridmaps[[k]](x[i*rank + k]); table pointers per dimension */
API void oracle_splinemap(const double *x, int rank, int64_t npts, const int *sn, const double *const *sx,
                          const double *const *sy, const double *const *sb, const double *const *sc,
                          const double *const *sd, int threads, double *out)
{
#pragma omp parallel for num_threads(threads) schedule(static) if (npts > 1 && threads > 1)
    for (int64_t i = 0; i < npts; ++i)
        for (int k = 0; k < rank; ++k)
            out[i * rank + k] = spline_point(sn[k], sx[k], sy[k], sb[k], sc[k], sd[k], x[i * rank + k]);
}

/* ------------------------------------------------------------------------
 * Multilinear interpolation with blending.
 * Follows binsearch (src/chebpol.c:543-556), blendfun (src/chebpol.h:24-45)
 * and C_evalmlip (src/chebpol.c:617-667).
 * ------------------------------------------------------------------------ */
/* smallest index with arr[idx] >= val, clamped to [1, n-1] (comment at
 * src/chebpol.c:537-542).  Unlike the reference loop (:547-554) this one
 * terminates on NaN: NaN compares false, so it is treated as "too small". */
static inline int cell_upper(int n, const double *arr, double val)
{
    int lo = 0, hi = n - 1;
    while (lo + 1 < hi) {
        const int mid = (lo + hi) >> 1;
        if (arr[mid] >= val) hi = mid;
        else lo = mid;
    }
    return hi;
}

static inline double blend_weight(double w, int blend)
{
    switch (blend) {
    case 1: return w < 0.5 ? 0.5 * exp(2 - 1 / w) : 1 - 0.5 * exp(2 - 1 / (1 - w));
    case 2: return w < 0.5 ? 0.5 * exp(4 - 1 / (w * w)) : 1 - 0.5 * exp(4 - 1 / ((1 - w) * (1 - w)));
    case 3: return (-2 * w + 3) * w * w;
    case 4: return (w < 0.5) ? 0 : 1;
    case 5: return 0.5;
    default: return w; /* 0 = linear */
    }
}

static double mlip_point(int rank, const double *x, const double *const *grid, const int *dims,
                         const double *values, int blend)
{
    double w[ORACLE_MAXRANK];
    int64_t stride[ORACLE_MAXRANK];
    int64_t top = 0, s = 1;
    for (int d = 0; d < rank; ++d) {
        const double *g = grid[d];
        const int gp = cell_upper(dims[d], g, x[d]);
        w[d] = (x[d] - g[gp - 1]) / (g[gp] - g[gp - 1]);
        stride[d] = s;
        top += s * gp;
        s *= dims[d];
    }
    double wsum = 0.0, acc = 0.0;
    const unsigned long ncorner = 1ul << rank;
    for (unsigned long corner = 0; corner < ncorner; ++corner) {
        int64_t pos = top;
        double cw = 1;
        for (int d = 0; d < rank; ++d) {
            if (corner & (1ul << d)) {
                cw *= blend_weight(w[d], blend);
            } else {
                cw *= blend_weight(1 - w[d], blend);
                pos -= stride[d];
            }
        }
        wsum += cw;
        acc += cw * values[pos];
    }
    return acc / wsum;
}

API int oracle_evalmlip(int rank, const double *const *grid, const int *dims, const double *values,
                        const double *x, int64_t npts, int blend, int threads, double *out)
{
    if (rank < 1 || rank > ORACLE_MAXRANK - 1) return 1;
#pragma omp parallel for num_threads(threads) schedule(static) if (npts > 1 && threads > 1)
    for (int64_t i = 0; i < npts; ++i)
        out[i] = mlip_point(rank, x + i * rank, grid, dims, values, blend);
    return 0;
}

API double oracle_blendfun(double w, int blend) { return blend_weight(w, blend); }
API int oracle_binsearch(int n, const double *arr, double val) { return cell_upper(n, arr, val); }

/* ------------------------------------------------------------------------
 * Floater-Hormann rational interpolation, nested barycentric form.
 * Follows FH, src/chebpol.c:182-218: dimension rank-1 outermost; at each
 * level, first i with |x - knot_i| < 10*DBL_EPSILON (absolute) short-cuts to
 * slice i (:195-197); otherwise sum(pole*val)/sum(pole) with
 * pole = w_i/(x - knot_i), accumulated for i ascending (:211-217).
 * Iterative: an explicit frame per level replaces the recursion.
 * ------------------------------------------------------------------------ */
static double fh_point(const double *fv, const double *x, const double *const *knots,
                       const int *dims, int rank, const double *const *weights,
                       const int64_t *stride)
{
    if (rank == 0) return fv[0];
    /* frame of level l (the recursion with rank == l+1) */
    int i[ORACLE_MAXRANK], hit[ORACLE_MAXRANK];
    double num[ORACLE_MAXRANK], den[ORACLE_MAXRANK];
    const double *base[ORACLE_MAXRANK];
    int l = rank - 1;
    base[l] = fv;
    double ret = 0.0;
    int entering = 1;
    for (;;) {
        if (entering) {
            /* open level l on sub-tensor base[l] */
            const double xx = x[l];
            const double *kn = knots[l];
            hit[l] = -1;
            for (int k = 0; k < dims[l]; ++k)
                if (fabs(xx - kn[k]) < 10.0 * DBL_EPSILON) { hit[l] = k; break; }
            num[l] = den[l] = 0.0;
            i[l] = (hit[l] >= 0) ? hit[l] : 0;
            if (l == 0) {
                if (hit[0] >= 0) {
                    ret = base[0][hit[0]];
                } else {
                    double nu = 0, de = 0;
                    const double *w = weights[0];
                    for (int k = 0; k < dims[0]; ++k) {
                        const double val = base[0][k];
                        const double pole = w[k] / (xx - kn[k]);
                        nu += pole * val;
                        de += pole;
                    }
                    ret = nu / de;
                }
                entering = 0;
                ++l;
                if (l == rank) return ret;
                continue;
            }
            base[l - 1] = base[l] + (int64_t)i[l] * stride[l];
            --l;
            continue;
        }
        /* returning into level l with the value `ret` of slice i[l] */
        if (hit[l] >= 0) {
            /* knot hit: the level's value is the slice value */
        } else {
            const double pole = weights[l][i[l]] / (x[l] - knots[l][i[l]]);
            num[l] += pole * ret;
            den[l] += pole;
            if (++i[l] < dims[l]) {
                base[l - 1] = base[l] + (int64_t)i[l] * stride[l];
                --l;
                entering = 1;
                continue;
            }
            ret = num[l] / den[l];
        }
        ++l;
        if (l == rank) return ret;
    }
}

API int oracle_fh(const double *vals, const double *x, const double *const *knots, const int *dims,
                  int rank, const double *const *weights, int64_t npts, int threads, double *out)
{
    if (rank < 0 || rank > ORACLE_MAXRANK) return 1;
    int64_t stride[ORACLE_MAXRANK + 1];
    stride[0] = 1;
    for (int l = 0; l < rank; ++l) stride[l + 1] = stride[l] * dims[l];
#pragma omp parallel for num_threads(threads) schedule(static) if (npts > 1 && threads > 1)
    for (int64_t p = 0; p < npts; ++p)
        out[p] = fh_point(vals, x + p * rank, knots, dims, rank, weights, stride);
    return 0;
}

/* ------------------------------------------------------------------------
 * Polyharmonic spline evaluation.
 * Follows C_evalpolyh, src/chebpol.c:383-415, and the loop of R_evalpolyh, :443-446:
 *   f(x) = sum_i w_i phi(|x - k_i|) + v_0 + sum_j v_{j+1} x_j
 *   phi(r) = exp(k r^2)        k < 0            (:388-394)
 *            r^k               k odd  (int)k    (:395-402, knots at distance 0 skipped)
 *            log(r) r^k        k even           (:403-411, likewise)
 * knots is rank x nknots column-major (knot i at knots + i*rank); the sum runs over the knots
 * in order; r^k is R's R_pow_di (repeated squaring, R src/main/arithmetic.c).
 * ------------------------------------------------------------------------ */
static double oracle_pow_di(double x, int n)
{
    double xn = 1.0;
    if (isnan(x)) return x;
    if (n != 0) {
        if (!isfinite(x)) return pow(x, (double)n);
        const int neg = n < 0;
        if (neg) n = -n;
        for (;;) {
            if (n & 1) xn *= x;
            n >>= 1;
            if (!n) break;
            x *= x;
        }
        if (neg) xn = 1.0 / xn;
    }
    return xn;
}

static double polyh_point(const double *x, const double *knots, const double *w, const double *lw,
                          int rank, int nknots, double k)
{
    const int ki = (int)k;
    const int mode = (k < 0) ? 0 : ((ki % 2 == 1) ? 1 : 2);
    double sum = 0.0;
    for (int i = 0; i < nknots; ++i) {
        const double *kn = knots + (size_t)i * rank;
        double d2 = 0.0;
        for (int j = 0; j < rank; ++j) d2 += (x[j] - kn[j]) * (x[j] - kn[j]);
        if (mode == 0) {
            sum += w[i] * exp(k * d2);
        } else if (d2 != 0) {
            if (mode == 1) sum += w[i] * oracle_pow_di(sqrt(d2), ki);
            else sum += w[i] * log(sqrt(d2)) * oracle_pow_di(sqrt(d2), ki);
        }
    }
    sum += lw[0];
    for (int j = 0; j < rank; ++j) sum += lw[j + 1] * x[j];
    return sum;
}

API int oracle_evalpolyh(const double *x, int64_t npts, const double *knots, const double *weights,
                         const double *lweights, int rank, int nknots, double k, int threads, double *out)
{
    if (rank <= 0 || nknots < 0) return 1;
#pragma omp parallel for num_threads(threads) schedule(static) if (npts > 1 && threads > 1)
    for (int64_t p = 0; p < npts; ++p)
        out[p] = polyh_point(x + p * rank, knots, weights, lweights, rank, nknots, k);
    return 0;
}

/* ------------------------------------------------------------------------
 * Stalker splines: src/stalker.c built WITHOUT GSL (HAVE_GSL
 * undefined, what configure produces when GSL is absent -- as in this image): non-uniform
 * grids then use degree r = 2 in makestalk (:340-347,355-359).
 *   makestalk :271-364, evalstalk :366-442   ("stalker":  f_corner = v + sum_j b_j z_j + c_j |z_j|^r_j)
 *   makehyp   :446-535, evalhyp   :537-613   ("hstalker": f_corner = v + a + sum_j b_j z_j + d_j/(1 + c_j z_j),
 *                                             c_j = NA marks the parabola b_j z_j^2)
 * Tables are rank x prod(dims), column-major (entry j of grid point `index` at [rank*index + j]).
 * The cell is R's findInterval(rightmost_closed = all_inside = TRUE): gp with g[gp] <= x < g[gp+1],
 * clamped to 0..n-2; weights clamped to [0,1] (linear extrapolation); corners weighted by
 * blendfun (src/chebpol.h:24-45), corners of weight zero skipped.
 * ------------------------------------------------------------------------ */
#define ORACLE_MYEPS (1e3 * sqrt(DBL_EPSILON))
static double oracle_sign(double x) { return isnan(x) ? x : ((x > 0) ? 1.0 : ((x == 0) ? 0.0 : -1.0)); }
static double oracle_na_real(void)
{
    union { double d; unsigned long long u; } v;
    v.u = 0x7FF00000000007A2ULL; /* R's NA_real_ */
    return v.d;
}

static int find_interval0(const double *g, int n, double x)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) / 2;
        if (g[mid] <= x) lo = mid + 1; else hi = mid;
    }
    if (lo < 1) lo = 1;
    if (lo > n - 1) lo = n - 1;
    return lo - 1;
}

API int oracle_makestalk(int rank, const int *dims, const double *const *grid, const double *val,
                         double *b, double *c, double *r)
{
    if (rank <= 0 || rank > ORACLE_MAXRANK) return 1;
    int64_t len = 1;
    for (int i = 0; i < rank; ++i) len *= dims[i];
    for (int64_t index = 0; index < len; ++index) {
        double *bb = b + rank * index, *cc = c + rank * index, *rr = r + rank * index;
        int64_t stride = 1, rem = index;
        for (int i = 0; i < rank; ++i) {
            const double *gr = grid[i];
            const int idx = (int)(rem % dims[i]);
            rem /= dims[i];
            if (idx == 0) {
                bb[i] = (val[index + stride] - val[index]) / (gr[idx + 1] - gr[idx]);
                cc[i] = rr[i] = 0.0;
            } else if (idx == dims[i] - 1) {
                bb[i] = (val[index - stride] - val[index]) / (gr[idx - 1] - gr[idx]);
                cc[i] = rr[i] = 0.0;
            } else {
                const double v0 = val[index];
                const double vplus = val[index + stride] - v0, vmin = val[index - stride] - v0;
                const double kplus = gr[idx + 1] - gr[idx], kmin = gr[idx - 1] - gr[idx];
                const int uniform = fabs(kplus + kmin) < ORACLE_MYEPS * kplus;
                double rv;
                if (oracle_sign(vplus * vmin) <= 0) {
                    if (fabs(vplus - vmin) <= ORACLE_MYEPS * kplus) {
                        rv = 2.0;
                        bb[i] = cc[i] = 0.0;
                    } else if (fabs(kmin * vplus - kplus * vmin) <= ORACLE_MYEPS * kplus) {
                        rv = 2.0;
                    } else if (uniform) {
                        rv = fabs((vplus - vmin) / (vplus + vmin));
                        if (rv < 1) rv = 1;
                        if (isnan(rv) || rv > 2.0) rv = 2.0;
                    } else {
                        rv = 2.0; /* no GSL */
                    }
                } else {
                    rv = 2.0; /* no GSL */
                }
                rr[i] = rv;
                bb[i] = (vplus - vmin * pow(-kplus / kmin, rv)) / (kplus - kmin * pow(-kplus / kmin, rv));
                cc[i] = (kplus * vmin - kmin * vplus) / (kplus * pow(-kmin, rv) - kmin * pow(kplus, rv));
            }
            stride *= dims[i];
        }
    }
    return 0;
}

API int oracle_makehyp(int rank, const int *dims, const double *const *grid, const double *val,
                       double *a, double *b, double *c, double *d)
{
    if (rank <= 0 || rank > ORACLE_MAXRANK) return 1;
    int64_t len = 1;
    for (int i = 0; i < rank; ++i) len *= dims[i];
    for (int64_t index = 0; index < len; ++index) {
        double *bb = b + rank * index, *cc = c + rank * index, *dd = d + rank * index;
        a[index] = 0.0;
        int64_t stride = 1, rem = index;
        for (int i = 0; i < rank; ++i) {
            const double *gr = grid[i];
            const int idx = (int)(rem % dims[i]);
            rem /= dims[i];
            if (idx == 0) {
                bb[i] = (val[index + stride] - val[index]) / (gr[idx + 1] - gr[idx]);
                cc[i] = dd[i] = 0.0;
            } else if (idx == dims[i] - 1) {
                bb[i] = (val[index - stride] - val[index]) / (gr[idx - 1] - gr[idx]);
                cc[i] = dd[i] = 0.0;
            } else {
                const double v0 = val[index];
                const double vplus = val[index + stride] - v0, vmin = val[index - stride] - v0;
                const double kplus = gr[idx + 1] - gr[idx], kmin = gr[idx - 1] - gr[idx];
                const double D = kmin * vplus - kplus * vmin;
                const double D2 = kmin * kmin * vplus - kplus * kplus * vmin;
                if (oracle_sign(vplus * vmin) <= 0 || fabs(vplus) <= ORACLE_MYEPS || fabs(vmin) <= ORACLE_MYEPS) {
                    if (fabs(vplus - vmin) <= ORACLE_MYEPS * kplus) {
                        bb[i] = cc[i] = dd[i] = 0.0;
                    } else if (fabs(D) <= ORACLE_MYEPS * kplus) {
                        bb[i] = vplus / kplus;
                        cc[i] = dd[i] = 0.0;
                    } else {
                        bb[i] = 0.0;
                        cc[i] = -D / (kmin * kplus * (vplus - vmin));
                        dd[i] = -vplus * vmin * (kmin - kplus) / D;
                    }
                } else if (fabs(D2) <= ORACLE_MYEPS * kplus) {
                    bb[i] = vplus / (kplus * kplus);
                    dd[i] = 0.0;
                    cc[i] = oracle_na_real();
                } else {
                    bb[i] = vplus * vmin * (kmin - kplus) / D2;
                    cc[i] = -D2 / (kmin * kplus * D);
                    dd[i] = -kmin * vplus * kplus * vmin * (kmin - kplus) * D / (D2 * D2);
                }
            }
            a[index] -= dd[i];
            stride *= dims[i];
        }
    }
    return 0;
}

/* kind 0: stalker (t1 = c, t2 = r, a unused); kind 1: hstalker (t1 = c, t2 = d) */
static double stalk_point(int kind, const double *x, int rank, const int *dims, const double *const *grid,
                          const double *val, int blend, const double *a, const double *b, const double *t1,
                          const double *t2)
{
    double weight[ORACLE_MAXRANK], zx[ORACLE_MAXRANK];
    int gbase[ORACLE_MAXRANK];
    int64_t llcorner = 0, stride = 1;
    for (int g = 0; g < rank; ++g) {
        const double *gr = grid[g];
        const int gp = find_interval0(gr, dims[g], x[g]);
        gbase[g] = gp;
        weight[g] = 1 - (x[g] - gr[gp]) / (gr[gp + 1] - gr[gp]);
        if (weight[g] < 0) weight[g] = 0;
        if (weight[g] > 1) weight[g] = 1;
        llcorner += stride * gp;
        stride *= dims[g];
    }
    double V = 0.0, sumw = 0.0;
    for (int i = 0; i < (1 << rank); ++i) {
        int64_t corner = llcorner;
        double cw = 1.0;
        stride = 1;
        for (int g = 0; g < rank; ++g) {
            const double *gr = grid[g];
            if (i & (1 << g)) {
                cw *= oracle_blendfun(1 - weight[g], blend);
                corner += stride;
                zx[g] = x[g] - gr[gbase[g] + 1];
            } else {
                zx[g] = x[g] - gr[gbase[g]];
                cw *= oracle_blendfun(weight[g], blend);
            }
            stride *= dims[g];
        }
        if (cw == 0.0) continue;
        const double *bb = b + rank * corner, *p1 = t1 + rank * corner, *p2 = t2 + rank * corner;
        double fval;
        if (kind == 0) {
            fval = val[corner];
            for (int j = 0; j < rank; ++j) {
                if (p1[j] == 0) fval += bb[j] * zx[j];
                else fval += bb[j] * zx[j] + p1[j] * pow(fabs(zx[j]), p2[j]);
            }
        } else {
            fval = val[corner] + a[corner];
            for (int j = 0; j < rank; ++j) {
                if (!isnan(p1[j])) fval += bb[j] * zx[j] + p2[j] / (1.0 + p1[j] * zx[j]);
                else fval += bb[j] * zx[j] * zx[j];
            }
        }
        sumw += cw;
        V += cw * fval;
    }
    return V / sumw;
}

API int oracle_evalstalk(const double *x, int64_t npts, int rank, const int *dims, const double *const *grid,
                         const double *val, int blend, const double *b, const double *c, const double *r,
                         int threads, double *out)
{
    if (rank <= 0 || rank > 30) return 1;
#pragma omp parallel for num_threads(threads) schedule(guided) if (npts > 1 && threads > 1)
    for (int64_t p = 0; p < npts; ++p)
        out[p] = stalk_point(0, x + p * rank, rank, dims, grid, val, blend, NULL, b, c, r);
    return 0;
}

API int oracle_evalhyp(const double *x, int64_t npts, int rank, const int *dims, const double *const *grid,
                       const double *val, int blend, const double *a, const double *b, const double *c,
                       const double *d, int threads, double *out)
{
    if (rank <= 0 || rank > 30) return 1;
#pragma omp parallel for num_threads(threads) schedule(guided) if (npts > 1 && threads > 1)
    for (int64_t p = 0; p < npts; ++p)
        out[p] = stalk_point(1, x + p * rank, rank, dims, grid, val, blend, a, b, c, d);
    return 0;
}

/* ------------------------------------------------------------------------
 * Fit-side helpers (fixtures only; not on the timed path).
 * ------------------------------------------------------------------------ */
/* Floater-Hormann weights, eq. (18) of the FH paper, as fhweights at
 * src/chebpol.c:267-282 computes them: for knot k (0..n, n = nknots-1)
 *   w_k = (-1)^(k-d) * sum_{i=max(0,k-d)}^{min(k,n-d)} 1/prod_{j=i..i+d, j!=k} |x_k-x_j|
 * where the product is taken signed and its absolute value inverted. */
API void oracle_fhweights(int nknots, const double *grid, int d, double *w)
{
    const int n = nknots - 1;
    for (int k = 0; k <= n; ++k) {
        const int lo = (k < d) ? 0 : k - d;
        const int hi = (k < n - d) ? k : n - d;
        double sum = 0.0;
        for (int i = lo; i <= hi; ++i) {
            double prod = 1.0;
            for (int j = i; j <= i + d; ++j)
                if (j != k) prod *= grid[k] - grid[j];
            sum += 1.0 / fabs(prod);
        }
        w[k] = sum * ((k % 2 != d % 2) ? -1.0 : 1.0);
    }
}

/* cos(pi*x), exact at multiples of 1/2 (R's cospi, used at src/chebpol.c:94) */
static double cos_pi(double x)
{
    x = fmod(fabs(x), 2.0);
    if (fmod(x, 1.0) == 0.5) return 0.0;
    if (x == 1.0) return -1.0;
    if (x == 0.0) return 1.0;
    return cos(M_PI * x);
}

/* Chebyshev coefficients from values on the Chebyshev grid: the matrix path
 * of chebcoef, src/chebpol.c:65-140.  Per dimension (last first, :109):
 *   out[k] = alpha * sum_i cos(pi*k*(i+0.5)/N) * in[i],  alpha = 2/N (2 if dct)
 * and afterwards, unless dct, every index-0 hyperplane is halved (:126-140).
 * The transform of one dimension rotates it to the front, exactly as the
 * reference's dgemm("T","T") call does (:119), so after `rank` steps the
 * array is back in its original order and the roundings match. */
API void oracle_chebcoef(const double *val, const int *dims, int rank, double *F, int dct)
{
    int64_t siz = 1;
    for (int i = 0; i < rank; ++i) siz *= dims[i];
    double *a = (double *)malloc(sizeof(double) * siz);
    double *b = (double *)malloc(sizeof(double) * siz);
    memcpy(a, val, sizeof(double) * siz);
    for (int d = rank - 1; d >= 0; --d) {
        const int N = dims[d];
        const int64_t rest = siz / N;
        const double alpha = dct ? 2.0 : 2.0 / N;
        double *cm = (double *)malloc(sizeof(double) * N * N);
        for (int k = 0; k < N; ++k)
            for (int i = 0; i < N; ++i) cm[i + (int64_t)k * N] = cos_pi(k * (i + 0.5) / N);
        /* a is (rest x N) column-major with the transformed dim last; b becomes (N x rest) */
        for (int64_t j = 0; j < rest; ++j)
            for (int k = 0; k < N; ++k) {
                double s = 0.0;
                for (int i = 0; i < N; ++i) s += cm[i + (int64_t)k * N] * a[j + (int64_t)i * rest];
                b[k + (int64_t)j * N] = alpha * s;
            }
        free(cm);
        double *t = a; a = b; b = t;
    }
    memcpy(F, a, sizeof(double) * siz);
    free(a);
    free(b);
    if (!dct) {
        int64_t blocklen = 1, stride = 1;
        for (int i = 0; i < rank; ++i) {
            stride *= dims[i];
            for (int64_t j = 0; j < siz; j += stride)
                for (int64_t s = 0; s < blocklen; ++s) F[j + s] *= 0.5;
            blocklen = stride;
        }
    }
}

/* ------------------------------------------------------------------------
 * Simplex-linear interpolation on a Delaunay triangulation (slappx.real,
 * R/simplexlinear.R:27-40).
 *   analyzesimplex  R_analyzesimplex, src/chebpol.c:830-875: per simplex the matrix
 *                   [vertices; 1 ... 1] (column v = vertex v, last row ones),
 *                   LU-factorised by LAPACK dgetrf, and the bounding box;
 *   evalsl          R_evalsl :803-828 over findsimplex :738-801: the FIRST simplex
 *                   (index order) whose bounding box holds x and whose barycentric
 *                   coordinates (dgetrs) are all >= -1e-10; blended weights
 *                   (blendfun, src/chebpol.h:24-45) normalised by their sum; outside
 *                   the hull NA, or with epol: the first simplex containing the
 *                   nearest knot, plain barycentric combination.
 * dgetrf/dgetrs are LAPACK (not under /root/reference); the published reference
 * algorithm -- first-largest pivot, reciprocal-scaled multipliers, row-swapped
 * right-hand side, unit-lower then upper substitution skipping zero entries -- is
 * what lu_factor / lu_solve below issue, operation for operation.
 * dtri: (dim+1) x numsimplex, 1-based knot indices; knots: dim x nknots.
 * ------------------------------------------------------------------------ */
static void lu_factor(int n, double *a, int *ipiv)
{
    for (int j = 0; j < n; ++j) {
        int jp = j;
        double mx = fabs(a[j + j * n]);
        for (int i = j + 1; i < n; ++i)
            if (fabs(a[i + j * n]) > mx) { mx = fabs(a[i + j * n]); jp = i; }
        ipiv[j] = jp + 1;
        if (a[jp + j * n] != 0.0) {
            if (jp != j)
                for (int c = 0; c < n; ++c) { const double t = a[j + c * n]; a[j + c * n] = a[jp + c * n]; a[jp + c * n] = t; }
            if (j < n - 1) {
                if (fabs(a[j + j * n]) >= DBL_MIN) {
                    const double r = 1.0 / a[j + j * n];
                    for (int i = j + 1; i < n; ++i) a[i + j * n] *= r;
                } else
                    for (int i = j + 1; i < n; ++i) a[i + j * n] /= a[j + j * n];
            }
        }
        if (j < n - 1)
            for (int c = j + 1; c < n; ++c) {
                const double y = a[j + c * n];
                if (y != 0.0) {
                    const double t = -y;
                    for (int i = j + 1; i < n; ++i) a[i + c * n] += a[i + j * n] * t;
                }
            }
    }
}

static void lu_solve(int n, const double *a, const int *ipiv, double *b)
{
    for (int i = 0; i < n; ++i) {
        const int ip = ipiv[i] - 1;
        if (ip != i) { const double t = b[i]; b[i] = b[ip]; b[ip] = t; }
    }
    for (int k = 0; k < n; ++k)
        if (b[k] != 0.0)
            for (int i = k + 1; i < n; ++i) b[i] = b[i] - b[k] * a[i + k * n];
    for (int k = n - 1; k >= 0; --k)
        if (b[k] != 0.0) {
            b[k] = b[k] / a[k + k * n];
            for (int i = 0; i < k; ++i) b[i] = b[i] - b[k] * a[i + k * n];
        }
}

API int oracle_analyzesimplex(const int *dtri, int numsimplex, const double *knots, int dim, int nknots,
                              double *lumats, int *pivots, double *bbox)
{
    (void)nknots;
    if (dim < 1 || dim >= ORACLE_MAXRANK) return 1;
    const int N = dim + 1;
    for (int s = 0; s < numsimplex; ++s) {
        const int *tri = dtri + (int64_t)s * N;
        double *lu = lumats + (int64_t)s * N * N, *m = bbox + (int64_t)s * 2 * dim;
        for (int j = 0; j < dim; ++j) { m[2 * j] = 1e100; m[2 * j + 1] = -1e100; }
        for (int v = 0; v < N; ++v) {
            const double *knot = knots + (int64_t)(tri[v] - 1) * dim;
            for (int j = 0; j < dim; ++j) {
                lu[v * N + j] = knot[j];
                if (knot[j] < m[2 * j]) m[2 * j] = knot[j];
                if (knot[j] > m[2 * j + 1]) m[2 * j + 1] = knot[j];
            }
            lu[v * N + dim] = 1.0;
        }
        lu_factor(N, lu, pivots + (int64_t)s * N);
    }
    return 0;
}

static double sl_point(const double *x, const double *knots, const int *dtri, const double *lumats,
                       const int *pivots, const double *bbox, int epol, const double *val, int dim,
                       int numsimplex, int nknots, int blend)
{
    double vec[ORACLE_MAXRANK + 1];
    const int N = dim + 1;
    for (int s = 0; s < numsimplex; ++s) {
        const double *box = bbox + (int64_t)s * 2 * dim;
        int bad = 0;
        for (int i = 0; i < dim; ++i)
            if (x[i] < box[2 * i] || x[i] > box[2 * i + 1]) { bad = 1; break; }
        if (bad) continue;
        for (int d = 0; d < dim; ++d) vec[d] = x[d];
        vec[dim] = 1.0;
        lu_solve(N, lumats + (int64_t)s * N * N, pivots + (int64_t)s * N, vec);
        for (int d = 0; d < N; ++d)
            if (vec[d] < -1e-10) { bad = 1; break; }
        if (bad) continue;
        const int *tri = dtri + (int64_t)s * N;
        double sum = 0.0, wsum = 0.0;
        for (int d = 0; d < N; ++d) {
            const double w = blend_weight(vec[d], blend);
            wsum += w;
            sum += w * val[tri[d] - 1];
        }
        return sum / wsum;
    }
    if (epol) {
        double mindist = 1e99;
        int near = 0;
        for (int k = 0; k < nknots; ++k) {
            double dist = 0.0;
            for (int i = 0; i < dim; ++i) dist += (x[i] - knots[(int64_t)k * dim + i]) * (x[i] - knots[(int64_t)k * dim + i]);
            if (dist < mindist) { mindist = dist; near = k; }
        }
        int nearest = -1;
        for (int s = 0; s < numsimplex && nearest < 0; ++s)
            for (int j = 0; j < N; ++j)
                if (dtri[(int64_t)s * N + j] == near + 1) { nearest = s; break; }
        if (nearest < 0) return oracle_na_real();
        for (int d = 0; d < dim; ++d) vec[d] = x[d];
        vec[dim] = 1.0;
        lu_solve(N, lumats + (int64_t)nearest * N * N, pivots + (int64_t)nearest * N, vec);
        const int *tri = dtri + (int64_t)nearest * N;
        double sum = 0.0;
        for (int d = 0; d <= dim; ++d) sum += vec[d] * val[tri[d] - 1];
        return sum;
    }
    return oracle_na_real();
}

API int oracle_evalsl(const double *x, int64_t npts, const double *knots, int dim, int nknots, const int *dtri,
                      int numsimplex, const double *lumats, const int *pivots, const double *bbox,
                      const double *val, int epol, int threads, int blend, double *out)
{
    if (dim < 1 || dim >= ORACLE_MAXRANK) return 1;
#pragma omp parallel for num_threads(threads) schedule(guided) if (npts > 1 && threads > 1)
    for (int64_t i = 0; i < npts; ++i)
        out[i] = sl_point(x + i * dim, knots, dtri, lumats, pivots, bbox, epol, val, dim, numsimplex, nknots, blend);
    return 0;
}

/* index (0-based) of the simplex findsimplex accepts for each point, -1 outside the hull: the
 * scan of sl_point without the interpolation.  Lets tests check candidate-list structures. */
API int oracle_sl_index(const double *x, int64_t npts, int dim, int numsimplex, const double *lumats,
                        const int *pivots, const double *bbox, int threads, int *out)
{
    if (dim < 1 || dim >= ORACLE_MAXRANK) return 1;
    const int N = dim + 1;
#pragma omp parallel for num_threads(threads) schedule(guided) if (npts > 1 && threads > 1)
    for (int64_t i = 0; i < npts; ++i) {
        const double *xi = x + i * dim;
        double vec[ORACLE_MAXRANK + 1];
        int hit = -1;
        for (int s = 0; s < numsimplex && hit < 0; ++s) {
            const double *box = bbox + (int64_t)s * 2 * dim;
            int bad = 0;
            for (int k = 0; k < dim; ++k)
                if (xi[k] < box[2 * k] || xi[k] > box[2 * k + 1]) { bad = 1; break; }
            if (bad) continue;
            for (int d = 0; d < dim; ++d) vec[d] = xi[d];
            vec[dim] = 1.0;
            lu_solve(N, lumats + (int64_t)s * N * N, pivots + (int64_t)s * N, vec);
            for (int d = 0; d < N; ++d)
                if (vec[d] < -1e-10) { bad = 1; break; }
            if (!bad) hit = s;
        }
        out[i] = hit;
    }
    return 0;
}
