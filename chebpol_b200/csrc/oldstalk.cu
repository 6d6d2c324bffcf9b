// oldstalk.cu -- the deprecated recursive stalker evaluator, ipol(method='oldstalker'):
// R_evalstalker / evalstalker / stalk1 / stalkhyp, src/stalker.c:134-260, 617-678, 834-891.
// One thread per point, the reference's operation order (compiled with -fmad=false); the recursion
// over the dimensions (last dimension outermost, up to four lower-dimensional values per level) is
// unrolled into an explicit stack so no device call stack is needed.  Differences from the reference
// C are confined to pow for fractional degrees and exp (device libm, <= 2 ulp; |x|^1 and |x|^2 are formed
// exactly).  As in a reference build without GSL, a NaN
// degree (per-basis degree) is supported on uniform grids only.
#include "common.cuh"
#include <float.h>

struct OldStalkArgs {
    const double *val;    // prod(dims), R layout
    const double *grid;   // knots of all dimensions, concatenated at goff[]
    const double *det, *pmin, *pplus;  // stalkercontext vectors, same offsets
    int goff[CP_MAXR];
    int rank;
    int dims[CP_MAXR];
    int64_t stride[CP_MAXR];
    double degree[CP_MAXR];
    int uniform[CP_MAXR];
    int blend;
    const double *x;
    int64_t npts;
    double *out;
};

__device__ __forceinline__ double os_sign(double x) { return (x != x) ? x : ((x > 0) ? 1.0 : ((x == 0) ? 0.0 : -1.0)); }

__device__ __forceinline__ double os_blend(double w, int blend)
{
    switch (blend) {
    case 1: return w < 0.5 ? 0.5 * exp(2 - 1 / w) : 1 - 0.5 * exp(2 - 1 / (1 - w));
    case 2: return w < 0.5 ? 0.5 * exp(4 - 1 / (w * w)) : 1 - 0.5 * exp(4 - 1 / ((1 - w) * (1 - w)));
    case 3: return (-2 * w + 3) * w * w;
    case 4: return (w < 0.5) ? 0.0 : 1.0;
    case 5: return 0.5;
    default: return w;
    }
}

// stalkhyp, src/stalker.c:134-187 (the #if 1 branch)
__device__ double os_stalkhyp(double x, double vmin, double vplus, double dmin, double dplus)
{
    if (fabs(x - dplus) < 1e3 * DBL_EPSILON * dplus) return vplus;
    if (fabs(x + dmin) < 1e3 * DBL_EPSILON * dmin) return vmin;
    if (fabs(vplus * dmin + vmin * dplus) < 1e3 * DBL_EPSILON * dmin * dplus) return x * (vplus - vmin) / (dplus + dmin);
    if (os_sign(vplus * vmin) < 0) {
        if (fabs(vmin - vplus) < 1e3 * DBL_EPSILON) return 0;
        const double c = (vmin * dplus + vplus * dmin) / (dmin * dplus * (vmin - vplus));
        const double a = vplus + vplus / (c * dplus);
        return a - a / (1 + c * x);
    }
    const double D = dplus * dplus * vmin - dmin * dmin * vplus;
    if (fabs(D) < 1e3 * DBL_EPSILON) {
        const double a = vmin / (dmin * dmin);
        return a * x * x;
    }
    const double c = D / (dmin * dplus * (dmin * vplus + dplus * vmin));
    const double b = vplus * (1 + c * dplus) / (c * dplus * dplus);
    const double a = -b / c;
    return a + b * x - a / (1 + c * x);
}

// |x|^r as libm's pow returns it where the result is exact (r = 1) or a single rounding (r = 2); the
// device pow is only accurate to 2 ulp, and outside the grid the blend amplifies every ulp
__device__ __forceinline__ double os_powabs(double x, double r)
{
    const double ax = fabs(x);
    if (r == 1.0) return ax;
    if (r == 2.0) return ax * ax;
    return pow(ax, r);
}

// stalk1, src/stalker.c:190-260 without GSL
__device__ double os_stalk1(double x, double vmin, double v0, double vplus, double dmin, double dplus, double r,
                            double iD, double powmin, double powplus)
{
    if (vmin != vmin) return v0 + x / dplus * (vplus - v0);
    if (vplus != vplus) return v0 + x / dmin * (v0 - vmin);
    vmin -= v0;
    vplus -= v0;
    if (r == 0) return v0 + os_stalkhyp(x, vmin, vplus, dmin, dplus);
    if (r == r) {
        const double b = (vplus * powmin - vmin * powplus) * iD;
        const double c = (vplus * dmin + vmin * dplus) * iD;
        return v0 + b * x + c * os_powabs(x, r);
    }
    // per-basis degree on a uniform grid
    x /= dplus;
    const double b = 0.5 * (vplus - vmin);
    const double c = 0.5 * (vplus + vmin);
    if (c == 0.0) return v0 + b * x;
    const double ac = fabs(c), ab = fabs(b);
    if (os_sign(vplus * vmin) <= 0) r = (ab < 2.0 * ac) ? ab / ac : 2.0;
    else r = (ac < 2.0 * ab) ? ac / ab : 2.0;
    if (r == 2.0) return v0 + b * x + c * x * x;
    if (r != 1.0) return v0 + b * x + c * os_powabs(x, r);
    return v0 + b * x + c * fabs(x);
}

// R's findInterval(xt, n, x, rightmost_closed = TRUE, all_inside = TRUE) - 1 (0-based left knot)
__device__ __forceinline__ int os_interval(const double *__restrict__ xt, int n, double x)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(xt + mid) <= x) lo = mid + 1;
        else hi = mid;
    }
    if (lo < 1) lo = 1;
    if (lo > n - 1) lo = n - 1;
    return lo - 1;
}

__global__ void __launch_bounds__(128) oldstalk_kernel(const __grid_constant__ OldStalkArgs a)
{
    const int rank = a.rank;
    const double na = __longlong_as_double(0x7FF00000000007A2LL);  // NA_real_
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < a.npts; p += (int64_t)gridDim.x * blockDim.x) {
        const double *x = a.x + p * rank;
        // explicit stack: frame l evaluates the (l+1)-dimensional spline of the sub-array at base[l]
        int imin[CP_MAXR], k[CP_MAXR];
        int64_t base[CP_MAXR];
        double v[CP_MAXR][4];
        int l = rank - 1;
        base[l] = 0;
        k[l] = -1;
        double ret = 0.0;
        bool have_ret = false;
        while (l < rank) {
            const int n = a.dims[l];
            const double *gr = a.grid + a.goff[l];
            const int64_t st = a.stride[l];
            if (k[l] < 0) {  // entry
                imin[l] = os_interval(gr, n, x[l]);
                v[l][0] = v[l][3] = na;
                if (l == 0) {  // the reference saves the last recursion step (:646-651)
                    const int i = imin[l];
                    if (i > 0) v[l][0] = __ldg(a.val + base[l] + (int64_t)(i - 1) * st);
                    v[l][1] = __ldg(a.val + base[l] + (int64_t)i * st);
                    v[l][2] = __ldg(a.val + base[l] + (int64_t)(i + 1) * st);
                    if (i < n - 2) v[l][3] = __ldg(a.val + base[l] + (int64_t)(i + 2) * st);
                    k[l] = 4;
                } else
                    k[l] = 0;
            } else if (have_ret) {  // a child has just returned
                v[l][k[l]] = ret;
                have_ret = false;
                ++k[l];
            }
            // next child to evaluate (v1 only if imin > 0, v4 only if imin < n-2)
            while (k[l] < 4 && ((k[l] == 0 && imin[l] == 0) || (k[l] == 3 && imin[l] >= n - 2))) ++k[l];
            if (k[l] < 4) {
                base[l - 1] = base[l] + (int64_t)(imin[l] - 1 + k[l]) * st;
                k[l - 1] = -1;
                --l;
                continue;
            }
            // all values present: the two basis functions and their blend (:653-677)
            const int i = imin[l];
            const double xx = x[l];
            const double *det = a.det + a.goff[l], *pmn = a.pmin + a.goff[l], *ppl = a.pplus + a.goff[l];
            double dmin = na;
            if (i > 0) dmin = __ldg(gr + i) - __ldg(gr + i - 1);
            double dplus = __ldg(gr + i + 1) - __ldg(gr + i);
            const double low = os_stalk1(xx - __ldg(gr + i), v[l][0], v[l][1], v[l][2], dmin, dplus, a.degree[l],
                                         __ldg(det + i), __ldg(pmn + i), __ldg(ppl + i));
            dmin = dplus;
            if (i < n - 2) dplus = __ldg(gr + i + 2) - __ldg(gr + i + 1);
            else dplus = na;
            const double high = os_stalk1(xx - __ldg(gr + i + 1), v[l][1], v[l][2], v[l][3], dmin, dplus, a.degree[l],
                                          __ldg(det + i + 1), __ldg(pmn + i + 1), __ldg(ppl + i + 1));
            const double nx = (xx - __ldg(gr + i)) / (__ldg(gr + i + 1) - __ldg(gr + i));
            double w = 1 - nx;
            if (w == 0) ret = high;
            else if (w == 1) ret = low;
            else {
                w = os_blend(w, a.blend);
                ret = w * low + (1 - w) * high;
            }
            have_ret = true;
            ++l;  // pop
        }
        a.out[p] = ret;
    }
}

cudaError_t launch_oldstalk(const double *val, const double *grid, const double *det, const double *pmin,
                            const double *pplus, const int *goff, int rank, const int *dims, const int64_t *stride,
                            const double *degree, const int *uniform, int blend, const double *x, int64_t npts,
                            double *out, int num_sms, cudaStream_t s, LaunchInfo *li)
{
    OldStalkArgs a;
    a.val = val; a.grid = grid; a.det = det; a.pmin = pmin; a.pplus = pplus;
    a.rank = rank; a.blend = blend; a.x = x; a.npts = npts; a.out = out;
    for (int i = 0; i < rank; ++i) {
        a.goff[i] = goff[i]; a.dims[i] = dims[i]; a.stride[i] = stride[i]; a.degree[i] = degree[i]; a.uniform[i] = uniform[i];
    }
    const int block = 128;
    int64_t nb = (npts + block - 1) / block;
    if (nb > (int64_t)num_sms * 16) nb = (int64_t)num_sms * 16;
    if (nb < 1) nb = 1;
    oldstalk_kernel<<<(int)nb, block, 0, s>>>(a);
    if (li) { li->grid = (int)nb; li->block = block; li->smem = 0; li->kernel = K_OLDSTALK; }
    return cudaGetLastError();
}
